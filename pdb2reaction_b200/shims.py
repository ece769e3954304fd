"""pysisyphus boundary: the real ``Calculator`` base class and constants when pysisyphus (the
``torch_pysis`` fork pinned by the reference, ``pyproject.toml:15``) is installed, minimal
stand-ins otherwise, so the same ``uma_pysis`` class drops into a real pdb2reaction install.

Reference: ``pdb2reaction/uma_pysis.py:122-124`` imports ``Calculator``, ``BOHR2ANG``,
``ANG2BOHR``, ``AU2EV`` and ``run``.  pysisyphus derives the constants from scipy's CODATA
table; the stand-ins do the same.
"""
from __future__ import annotations

try:  # pragma: no cover - exercised only where pysisyphus exists
    from pysisyphus.calculators.Calculator import Calculator  # type: ignore
    from pysisyphus.constants import BOHR2ANG, ANG2BOHR, AU2EV  # type: ignore
    HAVE_PYSISYPHUS = True
except Exception:  # ImportError or a broken optional dependency of pysisyphus
    HAVE_PYSISYPHUS = False
    import scipy.constants as _spc

    BOHR2ANG = _spc.value("Bohr radius") * 1e10
    ANG2BOHR = 1.0 / BOHR2ANG
    AU2EV = _spc.value("Hartree energy in eV")

    class Calculator:  # noqa: D401 - mirrors pysisyphus.calculators.Calculator.Calculator
        """Subset of the pysisyphus base class the UMA calculator relies on: stores
        ``charge``/``mult`` and bookkeeping kwargs (``mem``, ``pal``, ``out_dir`` ... are accepted
        and kept, e.g. the ``mem=`` the Dimer wrapper forwards, ``tsopt.py:745``)."""

        def __init__(self, calc_number=0, charge=0, mult=1, base_name="calculator", pal=1, mem=1000,
                     keep_kind="all", check_mem=True, retry_calc=0, last_calc_cycle=None,
                     clean_after=True, out_dir="qm_calcs", force_num_hess=False, num_hess_kwargs=None,
                     **unused):
            self.calc_number = calc_number
            self.charge = int(charge)
            self.mult = int(mult)
            self.base_name = base_name
            self.pal = pal
            self.mem = mem
            self.out_dir = out_dir
            self.force_num_hess = force_num_hess
            self.num_hess_kwargs = num_hess_kwargs or {}
            self.calc_counter = 0

        def get_energy(self, atoms, coords, **kw):
            raise NotImplementedError

        def get_forces(self, atoms, coords, **kw):
            raise NotImplementedError

        def get_hessian(self, atoms, coords, **kw):
            raise NotImplementedError
