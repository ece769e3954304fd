"""Batched call paths for the drivers either side of the calculator.

The reference drives every image of a string through the calculator ONE AT A TIME:
* GSM / pysisyphus: ``ChainOfStates.calculate_forces`` loops ``image.calc_energy_and_forces()``
  with ``scheduler: None`` (``pdb2reaction/path_opt.py:184, 952-977``);
* DMF / torch_dmf: every image owns an ASE calculator (``FAIRChemCalculator``) and the IPOPT
  objective calls ``image.get_potential_energy()`` / ``get_forces()`` per image
  (``pdb2reaction/path_opt.py:355-363, 418-426``);
* ``trj2fig.recompute_energies`` / final-energy loops (``trj2fig.py:124-129``, ``path_opt.py:434-438``).

These adapters collapse such loops into one ``get_forces_batch`` call.  pysisyphus and ASE are not
installed in the build container, so the adapters are duck-typed against the attributes those
libraries use and are unit-tested with stand-ins.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .calculator import uma_pysis, EV2AU, F_EVAA_2_AU
from .shims import ANG2BOHR


# ------------------------------------------------------------------ pysisyphus chain of states
def batched_calculate_forces(images: Sequence, calc: uma_pysis, skip_fixed: Sequence[int] = ()):
    """Evaluate all ``images`` (pysisyphus ``Geometry``-like: ``.atoms``, ``.cart_coords`` or
    ``.coords`` in Bohr) in one engine call and store the results on each image the way
    ``Geometry.set_results`` does.  Returns the list of result dicts."""
    idx = [i for i in range(len(images)) if i not in set(skip_fixed)]
    if not idx:
        return []
    atoms = images[idx[0]].atoms
    coords = np.stack([np.asarray(getattr(images[i], "cart_coords", images[i].coords), dtype=np.float64).reshape(-1)
                       for i in idx])
    res = calc.get_forces_batch(atoms, coords)
    out = []
    for k, i in enumerate(idx):
        r = {"energy": float(res["energy"][k]), "forces": res["forces"][k]}
        img = images[i]
        if hasattr(img, "set_results"):
            img.set_results(r)
        else:
            img._energy, img._forces = r["energy"], r["forces"]
        out.append(r)
    return out


def install_batched_cos(cos, calc: uma_pysis):
    """Replace ``cos.calculate_forces`` (pysisyphus ``ChainOfStates``) by a batched evaluation of
    the non-fixed images; everything else of the chain (tangents, climbing image, growth) is
    untouched."""
    def calculate_forces():
        fixed = []
        if getattr(cos, "fix_first", False):
            fixed.append(0)
        if getattr(cos, "fix_last", False):
            fixed.append(len(cos.images) - 1)
        todo = [i for i in fixed if getattr(cos.images[i], "_forces", None) is None]   # endpoints once
        batched_calculate_forces(cos.images, calc, skip_fixed=[i for i in fixed if i not in todo])
        if hasattr(cos, "counter"):
            cos.counter += 1
        return {"energy": [im._energy if hasattr(im, "_energy") else im.energy for im in cos.images],
                "forces": [im._forces if hasattr(im, "_forces") else im.forces for im in cos.images]}
    cos.calculate_forces = calculate_forces
    return cos


def recompute_energies(calc: uma_pysis, atoms: Sequence[str], frames_ang: np.ndarray) -> np.ndarray:
    """Energies (Hartree) of trajectory frames [B,N,3] in Angstrom, one call
    (``trj2fig.recompute_energies`` / ``path_opt.py:434-438`` equivalents)."""
    c = np.asarray(frames_ang, dtype=np.float64)
    return calc.get_energy_batch(atoms, c.reshape(c.shape[0], -1) * ANG2BOHR)["energy"]


# ------------------------------------------------------------------ ASE (DMF path)
try:  # pragma: no cover
    from ase.calculators.calculator import Calculator as _ASEBase, all_changes  # type: ignore
    HAVE_ASE = True
except Exception:
    HAVE_ASE = False
    all_changes = ["positions", "numbers", "cell", "pbc", "initial_charges", "initial_magmoms"]

    class _ASEBase:
        """Stand-in of ``ase.calculators.calculator.Calculator`` with ASE's caching protocol: ``get_property``
        compares the queried Atoms with the stored COPY (``check_state``), clears ``results`` on any change and
        calls ``calculate`` only when the property is not cached; the base ``calculate`` stores ``atoms.copy()``."""
        implemented_properties: List[str] = []

        def __init__(self, **kwargs):
            self.results = {}
            self.atoms = None

        def check_state(self, atoms, tol=1e-15):
            if self.atoms is None:
                return list(all_changes)
            changes = []
            a = np.asarray(self.atoms.get_positions(), dtype=np.float64)
            b = np.asarray(atoms.get_positions(), dtype=np.float64)
            if a.shape != b.shape or np.abs(a - b).max(initial=0.0) > tol:
                changes.append("positions")
            if list(self.atoms.get_chemical_symbols()) != list(atoms.get_chemical_symbols()):
                changes.append("numbers")
            return changes

        def get_property(self, name, atoms=None, allow_calculation=True):
            if atoms is None:
                atoms, system_changes = self.atoms, []
            else:
                system_changes = self.check_state(atoms)
                if system_changes:
                    self.results = {}
            if name not in self.results:
                if not allow_calculation:
                    return None
                self.calculate(atoms, [name], system_changes)
            return self.results[name]

        def calculate(self, atoms=None, properties=("energy",), system_changes=tuple(all_changes)):
            if atoms is not None:
                self.atoms = atoms.copy()

        def get_potential_energy(self, atoms=None, force_consistent=False):
            return self.get_property("energy", atoms)

        def get_forces(self, atoms=None):
            return self.get_property("forces", atoms)


class UMAASECalculator(_ASEBase):
    """ASE calculator (eV, eV/A) on the B200 backend: the stand-in for fairchem's
    ``FAIRChemCalculator(predictor, task_name=...)`` in ``_run_dmf_mep`` (``path_opt.py:355-363``)."""

    implemented_properties = ["energy", "forces"]

    def __init__(self, calc: Optional[uma_pysis] = None, **calc_kwargs):
        super().__init__()
        self.calc = calc if calc is not None else uma_pysis(**calc_kwargs)
        self._pool: Optional["SharedImageBatch"] = None
        self._slot = -1

    def calculate(self, atoms=None, properties=("energy",), system_changes=tuple(all_changes)):
        if atoms is None:
            atoms = self.atoms
        # ASE's base class stores atoms.copy(): the cache check of the next query (check_state) must compare the
        # new geometry with a SNAPSHOT, never with the caller's live object (reference: FAIRChemCalculator.calculate
        # calls Calculator.calculate first)
        super().calculate(atoms, properties, system_changes)
        if self._pool is not None:
            e, f = self._pool.result_for(self._slot)
        else:
            sym = atoms.get_chemical_symbols()
            pos = np.asarray(atoms.get_positions(), dtype=np.float64)
            r = self.calc.get_forces_batch(sym, pos.reshape(1, -1) * ANG2BOHR)
            e, f = r["energy"][0] / EV2AU, r["forces"][0].reshape(-1, 3) / F_EVAA_2_AU
        self.results = {"energy": float(e), "free_energy": float(e), "forces": np.asarray(f, dtype=np.float64)}


class SharedImageBatch:
    """All images of a DMF/NEB-like path share ONE batched evaluation: the first query after any
    image moved evaluates every image in one engine call; the per-image ASE calculators then
    serve from the cache (``for image in mxflx.images: image.calc = ...``, ``path_opt.py:418-423``)."""

    def __init__(self, images: Sequence, calc: uma_pysis):
        self.images = list(images)
        self.calc = calc
        self._pos = None
        self._e = None
        self._f = None
        for k, im in enumerate(self.images):
            c = UMAASECalculator(calc)
            c._pool, c._slot = self, k
            im.calc = c

    def _current(self):
        return np.stack([np.asarray(im.get_positions(), dtype=np.float64) for im in self.images])

    def result_for(self, slot: int):
        pos = self._current()
        if self._pos is None or pos.shape != self._pos.shape or not np.array_equal(pos, self._pos):
            sym = self.images[0].get_chemical_symbols()
            r = self.calc.get_forces_batch(sym, pos.reshape(pos.shape[0], -1) * ANG2BOHR)
            self._e = r["energy"] / EV2AU
            self._f = r["forces"].reshape(pos.shape[0], -1, 3) / F_EVAA_2_AU
            self._pos = pos.copy()
        return self._e[slot], self._f[slot]
