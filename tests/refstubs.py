"""Run the REFERENCE's own wrapper code (``/root/reference/pdb2reaction``) offline.

The reference's calculator wrapper (``uma_pysis.py``), its Hessian post-processing helpers
(``freq.py:122-381``) and its restraint decorator (``opt.py:286-343``) are plain Python; only their
third-party imports (``ase``, ``fairchem.core``, ``pysisyphus``) are missing in this container.  This module
installs minimal stand-ins for exactly the names those files import, loads the reference files UNMODIFIED
from where they lie (never copied into the repo) and hands the live reference objects to the tests:

* ``load_reference_uma_pysis(predict_factory)``  -> the reference module; ``pretrained_mlip.get_predict_unit``
  is answered by ``predict_factory`` (tests pass an oracle-backed predictor, so
  reference-wrapper(oracle) can be compared bit for bit with repo-calculator(oracle));
* ``load_reference_functions(path, names, namespace)`` -> functions / classes cut out of a reference file by
  ``ast`` and exec'd as they stand (for files whose module-level imports pull in the whole CLI).

Nothing here restates reference logic: the stand-ins implement third-party library surface only
(``ase.Atoms`` as a record, ``AtomicData.from_ase`` as a float32 cast + the graph options,
``pysisyphus.constants`` from scipy's CODATA table as pysisyphus itself does).
TEST INFRASTRUCTURE; needs /root/reference (absent on the GPU box -> callers skip).
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types
from contextlib import contextmanager

import numpy as np
import scipy.constants as spc
import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("PDB2REACTION_REFERENCE", "/root/reference")
REF_PKG = os.path.join(REFERENCE_ROOT, "pdb2reaction")
HAVE_REFERENCE = os.path.exists(os.path.join(REF_PKG, "uma_pysis.py"))

# pysisyphus.constants (pysisyphus/constants.py derives them from scipy.constants)
BOHR2ANG = spc.value("Bohr radius") * 1e10
ANG2BOHR = 1.0 / BOHR2ANG
AU2EV = spc.value("Hartree energy in eV")
AMU2AU = 1.0 / spc.value("electron mass in u")

# ase.units (CODATA 2014, ASE's default) -- the five numbers freq.py:358-361 reads
ASE_UNITS = types.SimpleNamespace(_hbar=1.054571800e-34, _e=1.6021766208e-19, _amu=1.660539040e-27,
                                  _c=299792458.0, _hplanck=6.626070040e-34)
ASE_UNITS.invcm = 100.0 * ASE_UNITS._c * ASE_UNITS._hplanck / ASE_UNITS._e

SYMBOLS = ["X", "H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P", "S", "Cl", "Ar"]


class Atoms:
    """``ase.Atoms`` as the reference uses it (uma_pysis.py:361: ``Atoms(elem, positions=coord_ang)``)."""

    def __init__(self, symbols, positions=None):
        self.symbols = [str(s) for s in symbols]
        self.numbers = np.array([SYMBOLS.index(s) for s in self.symbols], dtype=np.int64)
        self.positions = np.array(positions, dtype=np.float64).reshape(-1, 3)
        self.info = {}

    def __len__(self):
        return len(self.symbols)

    def get_positions(self):
        return self.positions.copy()


class _Data:
    def __init__(self, atoms, max_neigh, radius, r_edges):
        self.pos = torch.from_numpy(atoms.positions.astype(np.float32))      # AtomicData.from_ase: float32
        self.atomic_numbers = torch.from_numpy(atoms.numbers.copy())
        self.charge = int(atoms.info.get("charge", 0))
        self.spin = int(atoms.info.get("spin", 0))
        self.max_neigh, self.radius, self.r_edges = max_neigh, radius, r_edges
        self.dataset = None
        self.device = torch.device("cpu")

    def to(self, device):
        self.device = torch.device(device)
        return self


class AtomicData:
    @staticmethod
    def from_ase(atoms, max_neigh=None, radius=None, r_edges=False):
        return _Data(atoms, max_neigh, radius, r_edges)


def data_list_collater(data_list, otf_graph=False):
    assert len(data_list) == 1 and otf_graph is True            # the reference's only call (uma_pysis.py:322)
    return data_list[0]


class _Backbone(nn.Module):
    def __init__(self, max_neighbors, cutoff):
        super().__init__()
        self.max_neighbors, self.cutoff = max_neighbors, cutoff


class _Model(nn.Module):
    def __init__(self, max_neighbors, cutoff):
        super().__init__()
        self.backbone = _Backbone(max_neighbors, cutoff)
        self.drop = nn.Dropout(0.1)                # the reference sets p = 0 (uma_pysis.py:262-264)
        self.w = nn.Parameter(torch.zeros(1))      # so .parameters() is non-empty (uma_pysis.py:396-398)


class OraclePredictUnit:
    """``MLIPPredictUnit`` stand-in: ``predict(batch)`` = the CPU oracle on ``batch.pos`` with the graph rebuilt
    from the float32 positions at every call (``otf_graph``), forces by autograd, energy left attached to the
    autograd graph so the reference's ``torch.autograd.functional.hessian`` block works."""

    def __init__(self, weights_for, hyper, dtype=torch.float32):
        """weights_for(z, charge, spin, task) -> merged (or un-merged) oracle weight dict."""
        from oracle import uma_ref
        self._uma_ref = uma_ref
        self._weights_for = weights_for
        self.hp = hyper
        self.dtype = dtype
        self.model = _Model(hyper.max_neighbors, hyper.cutoff)
        self.n_predict = 0

    def predict(self, batch):
        from oracle import graph as ograph
        self.n_predict += 1
        z = [int(v) for v in batch.atomic_numbers]
        w = self._weights_for(z, batch.charge, batch.spin, batch.dataset)
        hp = self._uma_ref.Hyper(num_experts=self.hp.num_experts, cutoff=float(batch.radius),
                                 max_neighbors=int(batch.max_neigh))
        pos = batch.pos
        ei = torch.from_numpy(ograph.radius_graph(pos.detach().numpy().astype(np.float32), [len(z)],
                                                  hp.cutoff, hp.max_neighbors))
        with torch.enable_grad():
            p = pos if pos.requires_grad else pos.detach().requires_grad_(True)
            e = self._uma_ref.energy(w, p.to(self.dtype), torch.tensor(z), [len(z)], ei, charge=batch.charge,
                                     spin=batch.spin, task_name=batch.dataset, hp=hp)
            keep = torch.is_grad_enabled() and pos.requires_grad and pos.grad_fn is not None
            (g,) = torch.autograd.grad(e.sum(), p, create_graph=keep, retain_graph=True)
        return {"energy": e if pos.requires_grad else e.detach(), "forces": (-g).to(torch.float32)}


class _PysisCalculator:
    """pysisyphus.calculators.Calculator.Calculator: the constructor surface the reference relies on."""

    def __init__(self, calc_number=0, charge=0, mult=1, base_name="calculator", pal=1, mem=1000, **kw):
        self.calc_number, self.charge, self.mult = calc_number, int(charge), int(mult)
        self.base_name, self.pal, self.mem = base_name, pal, mem
        self.kwargs = kw


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


@contextmanager
def stubbed_third_party(predict_factory=None, parallel_cls=None):
    """sys.modules entries for exactly what uma_pysis.py:108-124 imports; removed again on exit."""
    run_mod = _module("pysisyphus.run", CALC_DICT={}, run=lambda: None)
    pm = _module("fairchem.core.pretrained_mlip")

    def get_predict_unit(model, device="cpu", workers=1):
        return predict_factory(model, device, workers)

    pm.get_predict_unit = get_predict_unit
    pm.pretrained_checkpoint_path_from_name = lambda name: f"/nonexistent/{name}.pt"
    pm.get_reference_energies = lambda name, reference_type=None: {}
    mods = {
        "ase": _module("ase", Atoms=Atoms),
        "fairchem": _module("fairchem"),
        "fairchem.core": _module("fairchem.core", pretrained_mlip=pm),
        "fairchem.core.pretrained_mlip": pm,
        "fairchem.core.datasets": _module("fairchem.core.datasets", data_list_collater=data_list_collater),
        "fairchem.core.datasets.atomic_data": _module("fairchem.core.datasets.atomic_data", AtomicData=AtomicData),
        "pysisyphus": _module("pysisyphus", run=run_mod),
        "pysisyphus.run": run_mod,
        "pysisyphus.calculators": _module("pysisyphus.calculators"),
        "pysisyphus.calculators.Calculator": _module("pysisyphus.calculators.Calculator", Calculator=_PysisCalculator),
        "pysisyphus.constants": _module("pysisyphus.constants", BOHR2ANG=BOHR2ANG, ANG2BOHR=ANG2BOHR, AU2EV=AU2EV,
                                        AMU2AU=AMU2AU),
    }
    if parallel_cls is not None:
        mods["fairchem.core.units"] = _module("fairchem.core.units")
        mods["fairchem.core.units.mlip_unit"] = _module("fairchem.core.units.mlip_unit")
        mods["fairchem.core.units.mlip_unit.predict"] = _module("fairchem.core.units.mlip_unit.predict",
                                                                ParallelMLIPPredictUnit=parallel_cls)
        mods["fairchem.core.units.mlip_unit.api"] = _module("fairchem.core.units.mlip_unit.api")
        mods["fairchem.core.units.mlip_unit.api.inference"] = _module(
            "fairchem.core.units.mlip_unit.api.inference", guess_inference_settings=lambda name: {"name": name})
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        yield mods
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load_reference_uma_pysis(predict_factory, parallel_cls=None):
    """Import /root/reference/pdb2reaction/uma_pysis.py as it stands (module name ``_ref_uma_pysis``)."""
    path = os.path.join(REF_PKG, "uma_pysis.py")
    with stubbed_third_party(predict_factory, parallel_cls):
        spec = importlib.util.spec_from_file_location("_ref_uma_pysis", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    return mod


def load_reference_functions(filename, names, namespace):
    """Cut the named top-level functions / classes out of a reference file (ast, source untouched) and exec
    them in ``namespace``.  Returns the namespace.  Used where a module's imports drag in the whole CLI."""
    path = os.path.join(REF_PKG, filename)
    with open(path, "r", encoding="utf-8") as fh:
        src = fh.read()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    missing = set(names) - {n.name for n in wanted}
    if missing:
        raise KeyError(f"{filename}: {sorted(missing)} not found")
    ns = dict(namespace)
    ns.setdefault("__builtins__", __builtins__)
    code = compile(ast.Module(body=wanted, type_ignores=[]), path, "exec")
    exec(code, ns)
    return ns
