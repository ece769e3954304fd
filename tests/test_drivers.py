"""Batched driver adapters (pysisyphus ChainOfStates hook, ASE shim for DMF) with stand-in
images; the evaluator is the toy spring potential (tests/helpers.py)."""
import numpy as np

from pdb2reaction_b200 import uma_pysis, EV2AU, F_EVAA_2_AU
from pdb2reaction_b200.drivers import (SharedImageBatch, UMAASECalculator, batched_calculate_forces,
                                        install_batched_cos, recompute_energies)
from pdb2reaction_b200.shims import ANG2BOHR
from helpers import SpringBackend

ELEM = ["C", "H", "H", "O"]
X = np.array([[0, 0, 0], [1.1, 0, 0], [0, 1.0, 0.2], [0.3, -0.9, 0.8]], dtype=np.float64)


class Geom:                       # pysisyphus.Geometry-like
    def __init__(self, atoms, coords):
        self.atoms, self.coords = atoms, np.asarray(coords, dtype=np.float64).reshape(-1)
        self._energy = self._forces = None

    def set_results(self, r):
        self._energy, self._forces = r["energy"], r["forces"]


class COS:
    fix_first = fix_last = True

    def __init__(self, images):
        self.images, self.counter = images, 0


class Atoms:                      # ase.Atoms-like
    def __init__(self, sym, pos):
        self.sym, self.pos, self.calc = list(sym), np.array(pos, dtype=np.float64), None

    def copy(self):
        return Atoms(self.sym, self.pos.copy())

    def get_chemical_symbols(self):
        return self.sym

    def get_positions(self):
        return self.pos

    def get_potential_energy(self):
        return self.calc.get_potential_energy(self)

    def get_forces(self):
        return self.calc.get_forces(self)


def _imgs(n):
    return [X + 0.05 * k * np.array([[0, 0, 1], [0, 1, 0], [1, 0, 0], [0, 0, 0]]) for k in range(n)]


def test_chain_of_states_hook_batches_all_moving_images():
    be = SpringBackend()
    calc = uma_pysis(_backend=be)
    geoms = [Geom(ELEM, x * ANG2BOHR) for x in _imgs(5)]
    cos = install_batched_cos(COS(geoms), calc)
    res = cos.calculate_forces()
    assert be.calls == [(5, True)]                      # endpoints included the first time, ONE call
    res = cos.calculate_forces()
    assert be.calls[-1] == (3, True) and cos.counter == 2    # fixed endpoints are not re-evaluated
    for g, x in zip(geoms, _imgs(5)):
        ref = calc.get_forces(ELEM, x * ANG2BOHR)
        assert abs(g._energy - ref["energy"]) < 1e-12 and np.allclose(g._forces, ref["forces"], atol=1e-12)
    assert len(res["energy"]) == 5
    assert len(batched_calculate_forces(geoms, calc, skip_fixed=[0, 1, 2, 3, 4])) == 0


def test_ase_shim_units_and_shared_batch():
    be = SpringBackend()
    calc = uma_pysis(_backend=be)
    single = Atoms(ELEM, X)
    single.calc = UMAASECalculator(calc)
    e_ev, f_ev = be.evaluate(X[None])
    assert abs(single.get_potential_energy() - e_ev[0]) < 1e-9
    assert np.allclose(single.get_forces(), f_ev[0], atol=1e-6)
    images = [Atoms(ELEM, x) for x in _imgs(6)]
    SharedImageBatch(images, calc)
    be.calls.clear()
    es = [im.get_potential_energy() for im in images]
    fs = [im.get_forces() for im in images]
    assert be.calls == [(6, True)]                       # 12 queries, one batched evaluation
    images[2].pos = images[2].pos + 0.01
    images[2].get_forces()
    assert be.calls[-1] == (6, True) and len(be.calls) == 2
    ref_e, ref_f = be.evaluate(np.stack(_imgs(6)))
    assert np.allclose(es, ref_e, atol=1e-9) and np.allclose(np.stack(fs), ref_f, atol=1e-6)


def test_ase_shim_never_serves_stale_results_after_an_in_place_move():
    """ASE caches results against a COPY of the Atoms (check_state).  A calculator that aliased the caller's live
    object would compare it with itself and return the first geometry's results forever."""
    be = SpringBackend()
    calc = uma_pysis(_backend=be)
    at = Atoms(ELEM, X)
    at.calc = UMAASECalculator(calc)
    e0, f0 = at.get_potential_energy(), at.get_forces().copy()
    n0 = len(be.calls)
    assert at.get_potential_energy() == e0 and len(be.calls) == n0         # unchanged geometry: served from the cache
    assert at.calc.atoms is not at
    at.pos[1, 0] += 0.2                                                     # in-place move of the SAME object
    e1, f1 = at.get_potential_energy(), at.get_forces()
    assert len(be.calls) == n0 + 1 and e1 != e0 and np.abs(f1 - f0).max() > 1e-3
    ref_e, ref_f = be.evaluate(at.pos[None])
    assert abs(e1 - ref_e[0]) < 1e-9 and np.allclose(f1, ref_f[0], atol=1e-6)
    # shared batch: images moved in place are re-evaluated together, once
    images = [Atoms(ELEM, x) for x in _imgs(4)]
    SharedImageBatch(images, calc)
    [im.get_forces() for im in images]
    be.calls.clear()
    for im in images:
        im.pos += 0.01
    fs = np.stack([im.get_forces() for im in images])
    assert be.calls == [(4, True)]
    assert np.allclose(fs, be.evaluate(np.stack([im.pos for im in images]))[1], atol=1e-6)


def test_recompute_energies_is_one_call():
    be = SpringBackend()
    calc = uma_pysis(_backend=be)
    frames = np.stack(_imgs(7))
    e = recompute_energies(calc, ELEM, frames)
    assert be.calls == [(7, False)] and e.shape == (7,)
    assert np.allclose(e, be.evaluate(frames, forces=False)[0] * EV2AU)
