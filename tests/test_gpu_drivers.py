"""-m gpu: the driver adapters either side of the calculator (pysisyphus ChainOfStates hook, ASE calculator /
shared image batch for DMF, energy re-evaluation loops) on the REAL CUDA backend against the CPU oracle.
Reference call sites: path_opt.py:184,952-977 (GSM, one image per call), path_opt.py:355-363,418-426 (DMF: a
FAIRChemCalculator per image), trj2fig.py:124-129."""
import numpy as np
import pytest
import torch

from pdb2reaction_b200 import EV2AU, F_EVAA_2_AU, synth, uma_pysis
from pdb2reaction_b200.drivers import (SharedImageBatch, UMAASECalculator, install_batched_cos, recompute_energies)
from pdb2reaction_b200.shims import ANG2BOHR
from conftest import merged_for
from test_drivers import COS, Atoms, Geom

pytestmark = pytest.mark.gpu
N, B = 130, 5                         # >= 100 atoms: tensor-core GEMMs + cell-list neighbour search


@pytest.fixture()
def case(state4, arch4, hyper4):
    from oracle import uma_ref
    elem, imgs = synth.make_string(N, B, 41)
    z, merged = merged_for(state4, arch4, elem)
    e, f = uma_ref.OracleUMA(merged, z, dtype=torch.float32, hyper=hyper4, edge_chunk=8192).energy_forces(imgs)
    return elem, imgs, e.double().numpy(), f.double().numpy()


def test_chain_of_states_hook_on_cuda(built_lib, small_model, case):
    elem, imgs, e_ref, f_ref = case
    calc = uma_pysis(model="test-4x", freeze_atoms=[0, 7])
    cos = install_batched_cos(COS([Geom(elem, x * ANG2BOHR) for x in imgs]), calc)
    res = cos.calculate_forces()
    f_ref = f_ref.copy()
    f_ref[:, [0, 7]] = 0.0
    for k, g in enumerate(cos.images):
        assert abs(g._energy - e_ref[k] * EV2AU) < 1e-5 * N * EV2AU
        assert np.abs(g._forces - f_ref[k].reshape(-1) * F_EVAA_2_AU).max() < 1e-4 * F_EVAA_2_AU
    # second cycle: fixed endpoints keep their results, the moving images are re-evaluated as one batch
    first, last = cos.images[0]._forces.copy(), cos.images[-1]._forces.copy()
    eng = calc._core.backend.engines[0]
    cos.calculate_forces()
    assert eng.graph_counts()[0] == (B - 2) * N
    assert np.array_equal(cos.images[0]._forces, first) and np.array_equal(cos.images[-1]._forces, last)
    assert len(res["energy"]) == B


def test_ase_calculator_and_shared_image_batch_on_cuda(built_lib, small_model, case):
    elem, imgs, e_ref, f_ref = case
    calc = uma_pysis(model="test-4x")
    single = Atoms(elem, imgs[1])
    single.calc = UMAASECalculator(calc)
    assert abs(single.get_potential_energy() - e_ref[1]) < 1e-5 * N           # eV
    assert np.abs(single.get_forces() - f_ref[1]).max() < 1e-4                 # eV/A
    images = [Atoms(elem, x) for x in imgs]
    SharedImageBatch(images, calc)
    es = np.array([im.get_potential_energy() for im in images])
    fs = np.stack([im.get_forces() for im in images])
    assert calc._core.backend.engines[0].graph_counts()[0] == B * N             # ONE batched evaluation served all
    assert np.abs(es - e_ref).max() < 1e-5 * N and np.abs(fs - f_ref).max() < 1e-4
    # in-place move of one image: the next query re-evaluates, results follow the geometry
    images[2].pos[5] += np.array([0.05, -0.02, 0.03])
    f2 = images[2].get_forces()
    assert np.abs(f2 - fs[2]).max() > 1e-3
    moved = calc.get_forces_batch(elem, images[2].pos.reshape(1, -1) * ANG2BOHR)["forces"][0] / F_EVAA_2_AU
    assert np.array_equal(f2.reshape(-1), moved)


def test_recompute_energies_on_cuda(built_lib, small_model, case):
    elem, imgs, e_ref, _ = case
    e = recompute_energies(uma_pysis(model="test-4x"), elem, imgs)
    assert np.abs(e - e_ref * EV2AU).max() < 1e-5 * N * EV2AU
