"""-m gpu: VALUES at BASELINE.json's headline sizes and on the tensor-core Hessian path, against float64-oracle
fixtures committed under tests/golden/large/ (tests/golden/make_large_golden.py; inputs are regenerated from the
seeded generator, only outputs are stored).

  C4 image  1500 atoms   E within 1e-5 eV/atom, F within 1e-4 eV/A, edge list checksum bit-exact
  C5 image  10 000 atoms same, through one closed chunk AND through open chunks (image larger than the workspace)
  Hessian   160 atoms (>= 100: tcgen05 GEMMs, cell-list neighbour search): analytic dual-number columns and the
            calculator's finite-difference columns against fp64 double-backward columns
"""
import os

import numpy as np
import pytest
import torch

from pdb2reaction_b200 import synth, uma_pysis
from pdb2reaction_b200.shims import ANG2BOHR
from conftest import merged_for

pytestmark = pytest.mark.gpu

LARGE = os.path.join(os.path.dirname(__file__), "golden", "large")
TOL_E_PER_ATOM = 1e-5   # eV/atom
TOL_F = 1e-4            # eV/A


def _engine(state4, arch4, elem, **kw):
    from pdb2reaction_b200.engine import UmabEngine
    z, merged = merged_for(state4, arch4, elem)
    return UmabEngine(merged, z, arch4, **kw)


def _check_image(eng, g, img, n):
    pos = torch.from_numpy(img[None].astype(np.float32)).cuda()
    ei = eng.graph(pos).numpy()
    assert ei.shape[1] == int(g["n_edges"])
    assert int((ei[0] * 1000003 + ei[1]).sum()) == int(g["edge_checksum"])          # same (sorted) edge list
    e, f = eng.energy_forces_host(img[None].astype(np.float32))
    de = abs(e[0] - g["energy"][0]) / n
    df = np.abs(f[0].astype(np.float64) - g["forces"]).max()
    assert de < TOL_E_PER_ATOM and df < TOL_F, (de, df)
    return e, f


def test_c4_image_values_match_the_float64_oracle(built_lib, state4, arch4):
    g = np.load(os.path.join(LARGE, "c4_n1500.npz"))
    elem, imgs = synth.make_config("C4")
    eng = _engine(state4, arch4, elem)
    e, f = _check_image(eng, g, imgs[int(g["image"])], 1500)
    # the same image inside the full 32-image string (one product call, sub-batched by the engine): same bits
    e32, f32 = eng.energy_forces_host(imgs.astype(np.float32))
    k = int(g["image"])
    assert np.array_equal(e32[k], e[0]) and np.array_equal(f32[k], f[0])


def test_whole_c4_string_matches_the_float64_oracle(built_lib, small_model):
    """North-star target: "reproduces reference energies, forces ... on a 32-image, 1500-atom DMF string".  All 32 images
    through the PUBLIC calculator call (one get_forces_batch, sub-batched inside the library) against float64-oracle
    values of every image (fixture: 56 CPU-minutes, tests/golden/make_large_golden.py string)."""
    path = os.path.join(LARGE, "c4_string_n1500_b32.npz")
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    from pdb2reaction_b200 import EV2AU, F_EVAA_2_AU
    g = np.load(path)
    elem, imgs = synth.make_config("C4")
    calc = uma_pysis(model="test-4x")
    r = calc.get_forces_batch(elem, imgs.reshape(32, -1) * ANG2BOHR)
    de = np.abs(r["energy"] / EV2AU - g["energy"]) / 1500
    df = np.abs(r["forces"].reshape(32, 1500, 3) / F_EVAA_2_AU - g["forces"].astype(np.float64)).max(axis=(1, 2))
    assert de.max() < TOL_E_PER_ATOM and df.max() < TOL_F, (de.max(), df.max())
    assert calc._core.backend.engines[0].last_call_subcalls >= 1


def test_c5_image_values_closed_and_open_chunks(built_lib, state4, arch4):
    g = np.load(os.path.join(LARGE, "c5_n10000.npz"))
    elem, imgs = synth.make_config("C5")
    img = imgs[int(g["image"])]
    eng = _engine(state4, arch4, elem)                                   # default workspace: one closed chunk
    e, f = _check_image(eng, g, img, 10000)
    eng.close()
    # 6 GB workspace < one image's edges: node-range (open) chunks, per-edge G buffer + source_reduce
    eng_open = _engine(state4, arch4, elem, workspace_bytes=6 << 30)
    e2, f2 = _check_image(eng_open, g, img, 10000)
    assert np.array_equal(e, e2) and np.array_equal(f, f2)


@pytest.fixture()
def hess_case():
    g = np.load(os.path.join(LARGE, "hess_n160.npz"))
    elem, coords = synth.make_cluster(int(g["n_atoms"]), int(g["seed"]))
    return g, elem, coords


def test_analytic_hessian_columns_on_the_tensor_core_path(built_lib, small_model, hess_case):
    g, elem, coords = hess_case
    calc = uma_pysis(model="test-4x", hessian_calc_mode="Analytical")
    calc._ensure_core(elem)
    eng = calc._core.backend.engines[0]
    assert eng.n_atoms >= 100                                            # auto mode -> tcgen05 GEMMs
    cols = [int(k) for k in g["cols"]]
    h = calc._core.backend.hessian_columns(coords, cols).astype(np.float64)          # [24, 480] eV/A^2
    ref = g["hessian_columns"]
    scale = np.abs(ref).max()
    assert np.abs(h - ref).max() < 2e-4 * scale, np.abs(h - ref).max() / scale
    r = calc.get_forces(elem, coords * ANG2BOHR)
    from pdb2reaction_b200 import EV2AU, F_EVAA_2_AU
    assert abs(r["energy"] - g["energy"][0] * EV2AU) < TOL_E_PER_ATOM * 160 * EV2AU
    assert np.abs(r["forces"] - g["forces"].reshape(-1) * F_EVAA_2_AU).max() < TOL_F * F_EVAA_2_AU


def test_fd_hessian_columns_on_the_tensor_core_path(built_lib, small_model, hess_case):
    g, elem, coords = hess_case
    calc = uma_pysis(model="test-4x")
    calc._ensure_core(elem)
    cols = [int(k) for k in g["cols"]]
    hmat = torch.zeros((480, 480), device="cuda", dtype=torch.float64)
    calc._fd_columns_into(hmat, coords, cols)
    h = hmat[:, cols].T.cpu().numpy()
    ref = g["hessian_columns"]
    # central differences with the reference's fixed step h = 1e-3 A amplify the force error by 1/h: the bf16x3 forces
    # carry ~2e-5 eV/A (tolerance 1e-4) -> ~1e-2 eV/A^2 noise on entries of up to 2.2 eV/A^2 (measured 9.5e-3); the
    # analytic columns above have no such amplification (2e-4 relative)
    assert np.abs(h - ref).max() < 2.5e-2, np.abs(h - ref).max()
    untouched = [k for k in range(480) if k not in cols]
    assert float(hmat[:, untouched].abs().max()) == 0.0
