"""-m gpu: the CUDA path, called through the C ABI, against the oracle (same seeded inputs),
the committed golden vectors, and size-independent properties at BASELINE.json's full sizes.

Tolerances are the north star's: neighbour lists bit-exact; energies 1e-5 eV/atom; forces
1e-4 eV/A (float32 model arithmetic on both sides; forces are O(1) eV/A by construction of the
random-init weights, so the absolute force tolerance is also a 1e-4 relative one)."""
import glob
import os

import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation

from pdb2reaction_b200 import synth
from conftest import merged_for

pytestmark = pytest.mark.gpu

TOL_E_PER_ATOM = 1e-5   # eV/atom
TOL_F = 1e-4            # eV/A
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _engine(state4, arch4, elem, **kw):
    from pdb2reaction_b200.engine import UmabEngine
    z, merged = merged_for(state4, arch4, elem)
    return UmabEngine(merged, z, arch4, **kw), z, merged


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_cuda_reproduces_golden_vectors(path, built_lib, state4, arch4):
    g = np.load(path)
    elem = [str(e) for e in g["elem"]]
    eng, z, _ = _engine(state4, arch4, elem)
    coords = g["coords"]
    pos = torch.from_numpy(coords.astype(np.float32)).cuda()
    ei = eng.graph(pos).numpy()
    assert np.array_equal(ei, g["edge_index"].astype(np.int64))
    e, f = eng.energy_forces(pos)
    assert np.abs(e.cpu().numpy() - g["energy"]).max() / len(z) < TOL_E_PER_ATOM
    assert np.abs(f.cpu().numpy().astype(np.float64) - g["forces"]).max() < TOL_F
    # the host-buffer entry point gives the same bits as the device-pointer one
    e2, f2 = eng.energy_forces_host(coords.astype(np.float32))
    assert np.array_equal(e2, e.cpu().numpy()) and np.array_equal(f2, f.cpu().numpy())
    # energy-only call (no backward pass) returns the same energy
    e3, f3 = eng.energy_forces(pos, forces=False)
    assert f3 is None and torch.equal(e3, e)


@pytest.mark.parametrize("n,b,seed", [(20, 1, 1), (300, 2, 2), (500, 1, 3)])
def test_cuda_matches_oracle_on_seeded_clusters(n, b, seed, built_lib, state4, arch4, hyper4):
    from oracle import uma_ref
    elem, imgs = synth.make_string(n, b, seed)
    eng, z, merged = _engine(state4, arch4, elem)
    orc = uma_ref.OracleUMA(merged, z, dtype=torch.float32, hyper=hyper4, edge_chunk=8192)
    e_ref, f_ref = orc.energy_forces(imgs)
    e, f = eng.energy_forces_host(imgs.astype(np.float32))
    assert np.abs(e - e_ref.double().numpy()).max() / n < TOL_E_PER_ATOM
    assert np.abs(f - f_ref.numpy()).max() < TOL_F


def test_edge_chunking_does_not_change_the_result(built_lib, state4, arch4):
    elem, imgs = synth.make_string(60, 3, 11)
    eng, _, _ = _engine(state4, arch4, elem)
    eng_small, _, _ = _engine(state4, arch4, elem, workspace_bytes=9600 * 4 * 1500)
    pos = imgs.astype(np.float32)
    e1, f1 = eng.energy_forces_host(pos)
    e2, f2 = eng_small.energy_forces_host(pos)
    # chunk boundaries only change which launch an edge belongs to, never the arithmetic
    assert np.array_equal(e1, e2) and np.array_equal(f1, f2)


def test_closed_chunks_equal_open_chunks(built_lib, state4, arch4):
    """Whole-image (closed) chunks reduce the source halves of the edge adjoint by source node inside the chunk (no
    per-edge G buffer); node-range (open) chunks keep G + source_reduce.  Same additions in the same order: same bits."""
    elem, imgs = synth.make_string(60, 5, 21)
    pos = imgs.astype(np.float32)
    eng, _, _ = _engine(state4, arch4, elem)                                   # one closed chunk
    n_e = eng.graph(torch.from_numpy(pos).cuda())[0].shape[0]
    per_img = n_e // 5
    per_edge = 9600 * 4                                                          # upper bound of the workspace per edge
    eng_two, _, _ = _engine(state4, arch4, elem, workspace_bytes=int(2.5 * per_img) * per_edge)    # closed, 2 images / chunk
    eng_open, _, _ = _engine(state4, arch4, elem, workspace_bytes=int(0.4 * per_img) * per_edge)   # open: image > chunk
    e0, f0 = eng.energy_forces_host(pos)
    for other in (eng_two, eng_open):
        e1, f1 = other.energy_forces_host(pos)
        assert np.array_equal(e0, e1) and np.array_equal(f0, f1)


def test_store_mode_equals_recompute_mode(built_lib, state4, arch4):
    """Keeping the conv outputs for the backward (store mode) or recomputing them must give the
    same bits: the same kernels run on the same data in the same order."""
    elem, imgs = synth.make_string(150, 3, 12)
    eng_store, _, _ = _engine(state4, arch4, elem)                      # auto: fits -> store mode
    eng_rec, _, _ = _engine(state4, arch4, elem, store_bytes=-1)       # never store
    eng_rec_chunked, _, _ = _engine(state4, arch4, elem, store_bytes=-1, workspace_bytes=9600 * 4 * 4000)
    eng_store_chunked, _, _ = _engine(state4, arch4, elem, workspace_bytes=9600 * 4 * 4000)
    pos = imgs.astype(np.float32)
    e0, f0 = eng_store.energy_forces_host(pos)
    for eng in (eng_rec, eng_rec_chunked, eng_store_chunked):
        e1, f1 = eng.energy_forces_host(pos)
        assert np.array_equal(e0, e1) and np.array_equal(f0, f1)


def test_stagewise_intermediates_match_the_staged_twin(built_lib, state4, arch4, hyper4):
    from oracle import staged, uma_ref
    elem, imgs = synth.make_string(24, 2, 7)
    eng, z, merged = _engine(state4, arch4, elem, debug=True)
    orc = uma_ref.OracleUMA(merged, z, dtype=torch.float32, hyper=hyper4)
    pos, zz, nat, ei = orc._prep(imgs)
    _, _, inter = staged.energy_forces(merged, pos, zz, nat, ei, keep=True)
    eng.energy_forces_host(imgs.astype(np.float32))

    def close(name, ref, rtol=2e-5):
        t = eng.debug_tensor(name).double().reshape(-1)
        r = ref.double().reshape(-1)
        assert t.numel() == r.numel(), name
        assert (t - r).abs().max() <= rtol * r.abs().max() + 1e-7, name

    close("gauss", inter["geo"]["gauss"])
    close("env", inter["geo"]["env"])
    close("x0", inter["x0"])
    wig = eng.debug_tensor("wig").reshape(-1, 36)[:, :34]
    assert (wig - inter["geo"]["wig"]).abs().max() < 2e-6
    for l in range(4):
        d = inter[f"l{l}"]
        for k in ("n1", "rad", "y0", "y1", "y2", "x1", "x"):
            close(f"l{l}.{k}", d[k])
        close(f"l{l}.g_x", d["g_x_in"])
    close("g_gauss", inter["g_gauss"])
    close("g_env", inter["g_env"])
    close("g_vec", inter["g_vec"])


def test_full_size_properties_c4_image(built_lib, state4, arch4):
    """1500-atom images (BASELINE config 4): invariances that need no oracle at this size."""
    elem, imgs = synth.make_string(1500, 2, 4)
    eng, z, _ = _engine(state4, arch4, elem)
    pos = imgs.astype(np.float32)
    e, f = eng.energy_forces_host(pos)
    assert np.isfinite(e).all() and np.isfinite(f).all()
    # translation invariance -> forces of every image sum to zero
    assert np.abs(f.astype(np.float64).sum(axis=1)).max() < 5e-3
    # rigid rotation + translation: energy invariant, forces co-rotate
    rot = Rotation.random(random_state=2).as_matrix()
    pos_r = (imgs @ rot.T + np.array([0.3, -1.1, 2.0])).astype(np.float32)
    e_r, f_r = eng.energy_forces_host(pos_r)
    assert np.abs(e_r - e).max() / 1500 < TOL_E_PER_ATOM
    assert np.abs(f_r - f @ rot.T.astype(np.float32)).max() < 5 * TOL_F
    # an image evaluated alone equals the same image inside a batch, bit for bit
    e1, f1 = eng.energy_forces_host(pos[1:2])
    assert np.array_equal(e1[0], e[1]) and np.array_equal(f1[0], f[1])
    # image order does not matter
    e_s, f_s = eng.energy_forces_host(pos[::-1].copy())
    assert np.array_equal(e_s[::-1], e) and np.array_equal(f_s[::-1], f)
    # run-to-run reproducibility (no atomics on float data)
    e_2, f_2 = eng.energy_forces_host(pos)
    assert np.array_equal(e_2, e) and np.array_equal(f_2, f)


def test_forces_are_the_gradient_of_the_cuda_energy(built_lib, state4, arch4):
    """Central differences of the CUDA energy (fp64-accumulated) against the CUDA forces."""
    elem, coords = synth.make_cluster(40, 8)
    eng, _, _ = _engine(state4, arch4, elem)
    e0, f0 = eng.energy_forces_host(coords[None].astype(np.float32))
    h = 2e-2   # large step: the energy is float32 per atom
    rng = np.random.default_rng(0)
    d = rng.normal(size=coords.shape)
    d /= np.linalg.norm(d)
    batch = np.stack([coords + h * d, coords - h * d, coords + 2 * h * d, coords - 2 * h * d]).astype(np.float32)
    e, _ = eng.energy_forces_host(batch, forces=False)
    fd = (8 * (e[0] - e[1]) - (e[2] - e[3])) / (12 * h)           # 4th-order directional derivative
    an = -(f0[0].astype(np.float64) * d).sum()
    assert abs(fd - an) < 2e-3 * max(1.0, abs(an))
