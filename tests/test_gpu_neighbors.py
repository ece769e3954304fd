"""-m gpu: K1 neighbour search, bit-exact against the oracle after the canonical sort, on every
BASELINE.json size and on the edge cases (cap binding, coincident atoms, empty graph, ragged tail)."""
import numpy as np
import pytest
import torch

from oracle import graph as ograph
from pdb2reaction_b200 import synth
from conftest import merged_for

pytestmark = pytest.mark.gpu


def _eng(state4, arch4, elem, **kw):
    from pdb2reaction_b200.engine import UmabEngine
    z, merged = merged_for(state4, arch4, elem)
    return UmabEngine(merged, z, arch4, **kw)


@pytest.mark.parametrize("mode", ["brute", "cell"])
@pytest.mark.parametrize("n,b,seed", [(20, 1, 1), (300, 12, 2), (500, 1, 3), (1500, 4, 4), (10000, 1, 5)])
def test_edge_index_bit_exact_on_baseline_sizes(n, b, seed, mode, built_lib, state4, arch4):
    """Both searches (brute force, shared-memory cell list) against the oracle, bit for bit."""
    elem, imgs = synth.make_string(n, b, seed)
    eng = _eng(state4, arch4, elem)
    eng.set_neighbor_mode(mode)
    pos = imgs.astype(np.float32)
    ei = eng.graph(torch.from_numpy(pos).cuda()).numpy()
    ref = ograph.radius_graph(pos.reshape(-1, 3), [n] * b, 6.0, 300)
    assert ei.shape == ref.shape and np.array_equal(ei, ref)
    assert np.array_equal(ei, ograph.canonical_sort(ei))


@pytest.mark.parametrize("mode", ["brute", "cell"])
@pytest.mark.parametrize("cap", [4, 8, 30])
def test_max_neighbors_cap_non_strict_rule(cap, mode, built_lib, state4, arch4):
    elem, imgs = synth.make_string(120, 2, 6)
    eng = _eng(state4, arch4, elem, max_neighbors=cap)
    eng.set_neighbor_mode(mode)
    pos = imgs.astype(np.float32)
    ei = eng.graph(torch.from_numpy(pos).cuda()).numpy()
    ref = ograph.radius_graph(pos.reshape(-1, 3), [120, 120], 6.0, cap)
    assert np.array_equal(ei, ref)
    assert np.bincount(ei[1]).max() >= cap + 1


@pytest.mark.parametrize("mode", ["brute", "cell"])
def test_degenerate_inputs(mode, built_lib, state4, arch4):
    # coincident atoms (d2 <= 1e-4) are not neighbours; atoms beyond the cutoff give an empty graph
    elem = ["H", "H", "C", "O"]
    eng = _eng(state4, arch4, elem)
    eng.set_neighbor_mode(mode)
    pos = np.array([[[0, 0, 0], [0, 0, 0.005], [0, 0, 1.0], [30, 0, 0]],
                    [[0, 0, 0], [50, 0, 0], [0, 50, 0], [0, 0, 50]]], dtype=np.float32)
    ei = eng.graph(torch.from_numpy(pos).cuda()).numpy()
    ref = ograph.radius_graph(pos.reshape(-1, 3), [4, 4], 6.0, 300)
    assert np.array_equal(ei, ref)
    assert (ei[0] < 4).all() and ei.shape[1] == 4
    e, f = eng.energy_forces_host(pos)
    assert np.isfinite(e).all() and np.isfinite(f).all()
    assert np.all(f[1] == 0.0)                         # isolated atoms feel no force
    # exactly on the cutoff: d2 == 36 is kept (<=)
    pos2 = np.array([[[0, 0, 0], [6.0, 0, 0], [0, 6.0000005, 0], [0, 0, -6.0]]], dtype=np.float32)
    ei2 = eng.graph(torch.from_numpy(pos2).cuda()).numpy()
    assert np.array_equal(ei2, ograph.radius_graph(pos2.reshape(-1, 3), [4], 6.0, 300))


def test_cell_list_on_sparse_and_elongated_images(built_lib, state4, arch4):
    """Grids that hit the cell-count cap (atoms spread over kilometres) or are one cell thick still give the
    brute-force edge list; so does a jittered lattice with many atoms exactly one cell edge apart."""
    rng = np.random.default_rng(7)
    n = 200
    elem = ["C"] * n
    eng = _eng(state4, arch4, elem)
    chain = np.zeros((n, 3)); chain[:, 0] = np.arange(n) * 1.4                       # 280 A long, one cell thick
    far = rng.uniform(-5, 5, (n, 3)); far[::7] += rng.uniform(-1e5, 1e5, (len(far[::7]), 3))   # outliers
    lattice = np.stack(np.meshgrid(*[np.arange(6) * 6.0] * 3, indexing="ij"), -1).reshape(-1, 3)[:n]
    lattice = lattice + rng.normal(0, 1e-4, lattice.shape)
    pos = np.stack([chain, far, lattice]).astype(np.float32)
    ref = ograph.radius_graph(pos.reshape(-1, 3), [n] * 3, 6.0, 300)
    for mode in ("brute", "cell"):
        eng.set_neighbor_mode(mode)
        ei = eng.graph(torch.from_numpy(pos).cuda()).numpy()
        assert np.array_equal(ei, ref), mode
