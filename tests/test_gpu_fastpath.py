"""-m gpu: the launch-bound path of the engine -- sync-free graph build on capacity-sized edge arrays, overflow flag +
retry, CUDA-graph capture / replay of the host-buffer entry point -- returns the same BITS as the path that reads the
edge count back, for every call of a warm-up / capture / replay sequence.  (Reference calling pattern: optimizers call
the calculator serially, one geometry per call, path_opt.py:184,952-977.)"""
import numpy as np
import pytest
import torch

from pdb2reaction_b200 import synth
from conftest import merged_for

pytestmark = pytest.mark.gpu


def _engines(state4, arch4, elem):
    from pdb2reaction_b200.engine import UmabEngine
    z, merged = merged_for(state4, arch4, elem)
    fast = UmabEngine(merged, z, arch4)
    slow = UmabEngine(merged, z, arch4)
    slow.set_option("nosync", 0)
    slow.set_option("cuda_graphs", 0)
    assert fast.get_option("nosync") == 1 and fast.get_option("cuda_graphs") == 1 and slow.get_option("nosync") == 0
    return fast, slow


@pytest.mark.parametrize("n,b", [(20, 1), (20, 3), (64, 2), (130, 1), (300, 4)])
def test_graph_replay_equals_the_synchronising_path_bit_for_bit(n, b, built_lib, state4, arch4):
    elem, imgs = synth.make_string(n, b, 50 + n)
    fast, slow = _engines(state4, arch4, elem)
    rng = np.random.default_rng(n)
    for k in range(6):                               # call 0 warms up, call 1 captures, calls 2.. replay
        pos = (imgs + 0.03 * k * rng.normal(size=imgs.shape)).astype(np.float32)
        e1, f1 = fast.energy_forces_host(pos)
        e0, f0 = slow.energy_forces_host(pos)
        assert np.array_equal(e0, e1) and np.array_equal(f0, f1), k
        assert fast.last_call_edges == slow.last_call_edges and fast.graph_counts()[1] == slow.graph_counts()[1]
        en1, none = fast.energy_forces_host(pos, forces=False)        # energy-only calls have graphs of their own
        assert none is None and np.array_equal(en1, e0)
    assert fast.get_option("graph_captures") >= 1 and fast.get_option("graph_replays") >= 3
    assert slow.get_option("graph_captures") == 0
    # the device-pointer entry point takes the sync-free path too (no graph: caller-owned buffers)
    p = torch.from_numpy(pos).cuda()
    e2, f2 = fast.energy_forces(p)
    assert np.array_equal(e2.cpu().numpy(), e0) and np.array_equal(f2.cpu().numpy(), f0)
    # launches are still accounted for when a graph is replayed
    l0 = fast.stats()["kernel_launches"]
    fast.energy_forces_host(pos)
    assert fast.stats()["kernel_launches"] - l0 > 100


def test_edge_capacity_overflow_is_detected_and_retried(built_lib, state4, arch4):
    """>= 128 atoms: the capacity comes from the history.  A denser geometry than any seen before overflows it; the
    flag is raised on the device, the call repeats with the larger capacity, and the results are those of the
    synchronising path."""
    elem, coords = synth.make_cluster(200, 77)
    fast, slow = _engines(state4, arch4, elem)
    sparse = (coords * 1.25).astype(np.float32)[None]
    dense = (coords * 0.97).astype(np.float32)[None]
    for _ in range(3):
        fast.energy_forces_host(sparse)
    assert fast.get_option("overflow_retries") == 0
    e1, f1 = fast.energy_forces_host(dense)
    e0, f0 = slow.energy_forces_host(dense)
    assert fast.get_option("overflow_retries") >= 1
    assert slow.graph_counts()[1] > 1.1 * fast.get_option("edges_per_image_seen") / 1.5     # really denser
    assert np.array_equal(e0, e1) and np.array_equal(f0, f1)
    # device-pointer entry: the overflow surfaces through umab_last_call and the wrapper repeats the call
    fast2, _ = _engines(state4, arch4, elem)
    for _ in range(2):
        fast2.energy_forces(torch.from_numpy(sparse).cuda())
    e2, f2 = fast2.energy_forces(torch.from_numpy(dense).cuda())
    assert np.array_equal(e2.cpu().numpy(), e0) and np.array_equal(f2.cpu().numpy(), f0)
    # analytic Hessian columns (dual numbers) on the sync-free path
    t = torch.zeros(1, 200, 3, device="cuda")
    t[0, 5, 2] = 1.0
    fa, dfa = fast.forces_jvp(torch.from_numpy(dense).cuda(), t)
    fb, dfb = slow.forces_jvp(torch.from_numpy(dense).cuda(), t)
    assert torch.equal(fa, fb) and torch.equal(dfa, dfb)


def test_batch_composition_never_changes_an_image(built_lib, state4, arch4):
    """An image alone == the same image inside a batch, bit for bit, also between 100 and 256 atoms where the node-level
    GEMMs of a single image are small (the GEMM kernel choice depends on the image size only, never on the batch)."""
    for n in (60, 130, 260):
        elem, imgs = synth.make_string(n, 4, 90 + n)
        fast, slow = _engines(state4, arch4, elem)
        pos = imgs.astype(np.float32)
        e, f = slow.energy_forces_host(pos)
        for eng in (fast, slow):
            e1, f1 = eng.energy_forces_host(pos[2:3])
            assert np.array_equal(e1[0], e[2]) and np.array_equal(f1[0], f[2]), n


def test_release_workspace_gives_memory_back_and_results_do_not_change(built_lib, state4, arch4):
    elem, imgs = synth.make_string(300, 3, 5)
    fast, slow = _engines(state4, arch4, elem)
    slow.close()
    pos = imgs.astype(np.float32)
    e0, f0 = fast.energy_forces_host(pos)
    e0, f0 = fast.energy_forces_host(pos)
    before = fast.stats()["device_bytes"]
    fast.release_workspace()
    after = fast.stats()["device_bytes"]
    assert after < 0.2 * before                       # weights + graph arrays stay, the per-call buffers are gone
    for _ in range(3):                                # re-grow, capture, replay
        e1, f1 = fast.energy_forces_host(pos)
        assert np.array_equal(e0, e1) and np.array_equal(f0, f1)


@pytest.mark.parametrize("n,b", [(130, 2), (300, 3), (1500, 2)])
def test_fused_gate_epilogue_is_bit_identical_to_the_separate_combine_kernel(n, b, built_lib, state4, arch4):
    """conv-1 m = +-1 / +-2 GEMMs write the gated bf16 planes of conv-2 from their epilogue (gemm_tc2.cu): the same fp32
    product and the same split as combine_gate_fwd_kernel -> identical energies and forces, also with chunked /
    recomputing backward passes and for energy-only calls."""
    elem, imgs = synth.make_string(n, b, 70 + n)
    fused, plain = _engines(state4, arch4, elem)
    for eng in (fused, plain):
        eng.set_option("nosync", 1)
        eng.set_option("cuda_graphs", 0)
    fused.set_option("fuse_gate", 1)                  # off by default (measured: no gain)
    plain.set_option("fuse_gate", 0)
    assert fused.get_option("fuse_gate") == 1 and plain.get_option("fuse_gate") == 0
    pos = imgs.astype(np.float32)
    e0, f0 = plain.energy_forces_host(pos)
    e1, f1 = fused.energy_forces_host(pos)
    assert np.array_equal(e0, e1) and np.array_equal(f0, f1)
    e2, _ = fused.energy_forces_host(pos, forces=False)
    assert np.array_equal(e2, e0)
    from pdb2reaction_b200.engine import UmabEngine
    z, merged = merged_for(state4, arch4, elem)
    rec = UmabEngine(merged, z, arch4, store_bytes=-1, workspace_bytes=9600 * 4 * 6000)      # recompute + several chunks
    e3, f3 = rec.energy_forces_host(pos)
    assert np.array_equal(e0, e3) and np.array_equal(f0, f3)
