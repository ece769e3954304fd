"""Image sharding: partition arithmetic, record packing and the world_size-2 gloo path of
SpmdEvaluator (the N>1 bench path; NCCL on the GPU box)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pdb2reaction_b200.sharding import SpmdEvaluator, pack_results, shard_bounds, unpack_results


def test_shard_bounds_cover_everything_contiguously():
    for n in (0, 1, 7, 12, 32, 33):
        for w in (1, 2, 3, 4, 8):
            b = shard_bounds(n, w)
            assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [h - l for l, h in b]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert shard_bounds(12, 8) == [(0, 2), (2, 4), (4, 6), (6, 8), (8, 9), (9, 10), (10, 11), (11, 12)]


def test_pack_roundtrip_is_bit_exact_for_fp64_energies():
    e = torch.tensor([-1234567.123456789, 3.5e-9, 0.1], dtype=torch.float64)
    f = torch.randn(3, 5, 3)
    rec = pack_results(e, f, cap=4, n_atoms=5)
    assert rec.dtype == torch.float32 and rec.numel() == 4 * 2 + 4 * 15
    e2, f2 = unpack_results(rec, 3, 4, 5)
    assert torch.equal(e, e2) and torch.equal(f, f2)


def _toy_local(coords):
    c = torch.from_numpy(np.asarray(coords, dtype=np.float64))
    e = (c ** 2).sum((1, 2)) + 1e6
    f = (-2 * c).float()
    return e, f


def _worker(rank, world, port, n_img, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        coords = rng.normal(size=(n_img, 6, 3))
        ev = SpmdEvaluator(_toy_local, 6)
        e, f = ev.evaluate(coords)
        e_ref, f_ref = _toy_local(coords)
        ok = torch.equal(e, e_ref) and torch.equal(f, f_ref)
        q.put((rank, bool(ok), tuple(e.shape), tuple(f.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_img", [4, 5, 1])
def test_spmd_evaluator_gloo_world2(n_img):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_img, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, es, fs in res:
        assert ok and es == (n_img,) and fs == (n_img, 6, 3)


def _hess_worker(rank, world, port, frozen, partial, q):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from helpers import SpringBackend
    from pdb2reaction_b200 import uma_pysis
    from pdb2reaction_b200.sharding import sharded_fd_hessian
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        elem = ["C", "H", "H", "O", "N"]
        x = np.array([[0, 0, 0], [1.1, 0, 0], [0, 1.0, 0.2], [0.3, -0.9, 0.8], [-1.0, 0.2, 0.5]]) / 0.529177210903
        kw = dict(freeze_atoms=frozen, return_partial_hessian=partial, out_hess_torch=True)
        full = uma_pysis(_backend=SpringBackend(), **kw).get_hessian(elem, x.reshape(-1))
        shard = sharded_fd_hessian(uma_pysis(_backend=SpringBackend(), **kw), elem, x.reshape(-1))
        ok = (torch.equal(full["hessian"], shard["hessian"]) and full["energy"] == shard["energy"]
              and np.array_equal(full["forces"], shard["forces"]))
        q.put((rank, bool(ok), tuple(shard["hessian"].shape)))
    finally:
        dist.destroy_process_group()


def _ahess_worker(rank, world, port, frozen, partial, q):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from helpers import SpringBackendAnalytic as SpringBackend
    from pdb2reaction_b200 import uma_pysis
    from pdb2reaction_b200.sharding import sharded_analytic_hessian
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        elem = ["C", "H", "H", "O", "N"]
        x = np.array([[0, 0, 0], [1.1, 0, 0], [0, 1.0, 0.2], [0.3, -0.9, 0.8], [-1.0, 0.2, 0.5]]) / 0.529177210903
        kw = dict(freeze_atoms=frozen, return_partial_hessian=partial, out_hess_torch=True, hessian_calc_mode="Analytical")
        full = uma_pysis(_backend=SpringBackend(), **kw).get_hessian(elem, x.reshape(-1))
        shard = sharded_analytic_hessian(uma_pysis(_backend=SpringBackend(), **kw), elem, x.reshape(-1))
        ok = (torch.equal(full["hessian"], shard["hessian"]) and full["energy"] == shard["energy"]
              and np.array_equal(full["forces"], shard["forces"]))
        q.put((rank, bool(ok), tuple(shard["hessian"].shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("frozen,partial,shape", [([], False, (15, 15)), ([1, 3], False, (15, 15)), ([0, 1, 2, 4], True, (3, 3))])
def test_sharded_analytic_hessian_gloo_world2(frozen, partial, shape):
    """Analytic mode (BASELINE configs[2]): column blocks sharded over two ranks + one all_gather == the single-process
    ``get_hessian(hessian_calc_mode="Analytical")``, bit for bit, including frozen columns, the active-block reduction
    and a rank with fewer columns than the other (3 active columns over 2 ranks)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ahess_worker, args=(r, 2, port, frozen, partial, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shp in res:
        assert ok and shp == shape


@pytest.mark.parametrize("frozen,partial,shape", [([], False, (15, 15)), ([1, 3], False, (15, 15)), ([0, 1, 2, 4], True, (3, 3))])
def test_sharded_fd_hessian_gloo_world2(frozen, partial, shape):
    """Column blocks sharded over two ranks + one all_gather == the single-process Hessian, bit for bit
    (also with fewer active columns than ranks x 2 and with the active-block reduction)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_hess_worker, args=(r, 2, port, frozen, partial, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shp in res:
        assert ok and shp == shape


def _string_worker(rank, world, port, n_img, q):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from helpers import SpringBackend
    from pdb2reaction_b200 import uma_pysis
    from pdb2reaction_b200.sharding import sharded_get_forces_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        elem = ["C", "H", "H", "O", "N"]
        rng = np.random.default_rng(3)
        x = np.array([[0, 0, 0], [1.1, 0, 0], [0, 1.0, 0.2], [0.3, -0.9, 0.8], [-1.0, 0.2, 0.5]]) / 0.529177210903
        coords = np.stack([x + 0.05 * rng.normal(size=x.shape) for _ in range(n_img)]).reshape(n_img, -1)
        mine = coords if rank == 0 else np.zeros_like(coords)          # only the optimizer's rank holds the geometry
        full = uma_pysis(_backend=SpringBackend(), freeze_atoms=[2]).get_forces_batch(elem, coords)
        be = SpringBackend()
        shard = sharded_get_forces_batch(uma_pysis(_backend=be, freeze_atoms=[2]), elem, mine)
        ok = np.array_equal(full["energy"], shard["energy"]) and np.array_equal(full["forces"], shard["forces"])
        q.put((rank, bool(ok), sum(b for b, _ in be.calls), tuple(shard["forces"].shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_img", [6, 5, 1])
def test_sharded_get_forces_batch_gloo_world2(n_img):
    """The strong-scaled string step (bench.py --gpus N): coordinates broadcast from the optimizer's rank, images
    sharded, one all_gather -> every rank holds the single-process result, bit for bit; each rank evaluated only its
    block."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_string_worker, args=(r, 2, port, n_img, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bounds = shard_bounds(n_img, 2)
    for rank, ok, n_eval, shp in res:
        assert ok and shp == (n_img, 15)
        assert n_eval == bounds[rank][1] - bounds[rank][0]
