"""Image / displacement sharding across GPUs.

Units (images of a string, FD displacements) are independent evaluations of the same model,
so the data path has NO collective: each rank evaluates a contiguous block of units with its
own replica of the merged weights.  The only exchange is one ``all_gather`` of the packed
result ``[E (fp64) | F (fp32)]`` per optimizer step, so that every rank (and the optimizer that
lives on rank 0) sees the whole string -- the new meaning of the reference's ``workers`` option
(``pdb2reaction/uma_pysis.py:52-63, 213-242``; there: Ray actors + graph parallelism over ONE
structure).

Hessian column blocks shard the same way (BASELINE.json configs[2]): ``sharded_analytic_hessian`` (dual-number columns,
the mode configs[2] names) and ``sharded_fd_hessian`` (the reference's default mode) give every rank a contiguous block
of the active columns and end in ONE all_gather of the blocks.

Three drivers use this module:
* ``sharded_analytic_hessian`` / ``sharded_fd_hessian``: one process per GPU under ``torchrun`` (bench.py ``--hessian
  [--hessian-mode analytic|fd] --gpus N``);
* ``CudaBackend`` (calculator.py): one process, one host thread per GPU, no collective at all;
* ``SpmdEvaluator``: one process per GPU under ``torchrun`` (bench.py ``--gpus N``), NCCL
  all-gather over NVLink (gloo on CPU in the tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_units: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous block partition; the first ``n_units % world`` ranks get one extra unit."""
    base, extra = divmod(int(n_units), int(world))
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def pack_results(e: torch.Tensor, f: torch.Tensor, cap: int, n_atoms: int) -> torch.Tensor:
    """[b] fp64 energies + [b, n, 3] fp32 forces -> flat fp32 record padded to ``cap`` images:
    [cap x 2 floats holding the fp64 bits | cap x n x 3]."""
    b = e.shape[0]
    n3 = 3 * int(n_atoms)          # fixed record size on every rank, also for an empty shard
    rec = torch.zeros(cap * 2 + cap * n3, dtype=torch.float32, device=f.device)
    rec[: 2 * b] = e.to(torch.float64).contiguous().view(torch.float32)
    rec[2 * cap: 2 * cap + b * n3] = f.reshape(-1)
    return rec


def unpack_results(rec: torch.Tensor, b: int, cap: int, n_atoms: int):
    e = rec[: 2 * b].clone().view(torch.float64)          # clone: a row of an all_gather buffer may sit at an odd fp32 offset
    f = rec[2 * cap: 2 * cap + b * n_atoms * 3].reshape(b, n_atoms, 3)
    return e, f


class SpmdEvaluator:
    """Evaluate a full batch of images across the ranks of a process group.

    ``evaluate_local(coords [b,N,3] float64 numpy) -> (E torch fp64 [b], F torch fp32 [b,N,3])``
    on this rank's device.
    """

    def __init__(self, evaluate_local: Callable, n_atoms: int, group: Optional[dist.ProcessGroup] = None):
        self.evaluate_local = evaluate_local
        self.n_atoms = int(n_atoms)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def evaluate(self, coords_all: np.ndarray, gather: bool = True):
        """coords_all [B,N,3] (identical on every rank) -> (E [B], F [B,N,3]) on every rank."""
        b_tot = coords_all.shape[0]
        bounds = shard_bounds(b_tot, self.world)
        lo, hi = bounds[self.rank]
        e, f = self.evaluate_local(coords_all[lo:hi])
        if self.world == 1 or not gather:
            return e, f
        cap = max(h - l for l, h in bounds)
        rec = pack_results(e, f, cap, self.n_atoms)
        out = [torch.empty_like(rec) for _ in range(self.world)]
        dist.all_gather(out, rec, group=self.group)          # the ONE collective per step
        es, fs = [], []
        for r, (l, h) in enumerate(bounds):
            er, fr = unpack_results(out[r], h - l, cap, self.n_atoms)
            es.append(er)
            fs.append(fr)
        return torch.cat(es), torch.cat(fs)



def sharded_get_forces_batch(calc, elem, coords_bohr, group: Optional[dist.ProcessGroup] = None, src: int = 0):
    """``calc.get_forces_batch(elem, coords_bohr)`` with the IMAGES of the string sharded over the ranks of ``group``
    (one process per GPU under torchrun; the new meaning of ``workers``, reference ``uma_pysis.py:205-242``).

    The optimizer lives on rank ``src``: its host coordinates [B, 3N] (Bohr) are broadcast, rank r evaluates the
    contiguous block ``shard_bounds(B, world)[r]`` on its own device with its replica of the weights, ONE all_gather
    of the packed ``[E | F]`` records follows, and every rank returns the full host result
    ``{"energy": [B] Hartree, "forces": [B, 3N] Hartree/Bohr}`` -- identical bits to the single-process call."""
    from .calculator import EV2AU, F_EVAA_2_AU
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    calc._ensure_core(elem)
    if world == 1:
        return calc.get_forces_batch(elem, coords_bohr)
    core = calc._core
    n_atoms = len(core.elem)
    dev = core.device if dist.get_backend(group) == "nccl" else torch.device("cpu")
    c = torch.as_tensor(np.ascontiguousarray(np.asarray(coords_bohr, dtype=np.float64).reshape(-1, 3 * n_atoms)))
    c = c.to(dev)
    dist.broadcast(c, src=src, group=group)                  # coordinates of the step: B x 3N fp64 (C4: 1.15 MB)
    b_tot = c.shape[0]
    bounds = shard_bounds(b_tot, world)
    lo, hi = bounds[rank]
    cap = max(h - l for l, h in bounds)
    backend = core.backend
    if hi > lo and dev.type == "cuda" and hasattr(backend, "evaluate_device"):
        e, f = backend.evaluate_device(calc._coords_ang(c[lo:hi].cpu().numpy(), batch=True))     # results stay on the GPU
    elif hi > lo:
        r = core.compute_batch(calc._coords_ang(c[lo:hi].cpu().numpy(), batch=True), forces=True)
        e, f = torch.from_numpy(r["energy"]).to(dev), torch.from_numpy(r["forces"]).to(dev)
    else:
        e = torch.zeros(0, dtype=torch.float64, device=dev)
        f = torch.zeros((0, n_atoms, 3), dtype=torch.float32, device=dev)
    rec = pack_results(e, f, cap, n_atoms)
    out = torch.empty(world * rec.numel(), dtype=rec.dtype, device=dev)
    dist.all_gather_into_tensor(out, rec, group=group)       # the ONE collective of the step
    out = out.cpu().view(world, -1)
    es, fs = [], []
    for r_, (l, h) in enumerate(bounds):
        er, fr = unpack_results(out[r_], h - l, cap, n_atoms)
        es.append(er)
        fs.append(fr)
    e_all = torch.cat(es).numpy()
    f_all = calc._zero_frozen_forces_ev(torch.cat(fs).numpy())
    return {"energy": e_all * EV2AU, "forces": (np.asarray(f_all, dtype=np.float64) * F_EVAA_2_AU).reshape(b_tot, -1)}


def sharded_fd_hessian(calc, elem, coords_bohr, group: Optional[dist.ProcessGroup] = None):
    """``calc.get_hessian(elem, coords)`` in FiniteDifference mode with the ACTIVE COLUMNS sharded over the ranks of
    ``group`` (one process per GPU): rank r evaluates the +-h displacements of its contiguous block of active degrees
    of freedom on its own device, then one all_gather of the column blocks (``[cols_r, 3N]`` each; C3: 9 MB fp32 /
    18 MB fp64 in total) gives every rank the full matrix, which goes through the calculator's own symmetrise / unit /
    dtype formatting (reference ``uma_pysis.py:515-551``).  Same result bits as the single-process path."""
    from .calculator import FD_STEP_ANG
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    calc._ensure_core(elem)
    coord_ang = calc._coords_ang(coords_bohr)
    core = calc._core
    n_atoms = coord_ang.shape[0]
    dof = 3 * n_atoms
    _, active_dof, _ = calc._active_and_frozen_dof_idx(n_atoms)
    res0 = core.compute(coord_ang, forces=True)
    hdt = torch.float64 if calc.hessian_double else torch.float32
    hmat = torch.zeros((dof, dof), device=core.device, dtype=hdt)
    bounds = shard_bounds(len(active_dof), world)
    lo, hi = bounds[rank]
    calc._fd_columns_into(hmat, coord_ang, active_dof[lo:hi], FD_STEP_ANG)
    if world > 1:
        cap = max(h - l for l, h in bounds)
        mine = torch.zeros((cap, dof), device=hmat.device, dtype=hdt)
        if hi > lo:
            idx = torch.as_tensor(active_dof[lo:hi], device=hmat.device, dtype=torch.long)
            mine[: hi - lo] = hmat.index_select(1, idx).T
        blocks = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(blocks, mine, group=group)          # the ONE collective of a Hessian
        for r, (l, h) in enumerate(bounds):
            if h > l and r != rank:
                idx = torch.as_tensor(active_dof[l:h], device=hmat.device, dtype=torch.long)
                hmat[:, idx] = blocks[r][: h - l].T
    f_ev = calc._zero_frozen_forces_ev(res0["forces"])
    return {"energy": calc._au_energy(res0["energy"]), "forces": calc._au_forces(f_ev),
            "hessian": calc._au_hessian(calc._finish_fd_hessian(hmat, n_atoms))}


def sharded_analytic_hessian(calc, elem, coords_bohr, group: Optional[dist.ProcessGroup] = None):
    """``calc.get_hessian(elem, coords)`` in Analytical mode (reference ``uma_pysis.py:394-415``; BASELINE.json
    configs[2]: "full analytic Hessian, column blocks sharded across GPUs") with the ACTIVE COLUMNS sharded over the
    ranks of ``group`` (one process per GPU): rank r runs the dual-number forward + backward passes of its contiguous
    block of active degrees of freedom on its own device (``backend.hessian_columns``), ONE all_gather of the column
    blocks (``[cols_r, 3N]`` fp32 each) gives every rank all columns, and the calculator's own assembly / trim /
    symmetrise / unit / dtype formatting follows (``_assemble_analytic_hessian``, ``_au_hessian``; reference
    ``:515-551, :569-592``).  Same result bits as the single-process ``get_hessian``.  Unlike the reference
    (``:737``), more than one worker does not force the finite-difference mode."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    calc._ensure_core(elem)
    core = calc._core
    if not hasattr(core.backend, "hessian_columns"):
        raise RuntimeError("sharded_analytic_hessian: this backend has no analytic Hessian columns")
    coord_ang = calc._coords_ang(coords_bohr)
    n_atoms = coord_ang.shape[0]
    dof = 3 * n_atoms
    _, active_dof, _ = calc._active_and_frozen_dof_idx(n_atoms)
    res0 = core.compute(coord_ang, forces=True)
    bounds = shard_bounds(len(active_dof), world)
    lo, hi = bounds[rank]
    dev = core.device if (world == 1 or dist.get_backend(group) == "nccl") else torch.device("cpu")
    if hi > lo:
        mine = torch.from_numpy(np.ascontiguousarray(core.backend.hessian_columns(coord_ang, active_dof[lo:hi]),
                                                     dtype=np.float32)).to(dev)
    else:
        mine = torch.zeros((0, dof), dtype=torch.float32, device=dev)
    if world > 1:
        cap = max(h - l for l, h in bounds)
        rec = torch.zeros((cap, dof), dtype=torch.float32, device=dev)
        rec[: hi - lo] = mine
        out = torch.empty((world * cap, dof), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(out, rec, group=group)              # the ONE collective of a Hessian
        cols = torch.cat([out[r * cap: r * cap + (h - l)] for r, (l, h) in enumerate(bounds)])
    else:
        cols = mine
    res = calc._assemble_analytic_hessian(res0, cols.to(core.device), n_atoms)
    f_ev = calc._zero_frozen_forces_ev(res["forces"])
    return {"energy": calc._au_energy(res["energy"]), "forces": calc._au_forces(f_ev),
            "hessian": calc._au_hessian(res["hessian"])}
