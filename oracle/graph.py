"""Radius graph of non-periodic images, the way fairchem's ``generate_graph`` builds it.

Restates the published ``fairchem/core/graph/compute.py::generate_graph`` /
``radius_graph_pbc`` / ``get_max_neighbors_mask`` for pbc = False (one cell image), which the
reference triggers on every call through ``data_list_collater([data], otf_graph=True)``
(``pdb2reaction/uma_pysis.py:313-322``).  Positions are float32 (``AtomicData.from_ase``).

Rule: ordered pair (source j -> target i), same image, kept iff  d2 <= r_c^2  and  d2 > 1e-4
with  d2 = (dx*dx + dy*dy) + dz*dz  evaluated in float32 in exactly that order; per target at
most ``max_neighbors`` nearest are kept, non-strict: everything with
d2 <= d2_sorted[max_neighbors] + 0.01 survives (``enforce_max_neighbors_strictly=False``).
Canonical output order: sorted by (target, source).
"""
from __future__ import annotations

import numpy as np


def radius_graph(pos: np.ndarray, natoms, cutoff: float = 6.0, max_neighbors: int = 300,
                 block: int = 2048):
    """pos [sum(natoms), 3] float32, natoms: atoms per image.

    Returns edge_index [2, E] int64 (row 0 = source j, row 1 = target i), sorted by
    (target, source); indices are global (into the concatenated ``pos``).
    """
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    natoms = [int(n) for n in np.atleast_1d(natoms)]
    assert sum(natoms) == pos.shape[0]
    rc2 = np.float32(cutoff) * np.float32(cutoff)
    tiny = np.float32(1e-4)
    srcs, tgts = [], []
    start = 0
    for n in natoms:
        p = pos[start:start + n]
        for t0 in range(0, n, block):
            t1 = min(n, t0 + block)
            d = p[None, :, :] - p[t0:t1, None, :]            # [T, n, 3]  pos_j - pos_i
            dx2 = d[..., 0] * d[..., 0]
            dy2 = d[..., 1] * d[..., 1]
            dz2 = d[..., 2] * d[..., 2]
            d2 = (dx2 + dy2) + dz2                           # float32, fixed order
            keep = (d2 <= rc2) & (d2 > tiny)
            deg = keep.sum(axis=1)
            if max_neighbors is not None and deg.max(initial=0) > max_neighbors:
                for r in np.nonzero(deg > max_neighbors)[0]:
                    ds = np.sort(d2[r][keep[r]])
                    eff = ds[max_neighbors] + np.float32(0.01)
                    keep[r] &= d2[r] <= eff
            ti, sj = np.nonzero(keep)                        # row-major: sorted (target, source)
            tgts.append(ti.astype(np.int64) + t0 + start)
            srcs.append(sj.astype(np.int64) + start)
        start += n
    if not srcs:
        return np.zeros((2, 0), dtype=np.int64)
    return np.stack([np.concatenate(srcs), np.concatenate(tgts)])


def canonical_sort(edge_index: np.ndarray) -> np.ndarray:
    order = np.lexsort((edge_index[0], edge_index[1]))
    return edge_index[:, order]
