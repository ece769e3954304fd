"""Wigner-D matrices for l <= 2 in the e3nn real basis, the way fairchem builds them.

Restates (from the published algorithm; fairchem-core is not installable here, see
``oracle/__init__.py``) ``fairchem/core/models/uma/common/rotation.py``:
``init_edge_rot_euler_angles``, ``eulers_to_wigner``, ``wigner_D``, ``_z_rot_mat`` -- which the
reference reaches through ``predict_unit.predict`` (``pdb2reaction/uma_pysis.py:385``).

fairchem loads the constant ``J_l`` matrices from ``Jd.pt`` (not available offline); here they
are regenerated numerically as the real-SH representation of the pi rotation about (x+y)/sqrt2,
the involution K with K Ry(b) K = Rx(b), so that  Z(a) J Z(b) J Z(c) = D(Ry(a) Rx(b) Ry(c)).
Basis: l=1 -> (x, y, z) (m = -1, 0, 1; the m=0 axis is y);  l=2 -> (sqrt3 xz, sqrt3 xy,
y^2 - (x^2+z^2)/2, sqrt3 yz, sqrt3/2 (z^2 - x^2)).
"""
from __future__ import annotations

import math

import numpy as np
import torch

SQ3 = math.sqrt(3.0)


def l2_quadratic_forms() -> np.ndarray:
    """Q[m] (3x3 symmetric traceless) with Y_2m(v) = v^T Q[m] v, m = -2..2."""
    q = np.zeros((5, 3, 3))
    q[0, 0, 2] = q[0, 2, 0] = SQ3 / 2
    q[1, 0, 1] = q[1, 1, 0] = SQ3 / 2
    q[2] = np.diag([-0.5, 1.0, -0.5])
    q[3, 1, 2] = q[3, 2, 1] = SQ3 / 2
    q[4] = np.diag([-SQ3 / 2, 0.0, SQ3 / 2])
    return q


def real_sh(v: np.ndarray, l: int) -> np.ndarray:
    v = np.asarray(v, dtype=np.float64)
    if l == 0:
        return np.ones(v.shape[:-1] + (1,))
    if l == 1:
        return v.copy()
    q = l2_quadratic_forms()
    return np.einsum("...a,mab,...b->...m", v, q, v)


def d_from_rotation(rot: np.ndarray, l: int) -> np.ndarray:
    """D^l(R) with Y_l(R v) = D^l(R) Y_l(v) (float64, any batch shape)."""
    rot = np.asarray(rot, dtype=np.float64)
    if l == 0:
        return np.ones(rot.shape[:-2] + (1, 1))
    if l == 1:
        return rot.copy()
    q = l2_quadratic_forms()
    rq = np.einsum("...ca,mcd,...db->...mab", rot, q, rot)  # R^T Q_m R
    return (2.0 / 3.0) * np.einsum("...mab,nab->...mn", rq, q)


K_SWAP = np.array([[0.0, 1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, -1.0]])


def jd_matrices(lmax: int = 2):
    return [d_from_rotation(K_SWAP, l) for l in range(lmax + 1)]


def z_rot_mat(angle: torch.Tensor, l: int) -> torch.Tensor:
    """(2l+1)^2 rotation about the m=0 (y) axis: cos(k a) on the diagonal, sin(k a) on the
    anti-diagonal, k = l..-l."""
    m = angle.new_zeros(angle.shape + (2 * l + 1, 2 * l + 1))
    inds = torch.arange(0, 2 * l + 1, device=angle.device)
    rev = torch.arange(2 * l, -1, -1, device=angle.device)
    freq = torch.arange(l, -l - 1, -1, dtype=angle.dtype, device=angle.device)
    m[..., inds, rev] = torch.sin(freq * angle[..., None])
    m[..., inds, inds] = torch.cos(freq * angle[..., None])
    return m


def wigner_d(l: int, alpha, beta, gamma, jd) -> torch.Tensor:
    j = jd[l].to(dtype=alpha.dtype, device=alpha.device)
    return z_rot_mat(alpha, l) @ j @ z_rot_mat(beta, l) @ j @ z_rot_mat(gamma, l)


def edge_rot_euler_angles(edge_vec: torch.Tensor, gamma=None):
    """beta = acos(n_y), alpha = atan2(n_x, n_z), gamma random (or given); returns the
    intrinsic->extrinsic swapped triple (-gamma, -beta, -alpha)."""
    n = edge_vec / edge_vec.norm(dim=1, keepdim=True)
    x, y, z = n[:, 0], n[:, 1], n[:, 2]
    beta = torch.acos(y.clamp(-1.0, 1.0))
    alpha = torch.atan2(x, z)
    if gamma is None:
        gamma = torch.zeros_like(alpha)
    gamma = gamma.detach()
    return -gamma, -beta, -alpha


def edge_wigner(edge_vec: torch.Tensor, lmax: int = 2, gamma=None) -> torch.Tensor:
    """Block-diagonal D(l=0..lmax) [E, (lmax+1)^2, (lmax+1)^2] in l-primary order; it rotates
    the edge direction onto the m=0 axis."""
    jd = [torch.from_numpy(j) for j in jd_matrices(lmax)]
    a, b, c = edge_rot_euler_angles(edge_vec, gamma)
    size = (lmax + 1) ** 2
    w = edge_vec.new_zeros(edge_vec.shape[0], size, size)
    start = 0
    for l in range(lmax + 1):
        blk = wigner_d(l, a, b, c, jd)
        end = start + 2 * l + 1
        w[:, start:end, start:end] = blk
        start = end
    return w
