"""ORACLE (test infrastructure only): numpy restatement of the reference's Hessian post-processing.

Follows ``pdb2reaction/freq.py``: ``_build_tr_basis`` (:120-141), ``_tr_orthonormal_basis`` (:144-157),
``_mw_projected_hessian`` (:159-205), ``_mass_weighted_hessian`` (:209-225), ``_frequencies_cm_and_modes``
(:228-366, incl. both PHVA branches) and ``_mw_mode_to_cart`` (:369-381); and the finite-difference column formula
of ``uma_pysis._build_fd_hessian_gpu`` (``pdb2reaction/uma_pysis.py:652-675``).  Parity pinned only by the
properties in tests/test_hessian_post.py (the reference itself cannot be imported here: ase / pysisyphus absent).
Only tests/ may import this module.
"""
import numpy as np

from pdb2reaction_b200.hessian_post import AMU2AU, FREQ_EV_FACTOR, INVCM_EV


def build_tr_basis(coords_bohr, masses_au):
    x = np.asarray(coords_bohr, dtype=np.float64).reshape(-1, 3)
    m = np.asarray(masses_au, dtype=np.float64)
    n = x.shape[0]
    ms = np.sqrt(m).reshape(-1, 1)
    com = (m.reshape(-1, 1) * x).sum(0) / m.sum()
    x = x - com
    eye = np.eye(3)
    cols = []
    for i in range(3):
        cols.append((np.tile(eye[i], (n, 1)) * ms).reshape(-1, 1))
    for i in range(3):
        cols.append((np.cross(x, np.broadcast_to(eye[i], x.shape)) * ms).reshape(-1, 1))
    return np.concatenate(cols, axis=1)


def tr_orthonormal_basis(coords_bohr, masses_au, rtol=1e-12):
    b = build_tr_basis(coords_bohr, masses_au)
    u, s, _ = np.linalg.svd(b, full_matrices=False)
    r = int((s > rtol * s.max()).sum())
    return u[:, :r], r


def _inv_sqrt_m3(masses_au):
    m_amu = np.asarray(masses_au, dtype=np.float64) / AMU2AU
    return np.sqrt(1.0 / np.repeat(m_amu, 3))


def _project(h, q):
    qt = q.T
    qth = qt @ h
    h = h - q @ qth
    h = h - qth.T @ qt
    h = h + (q @ (qth @ q)) @ qt
    return 0.5 * (h + h.T)


def mw_projected_hessian(h, coords_bohr, masses_au):
    s = _inv_sqrt_m3(masses_au)
    h = np.asarray(h, dtype=np.float64) * s[:, None] * s[None, :]
    q, _ = tr_orthonormal_basis(coords_bohr, masses_au)
    return _project(h, q)


def frequencies_cm_and_modes(h, masses_amu, coords_bohr, tol=1e-6, freeze_idx=None):
    h = np.asarray(h, dtype=np.float64).copy()
    masses_au = np.asarray(masses_amu, dtype=np.float64) * AMU2AU
    x = np.asarray(coords_bohr, dtype=np.float64).reshape(-1, 3)
    n = x.shape[0]
    if freeze_idx is not None and len(freeze_idx) > 0:
        frozen = set(int(i) for i in freeze_idx if 0 <= int(i) < n)
        active = [i for i in range(n) if i not in frozen]
        if not active:
            return np.zeros(0), np.zeros((0, 3 * n))
        mask = np.ones(3 * n, dtype=bool)
        for i in frozen:
            mask[3 * i:3 * i + 3] = False
        if h.shape[0] == 3 * len(active):
            s = _inv_sqrt_m3(masses_au[active])
            hm = h * s[:, None] * s[None, :]
        else:
            s = _inv_sqrt_m3(masses_au)
            hm = (h * s[:, None] * s[None, :])[mask][:, mask]
        q, _ = tr_orthonormal_basis(x[active], masses_au[active])
        w2, v = np.linalg.eigh(_project(hm, q))
        sel = np.abs(w2) > tol
        w2, v = w2[sel], v[:, sel]
        modes = np.zeros((v.shape[1], 3 * n))
        modes[:, mask] = v.T
    else:
        w2, v = np.linalg.eigh(mw_projected_hessian(h, x, masses_au))
        sel = np.abs(w2) > tol
        w2 = w2[sel]
        modes = v[:, sel].T
    hnu = FREQ_EV_FACTOR * np.sqrt(np.abs(w2))
    hnu = np.where(w2 < 0, -hnu, hnu)
    return hnu / INVCM_EV, modes


def mw_mode_to_cart(mode, masses_au):
    v = _inv_sqrt_m3(masses_au) * np.asarray(mode, dtype=np.float64)
    return v / np.linalg.norm(v)


def fd_columns(forces, ks, dof, h_step, dtype=np.float64):
    """forces [2K, dof] (pairs +h, -h) -> dense [dof, dof] with columns ks filled."""
    f = np.asarray(forces, dtype=np.float64 if dtype == np.float64 else np.float32).astype(np.float64)
    out = np.zeros((dof, dof), dtype=dtype)
    for q, k in enumerate(ks):
        out[:, k] = (-(f[2 * q] - f[2 * q + 1]) / (2.0 * h_step)).astype(dtype)
    return out
