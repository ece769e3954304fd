"""Device-resident Hessian post-processing for ``freq`` / ``tsopt``.

Mirrors the helper functions of the reference's ``pdb2reaction/freq.py`` -- same names without the underscore,
same arguments, units and return conventions -- so that the vibrational analysis can consume the Hessian the
calculator leaves on the GPU without a host round trip:

* ``mw_projected_hessian``      freq.py:159-205   mass-weight + translation/rotation projection + symmetrise
* ``mass_weighted_hessian``     freq.py:209-225
* ``frequencies_cm_and_modes``  freq.py:228-366   full and PHVA (frozen atoms; full or active-block Hessian) branches
* ``mw_mode_to_cart``           freq.py:369-381

The dense chain of mul_/addmm_/transpose passes is ONE call into the CUDA library
(``umab_hessian_mw_project``: two passes over H, in place, fp64, deterministic); the rank-<=6 TR basis (an SVD of
a [3N,6] matrix) and the symmetric eigendecomposition stay on the device through torch (cuSOLVER).  There is no CPU
path: CPU tensors raise.  Constants: ``ase.units`` (CODATA 2014) and ``ase.data.atomic_masses`` values are
restated because ASE is not a dependency of this backend.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .shims import AU2EV, BOHR2ANG

# pysisyphus.constants.AMU2AU (atomic mass unit in electron masses, CODATA 2018 via scipy)
try:  # pragma: no cover
    from pysisyphus.constants import AMU2AU  # type: ignore
except Exception:
    import scipy.constants as _spc
    AMU2AU = 1.0 / _spc.value("electron mass in u")

# ase.units (CODATA 2014, ASE's default table): _hbar, _e, _amu, _c, _hplanck
_HBAR, _E, _AMU, _C, _HPLANCK = 1.054571800e-34, 1.6021766208e-19, 1.660539040e-27, 299792458.0, 6.626070040e-34
# freq.py:358: s_new = units._hbar * 1e10 / sqrt(units._e * units._amu) * sqrt(AU2EV) / BOHR2ANG   (eV per sqrt(Ha/(Bohr^2 amu)))
FREQ_EV_FACTOR = _HBAR * 1e10 / np.sqrt(_E * _AMU) * np.sqrt(AU2EV) / BOHR2ANG
INVCM_EV = 100.0 * _C * _HPLANCK / _E            # ase.units.invcm

# ase.data.atomic_masses (IUPAC 2016 abridged), index = Z; elements beyond the table need explicit ``masses_amu``
ATOMIC_MASSES = np.array([
    1.0, 1.008, 4.002602, 6.94, 9.0121831, 10.81, 12.011, 14.007, 15.999, 18.998403163, 20.1797,
    22.98976928, 24.305, 26.9815385, 28.085, 30.973761998, 32.06, 35.45, 39.948, 39.0983, 40.078,
    44.955908, 47.867, 50.9415, 51.9961, 54.938044, 55.845, 58.933194, 58.6934, 63.546, 65.38,
    69.723, 72.630, 74.921595, 78.971, 79.904, 83.798, 85.4678, 87.62, 88.90584, 91.224,
    92.90637, 95.95, 97.90721, 101.07, 102.90550, 106.42, 107.8682, 112.414, 114.818, 118.710,
    121.760, 127.60, 126.90447, 131.293,
])


def masses_amu_for(atomic_numbers: Sequence[int]) -> np.ndarray:
    z = np.asarray(list(atomic_numbers), dtype=int)
    if z.size and (z.min() < 1 or z.max() >= len(ATOMIC_MASSES)):
        raise ValueError("atomic number outside the built-in mass table (Z = 1..54): pass masses_amu explicitly")
    return ATOMIC_MASSES[z]


def _require_cuda(h: torch.Tensor):
    if not (isinstance(h, torch.Tensor) and h.is_cuda):
        raise RuntimeError("hessian_post works on CUDA tensors only (pdb2reaction_b200 has no CPU path)")


# ------------------------------------------------------------------ TR basis (freq.py:120-157), on the device
def build_tr_basis(coords_bohr_t: torch.Tensor, masses_au_t: torch.Tensor) -> torch.Tensor:
    """Mass-weighted translation/rotation vectors (Tx, Ty, Tz, Rx, Ry, Rz) as columns, [3N, 6]."""
    x = coords_bohr_t.reshape(-1, 3)
    dtype, device = x.dtype, x.device
    m = masses_au_t.to(dtype=dtype, device=device)
    ms = torch.sqrt(m).reshape(-1, 1)
    com = (m.reshape(-1, 1) * x).sum(0) / m.sum()
    x = x - com
    eye = torch.eye(3, dtype=dtype, device=device)
    n = x.shape[0]
    cols = [(eye[i].repeat(n, 1) * ms).reshape(-1, 1) for i in range(3)]
    cols += [(torch.cross(x, eye[i].expand_as(x), dim=1) * ms).reshape(-1, 1) for i in range(3)]
    return torch.cat(cols, dim=1)


def tr_orthonormal_basis(coords_bohr_t: torch.Tensor, masses_au_t: torch.Tensor, rtol: float = 1e-12) -> Tuple[torch.Tensor, int]:
    b = build_tr_basis(coords_bohr_t, masses_au_t)
    u, s, _ = torch.linalg.svd(b, full_matrices=False)
    r = int((s > rtol * s.max()).sum().item())
    return u[:, :r].contiguous(), r


def _inv_sqrt_m3(masses_au_t: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    m_amu = (masses_au_t / AMU2AU).to(dtype=torch.float64, device=like.device)
    return torch.sqrt(1.0 / torch.repeat_interleave(m_amu, 3)).contiguous()


def _mw_project_inplace(h: torch.Tensor, inv_sqrt_m: torch.Tensor, q: Optional[torch.Tensor]) -> torch.Tensor:
    """H <- sym(P (S H S) P) through the CUDA library, in place on a contiguous fp64 CUDA matrix."""
    from .engine import load_library, _check
    _require_cuda(h)
    assert h.dtype == torch.float64 and h.dim() == 2 and h.shape[0] == h.shape[1] and h.is_contiguous()
    lib = load_library()
    n = h.shape[0]
    r = 0 if q is None else int(q.shape[1])
    ws = None
    ws_n = 0
    if r > 0:
        q = q.to(dtype=torch.float64, device=h.device).contiguous()
        ws_n = int(lib.umab_hessian_mw_workspace(n, r))
        ws = torch.empty(ws_n, dtype=torch.float64, device=h.device)
    with torch.cuda.device(h.device):
        st = ctypes.c_void_p(torch.cuda.current_stream(h.device).cuda_stream)
        _check(lib, lib.umab_hessian_mw_project(h.data_ptr(), n, inv_sqrt_m.data_ptr(), q.data_ptr() if r else None, r,
                                                ws.data_ptr() if r else None, ws_n, st))
    return h


def mw_projected_hessian(H: torch.Tensor, coords_bohr_t: torch.Tensor, masses_au_t: torch.Tensor) -> torch.Tensor:
    """freq._mw_projected_hessian: updates H IN PLACE when it already is a contiguous fp64 tensor, returns it."""
    _require_cuda(H)
    if H.dtype != torch.float64:
        H = H.to(dtype=torch.float64)
    H = H if H.is_contiguous() else H.contiguous()
    with torch.no_grad():
        q, _ = tr_orthonormal_basis(coords_bohr_t.to(dtype=torch.float64, device=H.device),
                                    masses_au_t.to(dtype=torch.float64, device=H.device))
        return _mw_project_inplace(H, _inv_sqrt_m3(masses_au_t, H), q)


def mass_weighted_hessian(H: torch.Tensor, masses_au_t: torch.Tensor) -> torch.Tensor:
    """freq._mass_weighted_hessian: Hmw = M^-1/2 H M^-1/2, in place, no projection, no symmetrisation."""
    _require_cuda(H)
    with torch.no_grad():
        s = _inv_sqrt_m3(masses_au_t, H).to(dtype=H.dtype)
        H.mul_(s.view(-1, 1))
        H.mul_(s.view(1, -1))
        return H


def frequencies_cm_and_modes(H: torch.Tensor, atomic_numbers: List[int], coords_bohr: np.ndarray,
                             device: Optional[torch.device] = None, tol: float = 1e-6,
                             freeze_idx: Optional[List[int]] = None,
                             masses_amu: Optional[Sequence[float]] = None) -> Tuple[np.ndarray, torch.Tensor]:
    """freq._frequencies_cm_and_modes: frequencies in cm^-1 (negative = imaginary) and mass-weighted modes
    [nmode, 3N] on the device.  ``H`` is consumed (mass-weighted / projected in place when fp64)."""
    _require_cuda(H)
    device = H.device if device is None else torch.device(device)
    with torch.no_grad():
        if H.dtype != torch.float64:
            H = H.to(dtype=torch.float64)
        H = H.to(device)
        H = H if H.is_contiguous() else H.contiguous()
        m_amu = np.asarray(masses_amu, dtype=float) if masses_amu is not None else masses_amu_for(atomic_numbers)
        n = len(m_amu)
        masses_au_t = torch.as_tensor(m_amu * AMU2AU, dtype=torch.float64, device=device)
        coords_t = torch.as_tensor(np.asarray(coords_bohr, dtype=float).reshape(-1, 3), dtype=torch.float64, device=device)
        if freeze_idx is not None and len(freeze_idx) > 0:
            frozen = set(int(i) for i in freeze_idx if 0 <= int(i) < n)
            active = [i for i in range(n) if i not in frozen]
            if not active:
                return np.zeros((0,), dtype=float), torch.zeros((0, 3 * n), dtype=H.dtype, device=device)
            act_t = torch.as_tensor(active, dtype=torch.long, device=device)
            mask = torch.ones(3 * n, dtype=torch.bool, device=device)
            mask.view(n, 3)[torch.as_tensor(sorted(frozen), dtype=torch.long, device=device)] = False
            q, _ = tr_orthonormal_basis(coords_t[act_t], masses_au_t[act_t])
            if H.shape[0] == 3 * len(active):                         # active-block Hessian (freq.py:290-315)
                hm = _mw_project_inplace(H, _inv_sqrt_m3(masses_au_t[act_t], H), q)
            else:                                                     # full Hessian (freq.py:317-345)
                s = _inv_sqrt_m3(masses_au_t, H)
                hm = H[mask][:, mask].contiguous()                    # S commutes with the sub-block selection
                hm = _mw_project_inplace(hm, s[mask].contiguous(), q)
            w2, v = torch.linalg.eigh(hm)
            sel = torch.abs(w2) > tol
            w2, v = w2[sel], v[:, sel]
            modes = torch.zeros((v.shape[1], 3 * n), dtype=v.dtype, device=device)
            modes[:, mask] = v.T
        else:
            hm = mw_projected_hessian(H, coords_t, masses_au_t)
            w2, v = torch.linalg.eigh(hm)
            sel = torch.abs(w2) > tol
            w2 = w2[sel]
            modes = v[:, sel].T
        hnu = FREQ_EV_FACTOR * torch.sqrt(torch.abs(w2))
        hnu = torch.where(w2 < 0, -hnu, hnu)
        return (hnu / INVCM_EV).detach().cpu().numpy(), modes


def mw_mode_to_cart(mode_mw_3N_t: torch.Tensor, masses_au_t: torch.Tensor) -> np.ndarray:
    """freq._mw_mode_to_cart: one mass-weighted eigenvector -> L2-normalised Cartesian displacement (numpy)."""
    with torch.no_grad():
        s = _inv_sqrt_m3(masses_au_t, mode_mw_3N_t).to(dtype=mode_mw_3N_t.dtype)
        v = s * mode_mw_3N_t
        return (v / torch.linalg.norm(v)).detach().cpu().numpy()


def fd_hessian_columns_(hmat: torch.Tensor, forces_pm: torch.Tensor, dof_idx: torch.Tensor, h_step: float) -> torch.Tensor:
    """hmat[:, k_q] = -(F[2q] - F[2q+1]) / (2 h) for every q, on the device (uma_pysis.py:652-675).
    forces_pm [2K, dof] fp32, dof_idx [K] int32, hmat [dof, dof] fp32 / fp64 contiguous; returns hmat."""
    from .engine import load_library, _check
    _require_cuda(hmat)
    assert forces_pm.is_cuda and forces_pm.dtype == torch.float32 and forces_pm.is_contiguous()
    assert dof_idx.is_cuda and dof_idx.dtype == torch.int32 and hmat.is_contiguous()
    assert hmat.dtype in (torch.float32, torch.float64) and forces_pm.shape[0] == 2 * dof_idx.numel()
    lib = load_library()
    dof = hmat.shape[0]
    with torch.cuda.device(hmat.device):
        st = ctypes.c_void_p(torch.cuda.current_stream(hmat.device).cuda_stream)
        _check(lib, lib.umab_hessian_fd_columns(forces_pm.data_ptr(), dof_idx.data_ptr(), int(dof_idx.numel()), dof,
                                                float(h_step), hmat.data_ptr(), hmat.shape[1],
                                                int(hmat.dtype == torch.float64), st))
    return hmat
