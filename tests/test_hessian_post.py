"""Hessian post-processing: the numpy oracle's properties on CPU, CUDA parity on the GPU."""
import numpy as np
import pytest
import torch

from oracle import hessian_ref as R
from pdb2reaction_b200 import hessian_post as hp
from pdb2reaction_b200 import synth
from pdb2reaction_b200.arch import atomic_numbers


def _spring_hessian(x, k=0.7, r0=1.4, seed=0):
    """Exact Hessian (Hartree/Bohr^2-like numbers) of a random-spring network: translation AND rotation invariant
    energy, so its mass-weighted Hessian has six zero modes at a stationary point only; we make it stationary by
    choosing r0_ij = |x_i - x_j|."""
    n = x.shape[0]
    rng = np.random.default_rng(seed)
    h = np.zeros((3 * n, 3 * n))
    for i in range(n):
        for j in range(i + 1, n):
            if rng.random() < 0.6 or j == i + 1:
                d = x[i] - x[j]
                u = d / np.linalg.norm(d)
                kij = k * (0.5 + rng.random())
                blk = kij * np.outer(u, u)            # at r = r0 the Hessian of 1/2 k (r - r0)^2 is k u u^T
                h[3 * i:3 * i + 3, 3 * i:3 * i + 3] += blk
                h[3 * j:3 * j + 3, 3 * j:3 * j + 3] += blk
                h[3 * i:3 * i + 3, 3 * j:3 * j + 3] -= blk
                h[3 * j:3 * j + 3, 3 * i:3 * i + 3] -= blk
    return h


def _system(n=14, seed=3):
    elem, pos = synth.make_cluster(n, seed)
    z = atomic_numbers(elem)
    x = pos / 0.529177210903
    return z, x, hp.masses_amu_for(z)


# --------------------------------------------------------------------------- oracle properties (CPU)
def test_constants_match_ase_units():
    # ase.units (CODATA 2014): invcm = 1.239841973964072e-4 eV; the frequency factor in cm^-1 per sqrt(Ha/Bohr^2/amu)
    assert hp.INVCM_EV == pytest.approx(1.239841973964072e-4, rel=1e-12)
    assert hp.FREQ_EV_FACTOR / hp.INVCM_EV == pytest.approx(5140.487, rel=2e-6)     # textbook sqrt(Eh/(a0^2 amu)) in cm^-1
    assert hp.AMU2AU == pytest.approx(1822.888486, rel=1e-9)
    assert hp.masses_amu_for([1, 6, 7, 8, 16]).tolist() == [1.008, 12.011, 14.007, 15.999, 32.06]
    with pytest.raises(ValueError):
        hp.masses_amu_for([92])


def test_oracle_tr_basis_is_orthonormal_and_spans_rigid_motions():
    z, x, m = _system()
    q, r = R.tr_orthonormal_basis(x, m * hp.AMU2AU)
    assert r == 6 and np.allclose(q.T @ q, np.eye(6), atol=1e-12)
    # a linear molecule has only 5 rigid-body modes
    xl = np.zeros((4, 3)); xl[:, 2] = np.arange(4) * 2.0
    assert R.tr_orthonormal_basis(xl, np.array([1.0, 12.0, 12.0, 16.0]))[1] == 5


def test_oracle_projection_equals_dense_formula_and_is_idempotent():
    z, x, m = _system()
    rng = np.random.default_rng(0)
    h = rng.normal(size=(3 * len(z),) * 2); h = 0.5 * (h + h.T)
    mau = m * hp.AMU2AU
    out = R.mw_projected_hessian(h, x, mau)
    s = np.sqrt(1.0 / np.repeat(m, 3))
    q, _ = R.tr_orthonormal_basis(x, mau)
    p = np.eye(h.shape[0]) - q @ q.T
    assert np.allclose(out, p @ (h * s[:, None] * s[None, :]) @ p, atol=1e-10)
    assert np.allclose(out, out.T) and np.allclose(out @ q, 0.0, atol=1e-10)


def test_oracle_frequencies_of_a_stationary_spring_network():
    z, x, m = _system()
    h = _spring_hessian(x)
    freqs, modes = R.frequencies_cm_and_modes(h, m, x, tol=1e-9)
    assert len(freqs) == 3 * len(z) - 6 and (freqs > 0).all()            # 6 rigid modes removed, all real
    assert np.allclose(modes @ modes.T, np.eye(len(freqs)), atol=1e-9)
    # PHVA: full Hessian and its active block give the same spectrum
    frz = [0, 5, 9]
    f_full, m_full = R.frequencies_cm_and_modes(h, m, x, tol=1e-9, freeze_idx=frz)
    act = np.array([i for i in range(len(z)) if i not in frz])
    dof = (3 * act[:, None] + np.arange(3)).reshape(-1)
    f_blk, m_blk = R.frequencies_cm_and_modes(h[np.ix_(dof, dof)], m, x, tol=1e-9, freeze_idx=frz)
    assert np.allclose(f_full, f_blk, rtol=1e-9)
    assert np.allclose(m_full.reshape(len(f_full), -1, 3)[:, frz], 0.0)
    assert np.allclose(np.abs(np.sum(m_full * m_blk, axis=1)), 1.0, atol=1e-6)


def test_diatomic_frequency_known_answer():
    # harmonic diatomic: nu = sqrt(k / mu); H2-like with k = 0.37 Ha/Bohr^2 -> sqrt(k/mu[amu]) * 5140.487 cm^-1
    k, m = 0.37, np.array([1.008, 1.008])
    x = np.array([[0, 0, 0.0], [0, 0, 1.4]])
    u = np.array([0, 0, 1.0])
    blk = k * np.outer(u, u)
    h = np.block([[blk, -blk], [-blk, blk]])
    freqs, _ = R.frequencies_cm_and_modes(h, m, x)
    assert len(freqs) == 1
    assert freqs[0] == pytest.approx(np.sqrt(k / (1.008 / 2)) * hp.FREQ_EV_FACTOR / hp.INVCM_EV, rel=1e-10)


def test_oracle_fd_columns():
    rng = np.random.default_rng(1)
    f = rng.normal(size=(6, 9)).astype(np.float32)
    out = R.fd_columns(f, [4, 0, 7], 9, 1e-3)
    assert np.allclose(out[:, 0], -(f[2].astype(float) - f[3]) / 2e-3) and np.all(out[:, [1, 2, 3, 5, 6, 8]] == 0)


def test_no_cpu_path():
    with pytest.raises(RuntimeError, match="CUDA"):
        hp.mw_projected_hessian(torch.zeros(6, 6, dtype=torch.float64), torch.zeros(2, 3), torch.ones(2))
    with pytest.raises(RuntimeError, match="CUDA"):
        hp.frequencies_cm_and_modes(torch.zeros(6, 6, dtype=torch.float64), [1, 1], np.zeros((2, 3)))


# --------------------------------------------------------------------------- CUDA parity
@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 14, 97, 500])
def test_gpu_mw_projection_matches_oracle(built_lib, n):
    z, x, m = _system(n, seed=n)
    rng = np.random.default_rng(n)
    h = rng.normal(size=(3 * n, 3 * n)); h = 0.5 * (h + h.T)
    ref = R.mw_projected_hessian(h, x, m * hp.AMU2AU)
    ht = torch.tensor(h, device="cuda")
    out = hp.mw_projected_hessian(ht, torch.tensor(x, device="cuda"), torch.tensor(m * hp.AMU2AU, device="cuda"))
    assert out.data_ptr() == ht.data_ptr()                                  # in place, as the reference
    scale = np.abs(ref).max()
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-12 * scale
    assert torch.equal(out, out.T)                                          # exactly symmetric
    # deterministic: bitwise identical on a second run
    out2 = hp.mw_projected_hessian(torch.tensor(h, device="cuda"), torch.tensor(x, device="cuda"),
                                   torch.tensor(m * hp.AMU2AU, device="cuda"))
    assert torch.equal(out, out2)


@pytest.mark.gpu
def test_gpu_frequencies_match_oracle_all_branches(built_lib):
    z, x, m = _system(40, seed=11)
    h = _spring_hessian(x)
    for frz, block in ((None, False), ([0, 7, 22], False), ([0, 7, 22], True)):
        hh = h
        if block:
            act = np.array([i for i in range(len(z)) if i not in frz])
            dof = (3 * act[:, None] + np.arange(3)).reshape(-1)
            hh = h[np.ix_(dof, dof)]
        f_ref, m_ref = R.frequencies_cm_and_modes(hh, m, x, freeze_idx=frz)
        f_gpu, m_gpu = hp.frequencies_cm_and_modes(torch.tensor(hh, device="cuda"), z, x, freeze_idx=frz)
        assert f_gpu.shape == f_ref.shape and np.allclose(f_gpu, f_ref, rtol=1e-8, atol=1e-6)
        assert m_gpu.is_cuda and tuple(m_gpu.shape) == m_ref.shape
        # modes agree up to sign where the spectrum is non-degenerate
        ov = np.abs(np.sum(m_gpu.cpu().numpy() * m_ref, axis=1))
        gaps = np.minimum(np.diff(f_ref, prepend=-1e9), np.diff(f_ref, append=1e9))
        assert (ov[gaps > 1e-3] > 1 - 1e-6).all()
    cart = hp.mw_mode_to_cart(m_gpu[0], torch.tensor(m * hp.AMU2AU, device="cuda"))
    assert np.allclose(cart, R.mw_mode_to_cart(m_gpu[0].cpu().numpy(), m * hp.AMU2AU), atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_gpu_fd_columns_bit_exact(built_lib, dtype):
    rng = np.random.default_rng(5)
    dof, ks = 33, [4, 0, 31, 17]
    f = rng.normal(size=(2 * len(ks), dof)).astype(np.float32)
    hm = torch.zeros(dof, dof, dtype=dtype, device="cuda")
    hp.fd_hessian_columns_(hm, torch.tensor(f, device="cuda"), torch.tensor(ks, dtype=torch.int32, device="cuda"), 1e-3)
    # the reference's formula evaluated by torch in the same dtype (uma_pysis.py:668-670)
    ft = torch.tensor(f, device="cuda").to(dtype)
    ref = torch.zeros(dof, dof, dtype=dtype, device="cuda")
    ref[:, torch.tensor(ks, device="cuda")] = (-(ft[0::2] - ft[1::2]) / (2.0 * 1e-3)).T
    assert torch.equal(hm, ref)
