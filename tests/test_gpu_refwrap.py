"""-m gpu: the CUDA-backed drop-in calculator against what the REFERENCE's own wrapper code returned
(fixtures written by tests/golden/make_refwrap_golden.py, which executes /root/reference/pdb2reaction/
uma_pysis.py / freq.py / opt.py under third-party stand-ins with a float64-oracle predictor; that tree does not
exist on the GPU box).  The wrapper logic is compared bit for bit on the CPU (tests/test_reference_wrapper.py);
here the model arithmetic is the CUDA path, so values agree within the north-star tolerances."""
import os

import numpy as np
import pytest
import torch

from pdb2reaction_b200 import EV2AU, F_EVAA_2_AU, H_EVAA_2_AU, uma_pysis
from pdb2reaction_b200 import hessian_post as hp

pytestmark = pytest.mark.gpu
REFWRAP = os.path.join(os.path.dirname(__file__), "golden", "refwrap")

HESS_CASES = {
    "fd": dict(),
    "fd_frozen_partial": dict(freeze_atoms=[0, 5], return_partial_hessian=True),
    "fd_frozen_f32_numpy": dict(freeze_atoms=[2], hessian_double=False, out_hess_torch=False),
    "an": dict(hessian_calc_mode="Analytical"),
    "an_frozen": dict(hessian_calc_mode="Analytical", freeze_atoms=[1, 6]),
    "an_frozen_partial_f32": dict(hessian_calc_mode="Analytical", freeze_atoms=[1, 6], return_partial_hessian=True,
                                  hessian_double=False),
}


@pytest.fixture(scope="module")
def calc_golden():
    return np.load(os.path.join(REFWRAP, "calc_n12.npz"))


def test_forces_match_the_reference_wrapper(built_lib, small_model, calc_golden):
    g = calc_golden
    elem = [str(e) for e in g["elem"]]
    r = uma_pysis(model="test-4x", freeze_atoms=[3, 4]).get_forces(elem, g["coords_bohr"])
    assert abs(r["energy"] - float(g["forces_frozen34.energy"])) < 1e-5 * 12 * EV2AU
    assert np.abs(r["forces"] - g["forces_frozen34.forces"]).max() < 1e-4 * F_EVAA_2_AU
    assert np.all(r["forces"].reshape(12, 3)[[3, 4]] == 0.0) and np.all(g["forces_frozen34.forces"].reshape(12, 3)[[3, 4]] == 0.0)


@pytest.mark.parametrize("name", list(HESS_CASES))
def test_hessian_matches_the_reference_wrapper(name, built_lib, small_model, calc_golden):
    g = calc_golden
    elem = [str(e) for e in g["elem"]]
    r = uma_pysis(model="test-4x", **HESS_CASES[name]).get_hessian(elem, g["coords_bohr"])
    h, ref = r["hessian"], g[name + ".hessian"]
    assert isinstance(h, torch.Tensor) == bool(g[name + ".is_torch"])
    if isinstance(h, torch.Tensor):
        assert h.is_cuda
        h = h.cpu().numpy()
    assert h.dtype == ref.dtype and h.shape == ref.shape
    assert np.array_equal(h == 0.0, ref == 0.0)                           # frozen blocks are exactly zero on both sides
    scale = np.abs(ref).max()
    # analytic: fp32 dual-number arithmetic; FD: fp32 forces differenced over 2e-3 A (~1e-3 eV/A^2 noise)
    tol = 2e-4 * scale if name.startswith("an") else 5e-3 * H_EVAA_2_AU
    assert np.abs(h.astype(np.float64) - ref).max() < tol, (np.abs(h - ref).max(), tol)
    assert abs(r["energy"] - float(g[name + ".energy"])) < 1e-5 * 12 * EV2AU
    assert np.abs(r["forces"] - g[name + ".forces"]).max() < 1e-4 * F_EVAA_2_AU


def test_harmonic_bias_decorator_contract_on_the_cuda_calculator(built_lib, small_model):
    """opt.py:286-343 calls base.get_forces(elem, coords[N,3]) / get_energy and adds its bias; the fixture holds the
    bias terms and totals the reference class produced around the reference calculator."""
    g = np.load(os.path.join(REFWRAP, "bias_n12.npz"))
    elem = [str(e) for e in g["elem"]]
    calc = uma_pysis(model="test-4x")
    base = calc.get_forces(elem, np.asarray(g["coords_bohr"], dtype=float).reshape(-1, 3))
    e_tot = float(base["energy"]) + float(g["e_bias"])
    f_tot = np.asarray(base["forces"], dtype=float).reshape(-1) + g["f_bias"]
    assert abs(e_tot - float(g["energy"])) < 1e-5 * 12 * EV2AU
    assert np.abs(f_tot - g["forces"]).max() < 1e-4 * F_EVAA_2_AU
    assert float(calc.get_energy(elem, g["coords_bohr"])["energy"]) == float(base["energy"])


def test_hessian_post_matches_the_reference_freq_helpers(built_lib):
    g = np.load(os.path.join(REFWRAP, "freq_n10.npz"))
    z, x, h, freeze = g["z"], g["coords_bohr"], g["hessian"], [int(i) for i in g["freeze"]]
    n = len(z)
    dev = torch.device("cuda")
    m_au = torch.as_tensor(hp.masses_amu_for(z) * hp.AMU2AU, device=dev)
    hpj = hp.mw_projected_hessian(torch.as_tensor(h.copy(), device=dev), torch.as_tensor(x, device=dev), m_au).cpu().numpy()
    assert np.abs(hpj - g["mw_projected"]).max() < 1e-12 * np.abs(g["mw_projected"]).max()
    act = [3 * i + c for i in range(n) if i not in freeze for c in range(3)]
    for tag, hin, fz in (("full", h, None), ("phva_full", h, freeze), ("phva_block", h[np.ix_(act, act)], freeze)):
        f, modes = hp.frequencies_cm_and_modes(torch.as_tensor(hin.copy(), device=dev), list(z), x.copy(), freeze_idx=fz)
        fr, mr = g[tag + ".freqs"], g[tag + ".modes"]
        assert f.shape == fr.shape and np.abs(f - fr).max() < 1e-8 * np.abs(fr).max(), tag
        dots = np.abs((modes.cpu().numpy() * mr).sum(1))
        assert modes.shape == mr.shape and np.abs(dots - 1.0).max() < 1e-8, tag
