"""-m gpu: the drop-in calculator on the real CUDA backend."""
import numpy as np
import pytest
import torch

from pdb2reaction_b200 import synth, uma_pysis, EV2AU, F_EVAA_2_AU, H_EVAA_2_AU
from pdb2reaction_b200 import calculator as calc_mod
from pdb2reaction_b200.shims import ANG2BOHR
from conftest import merged_for

pytestmark = pytest.mark.gpu


def test_get_forces_matches_oracle_in_atomic_units(built_lib, small_model, state4, arch4, hyper4):
    from oracle import uma_ref
    elem, imgs = synth.make_string(30, 3, 13)
    z, merged = merged_for(state4, arch4, elem)
    orc = uma_ref.OracleUMA(merged, z, dtype=torch.float32, hyper=hyper4)
    e_ref, f_ref = orc.energy_forces(imgs)
    calc = uma_pysis(model="test-4x", freeze_atoms=[2, 5])
    r = calc.get_forces(elem, imgs[0] * ANG2BOHR)
    assert abs(r["energy"] - e_ref[0].item() * EV2AU) < 1e-5 * 30 * EV2AU
    ref = f_ref[0].double().numpy().copy()
    ref[[2, 5]] = 0.0
    assert np.abs(r["forces"] - ref.reshape(-1) * F_EVAA_2_AU).max() < 1e-4 * F_EVAA_2_AU
    rb = calc.get_forces_batch(elem, imgs.reshape(3, -1) * ANG2BOHR)
    assert np.abs(rb["energy"] - e_ref.double().numpy() * EV2AU).max() < 1e-5 * 30 * EV2AU
    assert np.array_equal(rb["forces"][0], r["forces"])
    e_only = calc.get_energy(elem, imgs[0] * ANG2BOHR)["energy"]
    assert e_only == r["energy"]
    # a second calculator of the same composition reuses the cached engine
    n_eng = len(calc_mod._engine_cache)
    uma_pysis(model="test-4x").get_energy(elem, imgs[0] * ANG2BOHR)
    assert len(calc_mod._engine_cache) == n_eng


def test_fd_hessian_against_oracle_analytic_hessian(built_lib, small_model, state4, arch4, hyper4):
    from oracle import uma_ref
    elem, coords = synth.make_cluster(10, 3)
    z, merged = merged_for(state4, arch4, elem)
    h_ref = uma_ref.OracleUMA(merged, z, dtype=torch.float64, hyper=hyper4).hessian(coords).reshape(30, 30).numpy()
    h_ref = 0.5 * (h_ref + h_ref.T) * H_EVAA_2_AU
    calc = uma_pysis(model="test-4x")
    r = calc.get_hessian(elem, coords * ANG2BOHR)
    h = r["hessian"]
    assert h.is_cuda and h.dtype == torch.float64 and h.shape == (30, 30)
    # fp32 forces differenced over 2e-3 A: ~1e-3 eV/A^2 noise on O(10) eV/A^2 entries
    assert np.abs(h.cpu().numpy() - h_ref).max() < 5e-3 * H_EVAA_2_AU * max(1.0, np.abs(h_ref / H_EVAA_2_AU).max() / 10)
    part = uma_pysis(model="test-4x", freeze_atoms=[0, 1, 2], return_partial_hessian=True, out_hess_torch=False,
                     hessian_double=False).get_hessian(elem, coords * ANG2BOHR)["hessian"]
    assert isinstance(part, np.ndarray) and part.shape == (21, 21) and part.dtype == np.float32


def test_device_cpu_is_refused(built_lib):
    with pytest.raises(RuntimeError, match="no CPU"):
        uma_pysis(device="cpu").get_energy(["H", "H"], [0, 0, 0, 0, 0, 1.4])


def test_analytic_hessian_columns_match_oracle_double_backward(built_lib, small_model, state4, arch4, hyper4):
    """hessian_calc_mode='Analytical': dual-number passes through the CUDA kernels against
    torch.autograd.functional.hessian of the float64 oracle (reference uma_pysis.py:402-409)."""
    from oracle import uma_ref
    elem, coords = synth.make_cluster(12, 5)
    z, merged = merged_for(state4, arch4, elem)
    h_ref = uma_ref.OracleUMA(merged, z, dtype=torch.float64, hyper=hyper4).hessian(coords).reshape(36, 36).numpy()
    scale = np.abs(h_ref).max()
    calc = uma_pysis(model="test-4x", hessian_calc_mode="Analytical", hessian_double=False)
    raw = calc._core if calc._core else None
    r = calc.get_hessian(elem, coords * ANG2BOHR)
    h = r["hessian"]
    assert h.is_cuda and h.dtype == torch.float32 and h.shape == (36, 36)
    h_sym_ref = 0.5 * (h_ref + h_ref.T) * H_EVAA_2_AU
    assert np.abs(h.cpu().numpy() - h_sym_ref).max() < 2e-4 * scale * H_EVAA_2_AU
    # un-symmetrised columns straight from the backend: the analytic Hessian is symmetric by itself
    cols = calc._core.backend.hessian_columns(coords, list(range(36)))
    assert np.abs(cols.T - h_ref).max() < 2e-4 * scale
    assert np.abs(cols - cols.T).max() < 2e-4 * scale
    # frozen atoms: frozen columns zero, then symmetrised (Q6), partial block on request
    frozen = [0, 7]
    hf = uma_pysis(model="test-4x", hessian_calc_mode="analytic", freeze_atoms=frozen).get_hessian(
        elem, coords * ANG2BOHR)["hessian"].cpu().numpy() / H_EVAA_2_AU
    fd = [3 * a + c for a in frozen for c in range(3)]
    ad = [k for k in range(36) if k not in fd]
    assert np.all(hf[np.ix_(fd, fd)] == 0.0)
    assert np.abs(hf[np.ix_(fd, ad)] - 0.5 * h_ref[np.ix_(fd, ad)]).max() < 2e-4 * scale
    assert np.abs(hf[np.ix_(ad, ad)] - 0.5 * (h_ref + h_ref.T)[np.ix_(ad, ad)]).max() < 2e-4 * scale


def test_sharded_analytic_hessian_equals_get_hessian_on_one_rank(built_lib, small_model):
    """sharding.sharded_analytic_hessian (bench.py --hessian --hessian-mode analytic) on the CUDA backend, one rank:
    the same bits as uma_pysis.get_hessian(hessian_calc_mode="Analytical"), also with frozen atoms and the active block
    (the world-2 gather is covered by the gloo test in test_sharding.py)."""
    from pdb2reaction_b200.sharding import sharded_analytic_hessian
    elem, coords = synth.make_cluster(12, 5)
    for kw in (dict(), dict(freeze_atoms=[0, 7]), dict(freeze_atoms=[0, 7], return_partial_hessian=True)):
        a = uma_pysis(model="test-4x", hessian_calc_mode="Analytical", out_hess_torch=True, **kw).get_hessian(
            elem, coords * ANG2BOHR)
        b = sharded_analytic_hessian(uma_pysis(model="test-4x", hessian_calc_mode="Analytical", out_hess_torch=True, **kw),
                                     elem, coords * ANG2BOHR)
        assert torch.equal(a["hessian"], b["hessian"]) and a["energy"] == b["energy"]
        assert np.array_equal(a["forces"], b["forces"])


@pytest.mark.parametrize("n_atoms,n_cols,kw", [(20, 7, {}), (130, 5, {}), (130, 6, {"workspace_bytes": 9600 * 4 * 9000}),
                                               (130, 5, {"images_per_chunk": 2.5}),
                                               (130, 7, {"images_per_chunk": 3.5, "store_bytes": -1})])
def test_jvp_shared_base_gives_identical_bits(built_lib, state4, arch4, n_atoms, n_cols, kw):
    """Hessian columns of one base geometry ("jvp_shared_base": value planes computed and read once per batch): the same
    bits as the plain dual-number batch -- fp32 split-K path (20 atoms), tensor-core path in one closed chunk, in chunks
    of one image, in chunks of TWO images (2 + 2 + 1: the value pointers step back relative to the chunk) and in the
    recompute mode with chunks of two images (130 atoms) -- and a batch whose images differ is detected on the device
    and evaluated the plain way."""
    from pdb2reaction_b200.engine import UmabEngine
    elem, coords = synth.make_cluster(n_atoms, 21)
    z, merged = merged_for(state4, arch4, elem)
    kw = dict(kw)
    if "images_per_chunk" in kw:
        probe = UmabEngine(merged, z, arch4)
        e_img = probe.graph(torch.from_numpy(coords.astype(np.float32)).cuda().unsqueeze(0))[0].shape[0]
        per_edge = (8192 + (2560 if kw.get("store_bytes", 0) < 0 else 0)) * 4 * 2      # engine.cu EDGE_WS_FLOATS (+ recompute), two planes
        kw["workspace_bytes"] = int(kw.pop("images_per_chunk") * e_img) * per_edge
        del probe
    eng = UmabEngine(merged, z, arch4, **kw)
    pos = torch.from_numpy(coords.astype(np.float32)).cuda().unsqueeze(0).expand(n_cols, n_atoms, 3).contiguous()
    tan = torch.zeros(n_cols, 3 * n_atoms, device="cuda")
    tan[torch.arange(n_cols), torch.arange(n_cols) * 5 + 1] = 1.0
    tan = tan.view(n_cols, n_atoms, 3)
    f0, df0 = eng.forces_jvp(pos, tan)
    eng.set_option("jvp_shared_base", 1)
    c0 = eng.get_option("dedupe_gemms")
    f1, df1 = eng.forces_jvp(pos, tan)
    c1 = eng.get_option("dedupe_gemms")
    assert c1 > c0, "the value-plane GEMMs were not de-duplicated"
    assert torch.equal(f0, f1) and torch.equal(df0, df1)
    # images that differ: the device check turns the de-duplication off for the call
    pos2 = pos.clone()
    pos2[1, 3, 0] += 1e-3
    f2, df2 = eng.forces_jvp(pos2, tan)
    assert eng.get_option("dedupe_gemms") == c1
    eng.set_option("jvp_shared_base", 0)
    f3, df3 = eng.forces_jvp(pos2, tan)
    assert torch.equal(f2, f3) and torch.equal(df2, df3)


def test_forces_jvp_matches_central_difference_of_cuda_forces(built_lib, state4, arch4):
    from pdb2reaction_b200.engine import UmabEngine
    elem, imgs = synth.make_string(200, 2, 17)
    z, merged = merged_for(state4, arch4, elem)
    eng = UmabEngine(merged, z, arch4)
    rng = np.random.default_rng(1)
    t = rng.normal(size=imgs.shape)
    t /= np.linalg.norm(t.reshape(2, -1), axis=1)[:, None, None]
    pos = torch.from_numpy(imgs.astype(np.float32)).cuda()
    f, df = eng.forces_jvp(pos, torch.from_numpy(t.astype(np.float32)).cuda())
    e0, f0 = eng.energy_forces(pos)
    assert (f - f0).abs().max() < 1e-4                       # value plane = the plain force path (other FMA order)
    # t is dense with |t| = 1, so a step of 0.05 moves every coordinate by ~2e-3 A: large against the
    # fp32 rounding of the displaced positions (ulp ~1e-6), small for a central difference
    h = 5e-2
    _, fp = eng.energy_forces(torch.from_numpy((imgs + h * t).astype(np.float32)).cuda())
    _, fm = eng.energy_forces(torch.from_numpy((imgs - h * t).astype(np.float32)).cuda())
    fd = (fp - fm) / (2 * h)
    assert (df - fd).abs().max() < 1e-2 * max(1.0, fd.abs().max().item())
