"""Contract of the drop-in calculator, derived from the reference's pdb2reaction/uma_pysis.py
(units :127-129, :507-513; freeze :554-592; FD Hessian :595-686; formatting :515-551; mode
fallback :737-740).  Runs on the CPU with injected evaluators (tests/helpers.py)."""
import inspect

import numpy as np
import pytest
import torch

from pdb2reaction_b200 import CALC_KW, EV2AU, F_EVAA_2_AU, H_EVAA_2_AU, uma_pysis
from pdb2reaction_b200.shims import ANG2BOHR, AU2EV, BOHR2ANG
from pdb2reaction_b200 import synth
from helpers import OracleBackend, SpringBackend
from conftest import merged_for

ELEM6 = ["c", "H", "h", "O", "N", "H"]
X6 = np.array([[0, 0, 0], [1.1, 0.1, 0], [-0.3, 1.0, 0.2], [0.2, -0.9, 0.8], [1.4, 1.2, -0.5], [-1.0, -0.6, -0.7]],
              dtype=np.float64)


def test_constants_and_defaults_match_reference_text():
    assert abs(BOHR2ANG - 0.529177210544) < 1e-9 and abs(AU2EV - 27.211386245981) < 1e-8
    assert EV2AU == 1.0 / AU2EV and F_EVAA_2_AU == EV2AU / ANG2BOHR
    assert H_EVAA_2_AU == EV2AU / ANG2BOHR / ANG2BOHR
    assert list(CALC_KW) == ["charge", "spin", "model", "task_name", "device", "workers", "workers_per_node",
                             "max_neigh", "radius", "r_edges", "out_hess_torch", "freeze_atoms",
                             "hessian_calc_mode", "return_partial_hessian", "hessian_double"]
    assert CALC_KW["hessian_calc_mode"] == "FiniteDifference" and CALC_KW["return_partial_hessian"] is False
    assert CALC_KW["hessian_double"] is True and CALC_KW["out_hess_torch"] is True
    sig = inspect.signature(uma_pysis.__init__)
    for k, v in CALC_KW.items():
        p = sig.parameters[k]
        assert p.kind is inspect.Parameter.KEYWORD_ONLY and p.default == v
    assert uma_pysis.implemented_properties == ["energy", "forces", "hessian"]
    with pytest.raises(TypeError):
        uma_pysis(0)                       # keyword-only, as the reference


def test_energy_forces_units_and_shapes():
    be = SpringBackend()
    calc = uma_pysis(_backend=be, mem=2000)           # base-class kwargs are accepted (tsopt.py:745)
    assert calc.charge == 0 and calc.mult == 1
    coords_bohr = (X6 * ANG2BOHR).reshape(-1)
    e_ev, f_ev = be.evaluate(X6[None])
    r = calc.get_energy(ELEM6, coords_bohr)
    assert set(r) == {"energy"} and isinstance(r["energy"], float)
    assert abs(r["energy"] - e_ev[0] * EV2AU) < 1e-12
    assert be.calls[-1] == (1, False)                 # energy only: no force evaluation (Q1)
    r = calc.get_forces(ELEM6, coords_bohr.reshape(6, 3))     # any shape is reshaped (-1, 3)
    assert r["forces"].dtype == np.float64 and r["forces"].shape == (18,)
    assert np.allclose(r["forces"], f_ev[0].astype(np.float64).reshape(-1) * F_EVAA_2_AU, rtol=0, atol=1e-12)
    assert calc._core.elem == ["C", "H", "H", "O", "N", "H"]   # capitalised (:266)


def test_frozen_atoms_forces_are_exactly_zero_and_indices_are_deduplicated():
    calc = uma_pysis(_backend=SpringBackend(), freeze_atoms=[4, 1, 4])
    assert calc.freeze_atoms == [1, 4]
    f = calc.get_forces(ELEM6, X6 * ANG2BOHR)["forces"].reshape(6, 3)
    assert np.all(f[[1, 4]] == 0.0) and np.all(np.abs(f[[0, 2, 3, 5]]).sum(1) > 0)
    fb = calc.get_forces_batch(ELEM6, np.stack([X6, X6 + 0.01]).reshape(2, -1) * ANG2BOHR)["forces"]
    assert fb.shape == (2, 18) and np.all(fb.reshape(2, 6, 3)[:, [1, 4]] == 0.0)


def test_fd_hessian_matches_analytic_hessian_of_toy_potential():
    be = SpringBackend()
    calc = uma_pysis(_backend=be)
    r = calc.get_hessian(ELEM6, X6 * ANG2BOHR)
    h = r["hessian"]
    assert isinstance(h, torch.Tensor) and h.dtype == torch.float64 and h.shape == (18, 18)
    ref = be.hessian(X6.astype(np.float32).astype(np.float64)) * H_EVAA_2_AU
    assert np.abs(h.numpy() - ref).max() < 2e-4 * np.abs(ref).max()
    assert (h - h.T).abs().max() == 0.0
    # 1 base point + all 2 x 18 displacements, evaluated as batches (not 37 single calls)
    assert sum(b for b, _ in be.calls) == 37 and len(be.calls) <= 3
    assert set(r) == {"energy", "forces", "hessian"}


@pytest.mark.parametrize("mode", ["FiniteDifference", "Analytical", " analytic ", "", None, "bogus"])
def test_hessian_mode_strings_never_fail(mode):
    calc = uma_pysis(_backend=SpringBackend(), hessian_calc_mode=mode)
    with pytest.warns(RuntimeWarning) if (mode or "").strip().lower() in ("analytical", "analytic") else _nowarn():
        h = calc.get_hessian(ELEM6, X6 * ANG2BOHR)["hessian"]
    assert h.shape == (18, 18)


class _nowarn:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def test_hessian_freeze_semantics_full_and_partial():
    be = SpringBackend()
    frozen = [0, 3]
    full = uma_pysis(_backend=be, freeze_atoms=frozen).get_hessian(ELEM6, X6 * ANG2BOHR)["hessian"].numpy()
    ref = be.hessian(X6.astype(np.float32).astype(np.float64)) * H_EVAA_2_AU
    fdof = [3 * a + c for a in frozen for c in range(3)]
    adof = [k for k in range(18) if k not in fdof]
    # frozen columns are never filled, then 0.5 (H + H^T): frozen-frozen block 0, frozen/active halved (Q6)
    assert np.all(full[np.ix_(fdof, fdof)] == 0.0)
    assert np.abs(full[np.ix_(fdof, adof)] - 0.5 * ref[np.ix_(fdof, adof)]).max() < 2e-4 * np.abs(ref).max()
    assert np.abs(full[np.ix_(adof, adof)] - ref[np.ix_(adof, adof)]).max() < 2e-4 * np.abs(ref).max()
    part = uma_pysis(_backend=be, freeze_atoms=frozen, return_partial_hessian=True).get_hessian(
        ELEM6, X6 * ANG2BOHR)["hessian"].numpy()
    assert part.shape == (12, 12)
    assert np.abs(part - ref[np.ix_(adof, adof)]).max() < 2e-4 * np.abs(ref).max()


def test_hessian_output_formatting_flags():
    h = uma_pysis(_backend=SpringBackend(), out_hess_torch=False).get_hessian(ELEM6, X6 * ANG2BOHR)["hessian"]
    assert isinstance(h, np.ndarray) and h.dtype == np.float64
    h = uma_pysis(_backend=SpringBackend(), hessian_double=False).get_hessian(ELEM6, X6 * ANG2BOHR)["hessian"]
    assert isinstance(h, torch.Tensor) and h.dtype == torch.float32
    h2 = uma_pysis(_backend=SpringBackend()).get_hessian(ELEM6, X6 * ANG2BOHR)["hessian"]
    h2 += 1.0                                   # callers mutate the result in place (freq.py:179-180)


def test_element_list_is_latched_on_first_call():
    calc = uma_pysis(_backend=SpringBackend())
    calc.get_energy(ELEM6, X6 * ANG2BOHR)
    calc.get_energy(["H"] * 6, X6 * ANG2BOHR)   # silently reuses the first list (Q4)
    assert calc._core.elem[0] == "C"
    with pytest.raises(ValueError):
        calc.get_energy(ELEM6, np.zeros(9))     # wrong atom count is an error, not a reshape


def test_calculator_on_oracle_backend_batch_equals_loop(state4, arch4, hyper4):
    from oracle import uma_ref
    elem, imgs = synth.make_string(16, 3, 9)
    z, merged = merged_for(state4, arch4, elem)
    be = OracleBackend(uma_ref.OracleUMA(merged, z, dtype=torch.float32, hyper=hyper4))
    calc = uma_pysis(_backend=be)
    rb = calc.get_forces_batch(elem, imgs.reshape(3, -1) * ANG2BOHR)
    for k in range(3):
        r1 = calc.get_forces(elem, imgs[k] * ANG2BOHR)
        assert abs(r1["energy"] - rb["energy"][k]) < 1e-6 * EV2AU * 16
        assert np.abs(r1["forces"] - rb["forces"][k]).max() < 1e-5 * F_EVAA_2_AU
    eb = calc.get_energy_batch(elem, imgs.reshape(3, -1) * ANG2BOHR)["energy"]
    assert np.abs(eb - rb["energy"]).max() < 1e-9


def test_fd_hessian_of_oracle_matches_its_analytic_hessian(state4, arch4, hyper4):
    """The reference's two modes agree: FD (h = 1e-3 A, :600) vs autograd Hessian (:402-409)."""
    from oracle import uma_ref
    elem, coords = synth.make_cluster(8, 3)
    z, merged = merged_for(state4, arch4, elem)
    orc = uma_ref.OracleUMA(merged, z, dtype=torch.float64, hyper=hyper4)
    calc = uma_pysis(_backend=OracleBackend(orc))
    h_fd = calc.get_hessian(elem, coords * ANG2BOHR)["hessian"].numpy() / H_EVAA_2_AU
    h_an = orc.hessian(coords).reshape(24, 24).numpy()
    h_an = 0.5 * (h_an + h_an.T)
    # float32 forces (the backend interface returns fp32) differenced over 2e-3 A: ~1e-3 eV/A^2 noise
    assert np.abs(h_fd - h_an).max() < 5e-3


def test_model_id_without_checkpoint_raises_and_random_weights_are_opt_in(monkeypatch, arch4):
    """The reference fails hard when the checkpoint of a model ID cannot be obtained (uma_pysis.py:246-250); a drop-in
    must not silently answer with random weights."""
    from pdb2reaction_b200 import calculator as cm
    monkeypatch.delenv("UMAB_WEIGHTS", raising=False)
    monkeypatch.delenv("UMAB_ALLOW_RANDOM_WEIGHTS", raising=False)
    monkeypatch.setattr(cm, "_state_cache", {})
    with pytest.raises(FileNotFoundError, match="random:uma-s-1p1"):
        cm.load_model_state("uma-s-1p1", arch4)
    state, tr = cm.load_model_state("random:uma-s-1p1", arch4)
    assert tr.is_identity and "sphere_embedding.weight" in state
    monkeypatch.setenv("UMAB_ALLOW_RANDOM_WEIGHTS", "1")
    state2, _ = cm.load_model_state("uma-s-1p1", arch4)
    assert torch.equal(state2["sphere_embedding.weight"], state["sphere_embedding.weight"])


def test_engine_cache_is_bounded_lru(monkeypatch):
    from pdb2reaction_b200 import calculator as cm
    monkeypatch.setattr(cm, "_engine_cache", cm.OrderedDict())
    monkeypatch.setattr(cm, "ENGINE_CACHE_MAX", 2)
    for i in range(4):
        cm._engine_cache_put(("w", i, 0), f"eng{i}")
    cm._engine_cache_put(("w", 9, 1), "other-device")                  # the bound is per device
    assert list(cm._engine_cache) == [("w", 2, 0), ("w", 3, 0), ("w", 9, 1)]
    assert cm._engine_cache_get(("w", 2, 0)) == "eng2"                 # a hit refreshes the entry
    cm._engine_cache_put(("w", 5, 0), "eng5")
    assert list(cm._engine_cache) == [("w", 9, 1), ("w", 2, 0), ("w", 5, 0)]
