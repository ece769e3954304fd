"""The reference's OWN wrapper code, executed here, against the repo's drop-in calculator -- bit for bit.

``/root/reference/pdb2reaction/uma_pysis.py`` is loaded unmodified under stand-ins for its third-party imports
(tests/refstubs.py); its ``pretrained_mlip.get_predict_unit`` is answered by an oracle-backed predictor.  The repo
calculator gets the same oracle injected as its backend, so every difference would be a difference in the wrapper
semantics: units (:127-129, :507-513), frozen forces (:561-567), FD Hessian (:595-686), analytic Hessian
(:394-415) + active trim (:569-592), formatting (:515-551), mode selection (:737-740).  Likewise the reference's
``freq.py:122-381`` helpers and ``opt.py:286-343`` ``HarmonicBiasCalculator`` are exec'd as they stand and compared
with ``oracle/hessian_ref.py`` / composed with the repo calculator.

CPU only; skipped where /root/reference does not exist (the GPU box) -- there the committed fixtures under
tests/golden/refwrap/ (written by tests/golden/make_refwrap_golden.py from these same reference objects) stand in.
"""
import numpy as np
import pytest
import torch

import refstubs
from pdb2reaction_b200 import calculator as repo_mod
from pdb2reaction_b200 import uma_pysis as repo_uma_pysis
from pdb2reaction_b200 import weights as W
from pdb2reaction_b200.arch import atomic_numbers

pytestmark = pytest.mark.skipif(not refstubs.HAVE_REFERENCE, reason="/root/reference not present")

ELEM = ["c", "H", "h", "O", "N", "H", "C", "h"]
X = np.array([[0, 0, 0], [1.1, 0.1, 0], [-0.3, 1.0, 0.2], [0.2, -0.9, 0.8], [1.4, 1.2, -0.5], [-1.0, -0.6, -0.7],
              [2.3, -0.4, 0.6], [2.9, 0.3, 1.2]], dtype=np.float64)


@pytest.fixture(autouse=True)
def _deterministic_cpu_autograd():
    """torch's CPU index_add / gather adjoints accumulate in thread order unless told otherwise; both sides run the
    same oracle, so bit equality needs the deterministic kernels."""
    before = torch.are_deterministic_algorithms_enabled()
    torch.use_deterministic_algorithms(True)
    yield
    torch.use_deterministic_algorithms(before)


class SerialOracleBackend:
    """The repo calculator's backend interface over the oracle, ONE image per oracle call (the reference's calling
    pattern) so both sides execute identical CPU arithmetic."""

    def __init__(self, oracle):
        self.oracle = oracle
        self.n_eval = 0

    def evaluate(self, coords_ang, forces=True):
        es, fs = [], []
        for c in coords_ang:
            self.n_eval += 1
            e, f = self.oracle.energy_forces(c, forces=True)     # the reference always differentiates (pos.requires_grad_)
            es.append(e.double().numpy()[0])
            fs.append(f.float().numpy()[0])
        return np.array(es), (np.stack(fs) if forces else None)

    def hessian_columns(self, coord_ang, dofs):
        n = coord_ang.shape[0]
        h = self.oracle.hessian(coord_ang).reshape(3 * n, 3 * n)
        return np.ascontiguousarray(h[:, list(dofs)].T.float().numpy())


@pytest.fixture(scope="module")
def sides(state4, arch4, hyper4):
    """(reference module, factory of reference calculators, factory of repo calculators)."""
    from oracle import uma_ref
    cache = {}

    def weights_for(z, charge, spin, task):
        key = (tuple(z), charge, spin, task)
        if key not in cache:
            cache[key] = W.merge_mole(state4, arch4, list(z), charge, spin, task)
        return cache[key]

    units = []

    def predict_factory(model, device, workers):
        u = refstubs.OraclePredictUnit(weights_for, hyper4)
        units.append(u)
        return u

    ref = refstubs.load_reference_uma_pysis(predict_factory)

    def make_ref(**kw):
        return ref.uma_pysis(device="cpu", **kw)

    def make_repo(elem, charge=0, spin=1, **kw):
        z = atomic_numbers([e.capitalize() for e in elem])
        orc = uma_ref.OracleUMA(weights_for(z, charge, spin, kw.get("task_name", "omol")), z, charge=charge, spin=spin,
                                task_name=kw.get("task_name", "omol"), dtype=torch.float32, hyper=hyper4)
        return repo_uma_pysis(charge=charge, spin=spin, _backend=SerialOracleBackend(orc), **kw)

    return ref, make_ref, make_repo


def _same(a, b):
    if isinstance(a, torch.Tensor):
        assert isinstance(b, torch.Tensor) and a.dtype == b.dtype and a.shape == b.shape
        return torch.equal(a, b)
    a, b = np.asarray(a), np.asarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)


def test_constants_and_defaults_equal_the_reference_module(sides):
    ref, _, _ = sides
    for name in ("EV2AU", "F_EVAA_2_AU", "H_EVAA_2_AU"):
        assert getattr(ref, name) == getattr(repo_mod, name), name
    assert ref.CALC_KW == repo_mod.CALC_KW and list(ref.CALC_KW) == list(repo_mod.CALC_KW)
    assert ref.GEOM_KW_DEFAULT == repo_mod.GEOM_KW_DEFAULT
    assert ref.uma_pysis.implemented_properties == repo_uma_pysis.implemented_properties
    from pdb2reaction_b200 import shims
    assert (shims.BOHR2ANG, shims.ANG2BOHR, shims.AU2EV) == (refstubs.BOHR2ANG, refstubs.ANG2BOHR, refstubs.AU2EV)


@pytest.mark.parametrize("kw", [{}, {"freeze_atoms": [4, 1, 4]}, {"charge": -1, "spin": 2}])
def test_energy_and_forces_bit_equal(sides, kw):
    _, make_ref, make_repo = sides
    a, b = make_ref(**kw), make_repo(ELEM, **kw)
    coords = (X * refstubs.ANG2BOHR).reshape(-1)
    ra, rb = a.get_energy(ELEM, coords), b.get_energy(ELEM, coords)
    assert set(ra) == set(rb) == {"energy"} and type(ra["energy"]) is type(rb["energy"]) is float
    assert ra["energy"] == rb["energy"]
    ra, rb = a.get_forces(ELEM, coords.reshape(-1, 3)), b.get_forces(ELEM, coords.reshape(-1, 3))
    assert set(ra) == set(rb) == {"energy", "forces"}
    assert ra["energy"] == rb["energy"] and _same(ra["forces"], rb["forces"])
    assert ra["forces"].dtype == np.float64 and ra["forces"].shape == (24,)
    assert a.freeze_atoms == b.freeze_atoms and a._core.elem == b._core.elem
    assert (a.charge, a.mult) == (b.charge, b.mult)
    # the batched extension returns, per image, exactly what the reference returns one image at a time
    imgs = np.stack([X, X + 0.05, X - 0.03])
    rbb = b.get_forces_batch(ELEM, imgs.reshape(3, -1) * refstubs.ANG2BOHR)
    for i in range(3):
        ri = a.get_forces(ELEM, imgs[i] * refstubs.ANG2BOHR)
        assert ri["energy"] == rbb["energy"][i] and _same(ri["forces"], rbb["forces"][i])


HESS_CASES = [
    dict(),
    dict(freeze_atoms=[0, 5]),
    dict(freeze_atoms=[0, 5], return_partial_hessian=True),
    dict(freeze_atoms=[2], hessian_double=False, out_hess_torch=False),
    dict(hessian_calc_mode="bogus", hessian_double=False),
    dict(hessian_calc_mode=None, out_hess_torch=False),
    dict(hessian_calc_mode="Analytical"),
    dict(hessian_calc_mode=" analytic ", freeze_atoms=[1, 6]),
    dict(hessian_calc_mode="ANALYTICAL", freeze_atoms=[1, 6], return_partial_hessian=True, hessian_double=False),
    dict(hessian_calc_mode="Analytical", freeze_atoms=[3], out_hess_torch=False),
]


@pytest.mark.parametrize("kw", HESS_CASES, ids=[",".join(f"{k}={v}" for k, v in c.items()) or "default"
                                                for c in HESS_CASES])
def test_hessian_bit_equal(sides, kw):
    _, make_ref, make_repo = sides
    a, b = make_ref(**kw), make_repo(ELEM, **kw)
    coords = X * refstubs.ANG2BOHR
    ra, rb = a.get_hessian(ELEM, coords), b.get_hessian(ELEM, coords)
    assert set(ra) == set(rb) == {"energy", "forces", "hessian"}
    assert ra["energy"] == rb["energy"] and _same(ra["forces"], rb["forces"])
    assert type(ra["hessian"]) is type(rb["hessian"])
    assert _same(ra["hessian"], rb["hessian"]), float(np.abs(np.asarray(ra["hessian"]) - np.asarray(rb["hessian"])).max())


def test_reference_workers_gt_1_forces_fd_and_matches_repo_fd(sides, state4, arch4, hyper4):
    """workers>1 in the reference = ParallelMLIPPredictUnit without ``.model`` -> FD forced (uma_pysis.py:737).
    The repo's FD mode returns the same Hessian; its Analytical mode stays available (documented difference)."""
    _, _, make_repo = sides
    made = []

    class Parallel:                                    # fairchem ParallelMLIPPredictUnit stand-in: no .model
        def __init__(self, **kw):
            self.kw = kw
            z_cache = {}

            def weights_for(z, charge, spin, task):
                return z_cache.setdefault((tuple(z), charge, spin, task),
                                          W.merge_mole(state4, arch4, list(z), charge, spin, task))
            self._u = refstubs.OraclePredictUnit(weights_for, hyper4)
            made.append(self)

        def predict(self, batch):
            if batch.max_neigh is None:                # no backbone defaults reachable (uma_pysis.py:298-309)
                batch.max_neigh = hyper4.max_neighbors
            return self._u.predict(batch)

    ref = refstubs.load_reference_uma_pysis(lambda *a: None, parallel_cls=Parallel)
    a = ref.uma_pysis(device="cpu", workers=2, workers_per_node=2, hessian_calc_mode="Analytical")
    b = make_repo(ELEM, hessian_calc_mode="FiniteDifference")
    coords = X * refstubs.ANG2BOHR
    ra, rb = a.get_hessian(ELEM, coords), b.get_hessian(ELEM, coords)
    assert made and made[0].kw["num_workers"] == 2 and made[0].kw["num_workers_per_node"] == 2
    assert a._core.parallel_predict and not a._core.has_torch_model
    assert _same(ra["hessian"], rb["hessian"]) and _same(ra["forces"], rb["forces"])


# ---------------------------------------------------------------------------------------- freq.py:122-381
FREQ_NAMES = ["_build_tr_basis", "_tr_orthonormal_basis", "_mw_projected_hessian", "_mass_weighted_hessian",
              "_frequencies_cm_and_modes", "_mw_mode_to_cart"]


@pytest.fixture(scope="module")
def freq_ref():
    from typing import List, Optional, Tuple
    from pdb2reaction_b200.hessian_post import ATOMIC_MASSES
    ns = dict(torch=torch, np=np, List=List, Optional=Optional, Tuple=Tuple, AMU2AU=refstubs.AMU2AU,
              AU2EV=refstubs.AU2EV, BOHR2ANG=refstubs.BOHR2ANG, units=refstubs.ASE_UNITS, atomic_masses=ATOMIC_MASSES)
    return refstubs.load_reference_functions("freq.py", FREQ_NAMES, ns)


def _toy_hessian(n, seed):
    rng = np.random.default_rng(seed)
    z = rng.choice([1, 6, 7, 8, 16], size=n)
    x = rng.normal(size=(n, 3)) * 3.0
    a = rng.normal(size=(3 * n, 3 * n))
    return z, x, a @ a.T / (3 * n) + np.diag(rng.uniform(0.1, 1.0, 3 * n))


def test_reference_freq_helpers_pin_the_hessian_oracle(freq_ref):
    from oracle import hessian_ref
    from pdb2reaction_b200.hessian_post import masses_amu_for
    z, x, h = _toy_hessian(9, 0)
    m_amu = masses_amu_for(z)
    m_au = torch.as_tensor(m_amu * refstubs.AMU2AU)
    xt = torch.as_tensor(x)
    b_ref = freq_ref["_build_tr_basis"](xt, m_au).numpy()
    assert np.abs(b_ref - hessian_ref.build_tr_basis(x, m_au.numpy())).max() < 1e-12 * np.abs(b_ref).max()
    q_ref, r_ref = freq_ref["_tr_orthonormal_basis"](xt, m_au)
    q, r = hessian_ref.tr_orthonormal_basis(x, m_au.numpy())
    assert r == r_ref == 6
    assert np.abs(q_ref.numpy() @ q_ref.numpy().T - q @ q.T).max() < 1e-12        # same subspace
    hp_ref = freq_ref["_mw_projected_hessian"](torch.as_tensor(h.copy()), xt, m_au).numpy()
    hp = hessian_ref.mw_projected_hessian(h, x, m_au.numpy())
    assert np.abs(hp_ref - hp).max() < 1e-12 * np.abs(hp).max()
    for freeze, partial in ((None, False), ([1, 4], False), ([1, 4], True)):
        hin = h
        if partial:
            act = [3 * i + c for i in range(9) if i not in freeze for c in range(3)]
            hin = h[np.ix_(act, act)]
        f_ref, modes_ref = freq_ref["_frequencies_cm_and_modes"](torch.as_tensor(hin.copy()), list(z), x.copy(),
                                                                 torch.device("cpu"), freeze_idx=freeze)
        f, modes = hessian_ref.frequencies_cm_and_modes(hin, m_amu, x, freeze_idx=freeze)
        assert f_ref.shape == f.shape and np.abs(f_ref - f).max() < 1e-8 * np.abs(f).max()
        assert modes_ref.shape == modes.shape
        # eigenvectors up to sign
        dots = np.abs((modes_ref.numpy() * modes).sum(1))
        assert np.abs(dots - 1.0).max() < 1e-8
    v = torch.as_tensor(np.random.default_rng(1).normal(size=27))
    assert np.abs(freq_ref["_mw_mode_to_cart"](v, m_au) - hessian_ref.mw_mode_to_cart(v.numpy(), m_au.numpy())).max() < 1e-14


# ---------------------------------------------------------------------------------------- opt.py:286-343 (a10)
@pytest.fixture(scope="module")
def harmonic_bias_cls(sides):
    from typing import List, Optional, Tuple
    ref, _, _ = sides
    ns = dict(np=np, List=List, Optional=Optional, Tuple=Tuple, H_EVAA_2_AU=ref.H_EVAA_2_AU, ANG2BOHR=refstubs.ANG2BOHR)
    return refstubs.load_reference_functions("opt.py", ["HarmonicBiasCalculator"], ns)["HarmonicBiasCalculator"]


def test_reference_harmonic_bias_composes_with_the_repo_calculator(sides, harmonic_bias_cls):
    _, make_ref, make_repo = sides
    pairs = [(0, 4, 1.9), (2, 7, 2.5), (1, 99, 1.0)]               # the out-of-range pair is skipped (opt.py:305)
    a = harmonic_bias_cls(make_ref(freeze_atoms=[3]), k=7.5, pairs=pairs)
    b = harmonic_bias_cls(make_repo(ELEM, freeze_atoms=[3]), k=7.5, pairs=pairs)
    coords = (X * refstubs.ANG2BOHR).reshape(-1)
    ra, rb = a.get_forces(ELEM, coords), b.get_forces(ELEM, coords)
    assert ra["energy"] == rb["energy"] and _same(ra["forces"], rb["forces"])
    assert a.get_energy(ELEM, coords)["energy"] == b.get_energy(ELEM, coords)["energy"]
    e, g = b.get_energy_and_gradient(ELEM, coords)
    assert e == rb["energy"] and np.array_equal(g, -rb["forces"])
    # bias really is on top of the base result
    base = b.base.get_forces(ELEM, coords)
    eb, fb = b._bias_energy_forces_bohr(coords)
    assert eb > 0 and rb["energy"] == float(base["energy"]) + eb and np.array_equal(rb["forces"], base["forces"] + fb)
    # attribute forwarding (opt.py:342-343) reaches the repo calculator, incl. the Hessian entry point
    assert b.freeze_atoms == [3] and b.implemented_properties == ["energy", "forces", "hessian"]
    assert _same(a.get_hessian(ELEM, coords)["hessian"], b.get_hessian(ELEM, coords)["hessian"])
