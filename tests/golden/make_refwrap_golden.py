"""Generate tests/golden/refwrap/*.npz by RUNNING THE REFERENCE's own wrapper code (tests/refstubs.py loads
/root/reference/pdb2reaction/{uma_pysis,freq,opt}.py unmodified) so that the -m gpu tests, which run where
/root/reference does not exist, can compare the CUDA-backed calculator with what the reference wrapper returns.

  calc_n12.npz   reference ``uma_pysis`` (predictor = float64 oracle, float32 positions / outputs as fairchem)
                 on a 12-atom cluster: get_forces; get_hessian in FiniteDifference / Analytical mode with and
                 without frozen atoms, partial block, float32 / numpy variants
  freq_n10.npz   reference ``freq._frequencies_cm_and_modes`` / ``_mw_projected_hessian`` on a seeded Hessian:
                 full, PHVA with the full Hessian, PHVA with the active block
  bias_n12.npz   reference ``opt.HarmonicBiasCalculator`` bias energy / forces for three restraint pairs

Run from the repo root (needs /root/reference):  python tests/golden/make_refwrap_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refstubs  # noqa: E402
from oracle import uma_ref  # noqa: E402
from pdb2reaction_b200 import synth, weights as W  # noqa: E402
from pdb2reaction_b200.arch import UMAArch  # noqa: E402
from pdb2reaction_b200.hessian_post import ATOMIC_MASSES  # noqa: E402

OUT = os.path.join(HERE, "refwrap")
HESS_CASES = {
    "fd": dict(),
    "fd_frozen_partial": dict(freeze_atoms=[0, 5], return_partial_hessian=True),
    "fd_frozen_f32_numpy": dict(freeze_atoms=[2], hessian_double=False, out_hess_torch=False),
    "an": dict(hessian_calc_mode="Analytical"),
    "an_frozen": dict(hessian_calc_mode="Analytical", freeze_atoms=[1, 6]),
    "an_frozen_partial_f32": dict(hessian_calc_mode="Analytical", freeze_atoms=[1, 6], return_partial_hessian=True,
                                  hessian_double=False),
}


def main():
    from typing import List, Optional, Tuple
    os.makedirs(OUT, exist_ok=True)
    torch.use_deterministic_algorithms(True)
    arch = UMAArch(num_experts=4)
    sd = W.init_uma_weights(arch, 0)
    hp = uma_ref.Hyper(num_experts=4)
    cache = {}

    def weights_for(z, charge, spin, task):
        return cache.setdefault((tuple(z), charge, spin, task), W.merge_mole(sd, arch, list(z), charge, spin, task))

    ref = refstubs.load_reference_uma_pysis(lambda *a: refstubs.OraclePredictUnit(weights_for, hp, dtype=torch.float64))

    # ---- calculator
    elem, coords = synth.make_cluster(12, 5)
    coords_bohr = coords * refstubs.ANG2BOHR
    out = dict(elem=np.array(elem), coords_bohr=coords_bohr, num_experts=4, weight_seed=0)
    r = ref.uma_pysis(device="cpu", freeze_atoms=[3, 4]).get_forces(elem, coords_bohr)
    out["forces_frozen34.energy"], out["forces_frozen34.forces"] = r["energy"], r["forces"]
    for name, kw in HESS_CASES.items():
        r = ref.uma_pysis(device="cpu", **kw).get_hessian(elem, coords_bohr)
        h = r["hessian"]
        out[name + ".energy"], out[name + ".forces"] = r["energy"], r["forces"]
        out[name + ".hessian"] = h.numpy() if isinstance(h, torch.Tensor) else h
        out[name + ".is_torch"] = isinstance(h, torch.Tensor)
        print(name, out[name + ".hessian"].shape, out[name + ".hessian"].dtype, float(np.abs(out[name + ".hessian"]).max()))
    np.savez_compressed(os.path.join(OUT, "calc_n12.npz"), **out)

    # ---- freq helpers
    names = ["_build_tr_basis", "_tr_orthonormal_basis", "_mw_projected_hessian", "_mass_weighted_hessian",
             "_frequencies_cm_and_modes", "_mw_mode_to_cart"]
    ns = dict(torch=torch, np=np, List=List, Optional=Optional, Tuple=Tuple, AMU2AU=refstubs.AMU2AU, AU2EV=refstubs.AU2EV,
              BOHR2ANG=refstubs.BOHR2ANG, units=refstubs.ASE_UNITS, atomic_masses=ATOMIC_MASSES)
    fr = refstubs.load_reference_functions("freq.py", names, ns)
    rng = np.random.default_rng(7)
    n = 10
    z = rng.choice([1, 6, 7, 8, 16], size=n)
    x = rng.normal(size=(n, 3)) * 3.0
    a = rng.normal(size=(3 * n, 3 * n))
    h = a @ a.T / (3 * n) - np.diag(rng.uniform(0.0, 0.6, 3 * n))         # a few negative curvatures
    h = 0.5 * (h + h.T)
    freeze = [1, 4, 8]
    act = [3 * i + c for i in range(n) if i not in freeze for c in range(3)]
    m_au = torch.as_tensor(ATOMIC_MASSES[z] * refstubs.AMU2AU)
    out = dict(z=z, coords_bohr=x, hessian=h, freeze=np.array(freeze))
    out["mw_projected"] = fr["_mw_projected_hessian"](torch.as_tensor(h.copy()), torch.as_tensor(x), m_au).numpy()
    for tag, hin, fz in (("full", h, None), ("phva_full", h, freeze), ("phva_block", h[np.ix_(act, act)], freeze)):
        f, modes = fr["_frequencies_cm_and_modes"](torch.as_tensor(hin.copy()), list(z), x.copy(), torch.device("cpu"),
                                                   freeze_idx=fz)
        out[tag + ".freqs"], out[tag + ".modes"] = f, modes.numpy()
        print(tag, f.shape, f[:3])
    np.savez_compressed(os.path.join(OUT, "freq_n10.npz"), **out)

    # ---- harmonic bias
    ns = dict(np=np, List=List, Optional=Optional, Tuple=Tuple, H_EVAA_2_AU=ref.H_EVAA_2_AU, ANG2BOHR=refstubs.ANG2BOHR)
    cls = refstubs.load_reference_functions("opt.py", ["HarmonicBiasCalculator"], ns)["HarmonicBiasCalculator"]
    pairs = [(0, 4, 1.9), (2, 7, 2.5), (1, 11, 3.1)]
    hb = cls(ref.uma_pysis(device="cpu"), k=7.5, pairs=pairs)
    eb, fb = hb._bias_energy_forces_bohr(coords_bohr)
    tot = hb.get_forces(elem, coords_bohr)
    np.savez_compressed(os.path.join(OUT, "bias_n12.npz"), elem=np.array(elem), coords_bohr=coords_bohr,
                        pairs=np.array(pairs), k=7.5, e_bias=eb, f_bias=fb, energy=tot["energy"], forces=tot["forces"])
    print("bias", eb, tot["energy"])


if __name__ == "__main__":
    main()
