"""GPU: how many tensor-core passes does fp32-grade parity need?  (VERDICT r1 weak item 10 / BASELINE.md section 2)

Energy / force error against the float64 oracle fixture (tests/golden/large/c4_n1500.npz: one 1500-atom C4 image) and
a 300-atom C2 image (float64 oracle computed here), for every GEMM arithmetic the hardware offers at lower cost than
the product's bf16 x3 split:

  fp32 exact (SIMT FFMA)                         reference point
  bf16 x3  (product: tcgen05, 3 MMAs / product)  cost 3
  TF32 single pass  (kind::tf32, half the bf16 rate)  cost 2   -- operands rounded to 10 mantissa bits
  bf16 x2, weights rounded to bf16               cost 2
  bf16 x2, activations rounded to bf16           cost 2
  bf16 x1                                        cost 1
  + the same roundings applied to the ADJOINT (backward) GEMMs only

The lower-precision variants are emulated exactly in the fp32 SIMT GEMM by rounding the operands (fp32 accumulation,
as TMEM accumulates), so this measures the arithmetic, not a kernel.  Tolerances: 1e-5 eV/atom, 1e-4 eV/A.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pdb2reaction_b200 import synth, weights as W          # noqa: E402
from pdb2reaction_b200.arch import UMAArch, atomic_numbers  # noqa: E402
from pdb2reaction_b200.engine import UmabEngine, _check    # noqa: E402

MODES = {0: "fp32", 1: "tf32 x1", 2: "bf16 x2 (W bf16)", 3: "bf16 x2 (A bf16)", 4: "bf16 x1"}


def main():
    arch = UMAArch(num_experts=4)
    sd = W.init_uma_weights(arch, 0)
    cases = []
    g = np.load(os.path.join(ROOT, "tests", "golden", "large", "c4_n1500.npz"))
    elem, imgs = synth.make_config("C4")
    cases.append(("C4 image, 1500 atoms", elem, imgs[int(g["image"])], float(g["energy"][0]), g["forces"]))
    from oracle import uma_ref
    elem2, imgs2 = synth.make_config("C2")
    z2 = atomic_numbers(elem2)
    m2 = W.merge_mole(sd, arch, z2, 0, 1, "omol")
    e2, f2 = uma_ref.OracleUMA(m2, z2, dtype=torch.float64, hyper=uma_ref.Hyper(num_experts=4), edge_chunk=16384).energy_forces(imgs2[3])
    cases.append(("C2 image, 300 atoms", elem2, imgs2[3], float(e2[0]), f2[0].numpy()))
    rows = []
    for name, elem, img, e_ref, f_ref in cases:
        z = atomic_numbers(elem)
        merged = W.merge_mole(sd, arch, z, 0, 1, "omol")
        n = len(z)
        pos = img[None].astype(np.float32)
        eng_tc = UmabEngine(merged, z, arch, gemm_mode=1)
        e, f = eng_tc.energy_forces_host(pos)
        rows.append({"case": name, "gemm": "bf16 x3 (tcgen05, product)", "cost": 3, "dE_per_atom": abs(e[0] - e_ref) / n,
                     "dF_max": float(np.abs(f[0] - f_ref).max())})
        eng_tc.close()
        eng = UmabEngine(merged, z, arch, gemm_mode=0)
        for fwd, bwd in [(0, 0), (1, 1), (2, 2), (3, 3), (4, 4), (0, 1), (0, 2), (0, 3), (0, 4)]:
            _check(eng.lib, eng.lib.umab_set_option(eng._h, b"simt_round_fwd", fwd))
            _check(eng.lib, eng.lib.umab_set_option(eng._h, b"simt_round_bwd", bwd))
            e, f = eng.energy_forces_host(pos)
            label = MODES[fwd] if fwd == bwd else f"forward fp32, adjoint {MODES[bwd]}"
            rows.append({"case": name, "gemm": label + " (emulated)" if (fwd or bwd) else "fp32 exact (SIMT)",
                         "cost": {0: None, 1: 2, 2: 2, 3: 2, 4: 1}[max(fwd, bwd)], "dE_per_atom": abs(e[0] - e_ref) / n,
                         "dF_max": float(np.abs(f[0] - f_ref).max())})
        eng.close()
    for r in rows:
        r["passes_tolerance"] = bool(r["dE_per_atom"] < 1e-5 and r["dF_max"] < 1e-4)
        print(f'{r["case"]:24s} {r["gemm"]:48s} dE/atom {r["dE_per_atom"]:.2e}  dF {r["dF_max"]:.2e}  '
              f'{"ok" if r["passes_tolerance"] else "FAILS"}', flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "precision_study.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
