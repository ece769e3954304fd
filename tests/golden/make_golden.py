"""Generate tests/golden/*.npz from the CPU oracle (float64) on deterministic synthetic inputs.

The reference ships no golden vectors and cannot be imported here, so these
fixtures pin the oracle against ITSELF over time (a regression guard for oracle edits) and give
the -m gpu tests inputs/outputs that do not depend on re-running the oracle.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import uma_ref  # noqa: E402
from pdb2reaction_b200 import synth, weights as W  # noqa: E402
from pdb2reaction_b200.arch import UMAArch, atomic_numbers  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"c1_n20": (20, 1, 1), "string_n40_b3": (40, 3, 21), "c2like_n120_b2": (120, 2, 2)}


def main():
    arch = UMAArch(num_experts=4)
    sd = W.init_uma_weights(arch, 0)
    hp = uma_ref.Hyper(num_experts=4)
    for name, (n, b, seed) in CASES.items():
        elem, imgs = synth.make_string(n, b, seed)
        z = atomic_numbers(elem)
        merged = W.merge_mole(sd, arch, z, 0, 1, "omol")
        orc = uma_ref.OracleUMA(merged, z, dtype=torch.float64, hyper=hp)
        e, f = orc.energy_forces(imgs)
        ei = orc.graph(imgs.reshape(-1, 3).astype(np.float32), [n] * b)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), elem=np.array(elem), coords=imgs,
                            energy=e.numpy(), forces=f.numpy(), edge_index=ei.astype(np.int32),
                            num_experts=4, weight_seed=0)
        print(name, e.numpy(), ei.shape)


if __name__ == "__main__":
    main()
