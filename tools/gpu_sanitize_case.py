"""GPU: the case run under compute-sanitizer (memcheck / racecheck / synccheck): energy+forces of two 130-atom images
(tcgen05 CTA-pair GEMMs with their hand-rolled mbarrier / cluster / TMEM protocol, TMA loads and stores, cell-list
neighbour search, every edge kernel) and one dual-number Hessian-column pass (umab_forces_jvp)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdb2reaction_b200 import synth, weights as W          # noqa: E402
from pdb2reaction_b200.arch import UMAArch, atomic_numbers  # noqa: E402
from pdb2reaction_b200.engine import UmabEngine            # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 130
arch = UMAArch(num_experts=4)
elem, imgs = synth.make_string(n, 2, 9)
z = atomic_numbers(elem)
eng = UmabEngine(W.merge_mole(W.init_uma_weights(arch, 0), arch, z, 0, 1, "omol"), z, arch)
e, f = eng.energy_forces_host(imgs.astype(np.float32))
print("E", e, "max|F|", float(np.abs(f).max()), flush=True)
pos = torch.from_numpy(imgs[:1].astype(np.float32)).cuda()
tan = torch.zeros_like(pos)
tan[0, 3, 1] = 1.0
f2, df = eng.forces_jvp(pos, tan)
torch.cuda.synchronize()
print("jvp max|dF|", float(df.abs().max()), "launches", eng.stats()["kernel_launches"], flush=True)
assert np.isfinite(f).all() and torch.isfinite(df).all()
