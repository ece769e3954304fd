"""GPU: ONE single-image energy+force call of a BASELINE config after warm-up (for an ncu launch list: run with
UMAB_CUDA_GRAPHS=0 so the kernels appear individually)."""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdb2reaction_b200 import synth, uma_pysis  # noqa: E402
from pdb2reaction_b200.shims import ANG2BOHR  # noqa: E402

warnings.simplefilter("ignore")
name = sys.argv[1] if len(sys.argv) > 1 else "C1"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
elem, imgs = synth.make_config(name)
c = imgs.reshape(imgs.shape[0], -1) * ANG2BOHR
calc = uma_pysis(model="random:uma-s-1p1")
for _ in range(reps):
    r = calc.get_forces(elem, c[0])
torch.cuda.synchronize()
print(name, r["energy"], float(np.abs(r["forces"]).max()))
