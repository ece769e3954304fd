"""GPU experiment: can tensor-bound GEMM phases of one half of the string overlap with the HBM-bound edge kernels of the
other half?  Two engines (own workspaces, own streams, two host threads) each evaluate 16 of the 32 C4 images
concurrently; compared with one engine evaluating all 32 serially."""
import json
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdb2reaction_b200 import synth, weights as W          # noqa: E402
from pdb2reaction_b200.arch import UMAArch, atomic_numbers  # noqa: E402
from pdb2reaction_b200.engine import UmabEngine            # noqa: E402

arch = UMAArch(num_experts=4)
elem, imgs = synth.make_config("C4")
z = atomic_numbers(elem)
merged = W.merge_mole(W.init_uma_weights(arch, 0), arch, z, 0, 1, "omol")
pos = torch.from_numpy(imgs.astype(np.float32)).cuda()
GB = 1 << 30
out = {}


def timeit(fn, reps=4):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


one = UmabEngine(merged, z, arch, workspace_bytes=28 * GB, store_bytes=75 * GB)
out["one_engine_32_images_ms"] = timeit(lambda: one.energy_forces(pos))
one.close()
torch.cuda.empty_cache()

engs = [UmabEngine(merged, z, arch, workspace_bytes=12 * GB, store_bytes=36 * GB) for _ in range(2)]
streams = [torch.cuda.Stream() for _ in range(2)]
halves = [pos[:16].contiguous(), pos[16:].contiguous()]
out["half_alone_16_images_ms"] = timeit(lambda: engs[0].energy_forces(halves[0]))


def both():
    def run(k):
        with torch.cuda.stream(streams[k]):
            engs[k].energy_forces(halves[k])
    ts = [threading.Thread(target=run, args=(k,)) for k in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()


out["two_engines_concurrent_2x16_images_ms"] = timeit(both)
print(json.dumps(out), flush=True)
