"""GPU: latency of the serial calling pattern (one geometry per get_forces call, as pysisyphus optimizers do) and of a
batched string step for the small / medium BASELINE configs, with the sync-free + CUDA-graph path on and off."""
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdb2reaction_b200 import synth, uma_pysis, calculator as cm  # noqa: E402
from pdb2reaction_b200.shims import ANG2BOHR  # noqa: E402

warnings.simplefilter("ignore")
for name in sys.argv[1:] or ["C1", "C2", "C3"]:
    elem, imgs = synth.make_config(name)
    b, n = imgs.shape[0], imgs.shape[1]
    c = imgs.reshape(b, -1) * ANG2BOHR
    row = {"config": name, "n_atoms": n, "n_images": b}
    for mode in ("fast", "sync"):
        calc = uma_pysis(model="random:uma-s-1p1")
        calc._ensure_core(elem)
        eng = calc._core.backend.engines[0]
        if mode == "sync":
            eng.set_option("nosync", 0)
            eng.set_option("cuda_graphs", 0)
        rng = np.random.default_rng(0)
        for _ in range(4):
            calc.get_forces(elem, c[0] + 1e-3 * rng.normal(size=c[0].shape))
        ts = []
        for _ in range(20):
            x = c[0] + 1e-3 * rng.normal(size=c[0].shape)
            t0 = time.perf_counter()
            calc.get_forces(elem, x)
            ts.append(1e3 * (time.perf_counter() - t0))
        row[f"single_image_ms_{mode}"] = float(np.median(ts))
        for _ in range(4):
            calc.get_forces_batch(elem, c + 1e-3 * rng.normal(size=c.shape))
        ts = []
        for _ in range(10):
            x = c + 1e-3 * rng.normal(size=c.shape)
            t0 = time.perf_counter()
            calc.get_forces_batch(elem, x)
            ts.append(1e3 * (time.perf_counter() - t0))
        row[f"batch_ms_{mode}"] = float(np.median(ts))
        row[f"graph_replays_{mode}"] = eng.get_option("graph_replays")
        row[f"launches_per_single_call_{mode}"] = None
        l0 = eng.stats()["kernel_launches"]
        calc.get_forces(elem, c[0])
        row[f"launches_per_single_call_{mode}"] = eng.stats()["kernel_launches"] - l0
        for e in cm._engine_cache.values():
            e.close()
        cm._engine_cache.clear()
        torch.cuda.empty_cache()
    print(json.dumps(row), flush=True)
