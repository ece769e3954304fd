"""GPU: energy+force throughput of all five BASELINE.json configs through the public calculator API
(host buffers in / out, the `e2e` path of bench.py), one JSON line per config."""
import sys, os, time, json, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pdb2reaction_b200 import uma_pysis, synth
from pdb2reaction_b200.shims import ANG2BOHR
warnings.simplefilter("ignore")
names = sys.argv[1:] or ["C1", "C2", "C3", "C4", "C5"]
for name in names:
    elem, imgs = synth.make_config(name)
    b, n = imgs.shape[0], imgs.shape[1]
    calc = uma_pysis(model="random:uma-s-1p1")
    c = imgs.reshape(b, -1) * ANG2BOHR
    calc.get_forces_batch(elem, c)                          # engine build + warm-up
    torch.cuda.synchronize()
    reps = 3 if n <= 1500 else 2
    t0 = time.perf_counter()
    for _ in range(reps):
        r = calc.get_forces_batch(elem, c)
    dt = (time.perf_counter() - t0) / reps
    eng = calc._core.backend.engines[0]
    # single-image latency (the reference's calling pattern: one image per call)
    calc.get_forces(elem, c[0])                             # warm-up at the single-image shapes (tensor maps)
    t0 = time.perf_counter()
    for _ in range(3):
        calc.get_forces(elem, c[0])
    lat = (time.perf_counter() - t0) / 3
    print(json.dumps({"config": name, "n_atoms": n, "n_images": b, "edges_per_step": eng.last_call_edges if b == 1 else None,
                      "image_evals_per_s": b / dt, "atoms_per_s": b * n / dt, "ms_per_batch": 1e3 * dt,
                      "single_image_latency_ms": 1e3 * lat, "finite": bool(np.isfinite(r["forces"]).all()),
                      "device_bytes": eng.stats()["device_bytes"]}), flush=True)
    calc = None
    from pdb2reaction_b200 import calculator as cm
    for e in cm._engine_cache.values():
        e.close()
    cm._engine_cache.clear()
    torch.cuda.empty_cache()
