"""Kernel families of one analytic-Hessian column batch (C3: 500 atoms), with the value sharing on and off.
usage: python tools/gpu_jvp_profile.py [n_atoms] [n_cols]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pdb2reaction_b200 import synth, uma_pysis
from pdb2reaction_b200.shims import ANG2BOHR

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
k = int(sys.argv[2]) if len(sys.argv) > 2 else 40
elem, coords = synth.make_cluster(n, 3)
calc = uma_pysis(model="random:uma-s-1p1", device="cuda:0", hessian_calc_mode="Analytical")
calc.get_forces(elem, (coords * ANG2BOHR).reshape(-1))
eng = calc._core.backend.engines[0]
pos = torch.from_numpy(coords.astype(np.float32)).cuda().unsqueeze(0).expand(k, n, 3).contiguous()
tan = torch.zeros(k, 3 * n, device="cuda"); tan[torch.arange(k), torch.arange(k)] = 1.0
tan = tan.view(k, n, 3)
out = {}
for mode in ("plain", "shared"):
    eng.set_option("jvp_shared_base", 1 if mode == "shared" else 0)
    eng.forces_jvp(pos, tan); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): eng.forces_jvp(pos, tan)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    eng.profile(True); eng.forces_jvp(pos, tan); torch.cuda.synchronize(); fam = eng.profile_read(); eng.profile(False)
    out[mode] = {"ms_per_batch": ms, "ms_per_column": ms / k, "families_ms": {f: round(v["ms"], 2) for f, v in fam.items() if v["ms"] > 0.5}}
    print(mode, f"{ms:.1f} ms per batch of {k} columns ({ms / k:.2f} ms per column)", out[mode]["families_ms"])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/jvp_profile.json", "w"), indent=1)
