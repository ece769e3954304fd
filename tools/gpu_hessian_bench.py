"""GPU: Hessian wall-time of BASELINE.json configs[2] (freq/tsopt full Hessian of a ~500-atom
cluster) through the public calculator API: FiniteDifference mode = 1 + 2*3N batched force
evaluations (the reference runs them one predict() at a time, uma_pysis.py:652-675)."""
import sys, os, time, json, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pdb2reaction_b200 import uma_pysis, synth
from pdb2reaction_b200.shims import ANG2BOHR
warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
workers = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mode = sys.argv[3] if len(sys.argv) > 3 else "FiniteDifference"
elem, coords = synth.make_cluster(n, 3)
calc = uma_pysis(model="random:uma-s-1p1", workers=workers, hessian_calc_mode=mode)
calc.get_forces(elem, coords * ANG2BOHR)                 # engine build + warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
r = calc.get_hessian(elem, coords * ANG2BOHR)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
h = r["hessian"]
sym = float((h - h.T).abs().max())
w = torch.linalg.eigvalsh(h)
work = f"{1 + 6 * n} force evaluations" if mode.lower().startswith("f") else f"{3 * n} dual-number (value+tangent) passes"
print(json.dumps({"config": f"C3: full {mode} Hessian, {n} atoms, {3 * n} columns, {work}",
                  "workers": workers, "hessian_wall_s": dt, "columns_per_s": 3 * n / dt,
                  "shape": list(h.shape), "dtype": str(h.dtype), "asym": sym,
                  "n_near_zero_modes(|w|<1e-4 au)": int((w.abs() < 1e-4).sum())}))
