"""GPU: single-image latency before / after a batched call on the same calculator (C2: 300 atoms)."""
import sys, os, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pdb2reaction_b200 import uma_pysis, synth
from pdb2reaction_b200.shims import ANG2BOHR
warnings.simplefilter("ignore")
elem, imgs = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "C2")
b = imgs.shape[0]
c = imgs.reshape(b, -1) * ANG2BOHR
calc = uma_pysis(model="random:uma-s-1p1")
def lat(tag, k=8):
    ts = []
    for i in range(k):
        t0 = time.perf_counter(); calc.get_forces(elem, c[i % b]); ts.append(1e3 * (time.perf_counter() - t0))
    print(tag, " ".join(f"{t:.2f}" for t in ts), flush=True)
lat("fresh single  ")
t0 = time.perf_counter(); calc.get_forces_batch(elem, c); print("batch", 1e3 * (time.perf_counter() - t0))
t0 = time.perf_counter(); calc.get_forces_batch(elem, c); print("batch", 1e3 * (time.perf_counter() - t0))
lat("after batch   ")
