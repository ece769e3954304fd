import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from pdb2reaction_b200 import synth, weights as W
from pdb2reaction_b200.arch import UMAArch, atomic_numbers
from pdb2reaction_b200.engine import UmabEngine
arch = UMAArch(num_experts=4); sd = W.init_uma_weights(arch, 0)
n, b = 300, 4
elem, imgs = synth.make_string(n, b, 50 + n)
z = atomic_numbers(elem); merged = W.merge_mole(sd, arch, z, 0, 1, "omol")
fast = UmabEngine(merged, z, arch); slow = UmabEngine(merged, z, arch)
slow.set_option("nosync", 0); slow.set_option("cuda_graphs", 0)
if len(sys.argv) > 1: fast.set_option("cuda_graphs", 0)
rng = np.random.default_rng(n)
def ctr(): return {k: fast.get_option(k) for k in ("graph_captures", "graph_replays", "overflow_retries", "edges_per_image_seen", "fast_calls")}
for k in range(6):
    pos = (imgs + 0.03 * k * rng.normal(size=imgs.shape)).astype(np.float32)
    e1, f1 = fast.energy_forces_host(pos); c1 = ctr(); ne1 = fast.last_call_edges
    e0, f0 = slow.energy_forces_host(pos)
    en1, _ = fast.energy_forces_host(pos, forces=False); c2 = ctr(); ne2 = fast.last_call_edges
    en0, _ = slow.energy_forces_host(pos, forces=False)
    print(k, "EF eq", np.array_equal(e0, e1), np.array_equal(f0, f1), "E-only eq fast/slowEF", np.array_equal(en1, e0), "slowE/slowEF", np.array_equal(en0, e0),
          "edges", ne1, ne2, slow.last_call_edges, c1, c2, en1[:2], e0[:2], flush=True)
