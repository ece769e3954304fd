import os, sys
os.environ["UMAB_DEBUG_FAST"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdb2reaction_b200 import synth, weights as W
from pdb2reaction_b200.arch import UMAArch, atomic_numbers
from pdb2reaction_b200.engine import UmabEngine
arch = UMAArch(num_experts=4); sd = W.init_uma_weights(arch, 0)
n, b = 300, 4
elem, imgs = synth.make_string(n, b, 50 + n)
z = atomic_numbers(elem); merged = W.merge_mole(sd, arch, z, 0, 1, "omol")
fast = UmabEngine(merged, z, arch, debug=True); slow = UmabEngine(merged, z, arch, debug=True)
slow.set_option("nosync", 0); slow.set_option("cuda_graphs", 0); fast.set_option("cuda_graphs", 0)
rng = np.random.default_rng(n)
names = ["x0", "gauss", "env", "wig"] + [f"l{l}.{k}" for l in range(4) for k in ("n1", "rad", "y0", "y1", "y2", "z0", "x1", "x")] + ["node_e"]
per = {"gauss": 64, "env": 1, "wig": 36, "rad": 1536, "y0": 640, "y1": 512, "y2": 256, "z0": 384}
for k in range(3):
    pos = (imgs + 0.03 * k * rng.normal(size=imgs.shape)).astype(np.float32)
    for forces in (True, False):
        e1, _ = fast.energy_forces_host(pos, forces=forces)
        e0, _ = slow.energy_forces_host(pos, forces=forces)
        ne = slow.last_call_edges
        print(k, "forces" if forces else "E-only", "E eq", np.array_equal(e0, e1), "edges", ne, fast.graph_counts(), "fast_calls", fast.get_option("fast_calls"), flush=True)
        if not np.array_equal(e0, e1):
            for nm in names:
                a, bb = fast.debug_tensor(nm), slow.debug_tensor(nm)
                w = per.get(nm.split(".")[-1])
                if w:
                    a = a[: ne * w]; bb = bb[: ne * w]
                if a.numel() != bb.numel():
                    print("  ", nm, "size", a.numel(), bb.numel()); continue
                d = (a - bb).abs()
                bad = torch.nonzero(d > 0).flatten()
                print("  ", nm, "max diff", float(d.max()) if d.numel() else 0.0, "n diff", bad.numel(), "first", (int(bad[0]) // (w or 1152)) if bad.numel() else None, flush=True)
            sys.exit(0)
