import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdb2reaction_b200 import synth, weights as W
from pdb2reaction_b200.arch import UMAArch, atomic_numbers
from pdb2reaction_b200.engine import UmabEngine
arch = UMAArch(num_experts=4); sd = W.init_uma_weights(arch, 0)
elem, imgs = synth.make_string(60, 3, 11)
z = atomic_numbers(elem); merged = W.merge_mole(sd, arch, z, 0, 1, "omol")
pos = imgs.astype(np.float32)
out = {}
for tag, kw in (("single", {}), ("multi", {"workspace_bytes": 9600 * 4 * 1500})):
    eng = UmabEngine(merged, z, arch, debug=True, **kw)
    e, f = eng.energy_forces_host(pos)
    out[tag + "_gn1"] = eng.debug_tensor("l3.g_n1").numpy()
    out[tag + "_ga0"] = eng.debug_tensor("bwd.l3.ga0").numpy().reshape(-1, 768)[:, :256].copy()
    out[tag + "_rad"] = eng.debug_tensor("bwd.l3.rad").numpy().reshape(-1, 1536)[:, :256].copy()
    out["ei"] = eng.graph(torch.from_numpy(pos).cuda()).numpy()
np.savez_compressed("gpurun_out/chunk3.npz", **out)
