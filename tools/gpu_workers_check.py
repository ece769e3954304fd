"""GPU (>= 2 devices): uma_pysis(model="random:uma-s-1p1", workers=2) shards a batch over two GPUs from one process and
returns the same bits as workers=1."""
import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pdb2reaction_b200 import uma_pysis, synth, calculator as cm
from pdb2reaction_b200.arch import UMAArch
from pdb2reaction_b200.shims import ANG2BOHR
warnings.simplefilter("ignore")
arch = UMAArch(num_experts=4)
orig = cm.CudaBackend.__init__
cm.CudaBackend.__init__ = lambda self, e, **kw: orig(self, e, **{**kw, "arch": arch})
elem, imgs = synth.make_string(300, 7, 2)
c = imgs.reshape(7, -1) * ANG2BOHR
r1 = uma_pysis(model="random:uma-s-1p1", workers=1).get_forces_batch(elem, c)
r2 = uma_pysis(model="random:uma-s-1p1", workers=2).get_forces_batch(elem, c)
print("workers=2 equals workers=1:", np.array_equal(r1["energy"], r2["energy"]), np.array_equal(r1["forces"], r2["forces"]))
h = uma_pysis(model="random:uma-s-1p1", workers=2, freeze_atoms=list(range(290))).get_hessian(elem, c[0])["hessian"]
h1 = uma_pysis(model="random:uma-s-1p1", workers=1, freeze_atoms=list(range(290))).get_hessian(elem, c[0])["hessian"]
print("hessian", tuple(h.shape), h.dtype, h.device, float(h.abs().max()), "workers=2 equals workers=1:", torch.equal(h, h1.to(h.device)))
ha = uma_pysis(model="random:uma-s-1p1", workers=2, freeze_atoms=list(range(290)), hessian_calc_mode="Analytical").get_hessian(elem, c[0])["hessian"]
print("analytic vs FD (workers=2): max abs diff", float((ha - h).abs().max()), "of", float(h.abs().max()))
ha1 = uma_pysis(model="random:uma-s-1p1", workers=1, freeze_atoms=list(range(290)), hessian_calc_mode="Analytical").get_hessian(elem, c[0])["hessian"]
print("analytic workers=2 equals workers=1:", torch.equal(ha, ha1.to(ha.device)))
