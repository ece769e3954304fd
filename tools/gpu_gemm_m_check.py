import sys, os, torch
sys.path.insert(0, os.getcwd())
from pdb2reaction_b200 import engine
g = torch.Generator(device="cuda").manual_seed(0)
for (m, n, k) in [(3540, 768, 640), (3540, 1024, 512), (3540, 512, 256), (3540, 384, 384), (3540, 512, 512), (3540, 256, 256), (3540, 640, 768), (3540, 128, 1536), (3540, 1536, 128)]:
    a = torch.randn(m, k, device="cuda", generator=g); w = torch.randn(n, k, device="cuda", generator=g) / k ** 0.5
    full = engine.gemm(a, w, None, mode=7)
    ok = True
    for lo, hi in ((0, 1180), (1180, 2360), (2360, 3540)):
        part = engine.gemm(a[lo:hi].contiguous(), w, None, mode=7)
        d = (part - full[lo:hi]).abs()
        if d.max() > 0:
            ok = False
            cols = torch.nonzero(d.amax(0) > 0).flatten()
            print((m, n, k), (lo, hi), "max diff", float(d.max()), "n", int((d > 0).sum()), "cols", int(cols.min()), int(cols.max()))
    print((m, n, k), "M-independent:", ok)
