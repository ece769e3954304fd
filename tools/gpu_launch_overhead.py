"""GPU: per-launch floor of the single-CTA and CTA-pair GEMM kernels (tiny M: launch + prologue + drain only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pdb2reaction_b200 import engine
for m in (256, 4096, 37888):
    for (n, k) in ((128, 128), (512, 512), (640, 768)):
        a = torch.randn(m, k, device="cuda"); w = torch.randn(n, k, device="cuda") / k ** 0.5
        line = f"M={m:6d} N={n:4d} K={k:4d}:"
        for mode, name in ((3, "single"), (5, "pair")):
            ms = engine.gemm_bench(a, w, mode, iters=200)
            line += f"  {name} {1e3 * ms:7.1f} us"
        print(line, flush=True)
