"""GPU: CTA-pair (cta_group::2) TMA-fed GEMM vs the single-CTA kernel: accuracy against fp64 and kernel-only throughput."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pdb2reaction_b200 import engine

# tails: M not a multiple of 256 / 128, a single tile, fewer tiles than SM pairs
SHAPES = [(300, 128, 64), (129, 256, 128), (4097, 256, 128), (4096, 128, 64), (8192, 1536, 128), (8192, 640, 768),
          (16384, 512, 512), (16384, 256, 256), (8192, 384, 384), (40000, 512, 1024), (8192, 768, 640),
          (8192, 128, 1536), (8192, 64, 128), (8192, 160, 64), (8192, 192, 192), (70001, 1024, 512)]


def main():
    torch.manual_seed(0)
    bad = 0
    for (m, n, k) in SHAPES:
        a = torch.randn(m, k, device="cuda")
        w = torch.randn(n, k, device="cuda") / k ** 0.5
        b = torch.randn(n, device="cuda")
        ref = a.double() @ w.double().T + b.double()
        out = {}
        for mode, name in ((3, "tc2/64"), (5, "pair/64")):
            c = engine.gemm(a, w, b, mode=mode)
            torch.cuda.synchronize()
            out[name] = (c.double() - ref).abs().max().item() / ref.abs().max().item()
        c3 = engine.gemm(a, w, b, mode=3)
        c5 = engine.gemm(a, w, b, mode=5)
        same = torch.equal(c3, c5)
        bad += any(not (v < 2e-5) for v in out.values())
        print(f"M={m:6d} N={n:5d} K={k:5d}  rel err " + "  ".join(f"{k_} {v:.2e}" for k_, v in out.items())
              + f"  bitwise==single-CTA: {same}", flush=True)
    print("ACCURACY", "FAIL" if bad else "OK", flush=True)
    E = 1 << 20
    shapes = [(E, 128, 64), (E, 128, 128), (E, 1536, 128), (E, 640, 768), (E, 512, 1024), (E, 256, 512),
              (E, 384, 384), (E, 512, 512), (E, 256, 256), (E, 768, 640), (E, 1024, 512), (E, 512, 256),
              (E, 128, 1536), (E, 384, 128), (E, 128, 384)]
    tot = {}
    for (m, n, k) in shapes:
        a = torch.randn(m, k, device="cuda"); w = torch.randn(n, k, device="cuda") / k ** 0.5
        line = f"M={m} N={n} K={k}:"
        for mode, name in ((3, "tc2/64"), (5, "pair/64")):
            ms = engine.gemm_bench(a, w, mode, iters=10)
            tot[name] = tot.get(name, 0.0) + ms
            line += f"  {name} {ms:.3f} ms {2.0 * m * n * k / ms / 1e9:6.1f} TF/s"
        print(line, flush=True)
        del a, w
    print("sum of shapes (ms): " + "  ".join(f"{k_} {v:.2f}" for k_, v in tot.items()), flush=True)


if __name__ == "__main__":
    main()
