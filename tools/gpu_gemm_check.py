"""GPU: accuracy + throughput of the library GEMMs (SIMT fp32 vs tcgen05 bf16x3) on model shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pdb2reaction_b200 import engine

SHAPES = [(300, 128, 64), (4097, 256, 128), (4096, 128, 64), (8192, 1536, 128), (8192, 640, 768), (16384, 512, 512),
          (16384, 256, 256), (8192, 384, 384), (16384, 512, 256), (16384, 256, 128), (8192, 768, 640),
          (8192, 128, 1536), (8192, 64, 128), (8192, 160, 64), (8192, 192, 192)]

def main():
    torch.manual_seed(0)
    for (m, n, k) in SHAPES:
        a = torch.randn(m, k, device="cuda")
        w = torch.randn(n, k, device="cuda") / k ** 0.5
        b = torch.randn(n, device="cuda")
        ref = a.double() @ w.double().T + b.double()
        out = {}
        for mode, name in ((0, "simt"), (1, "tc"), (3, "tc2/64"), (5, "pair/64"), (7, "simt split-K")):
            if mode in (3, 5) and (k % 64 or n % 32):
                out[name] = float("nan")
                continue
            c = engine.gemm(a, w, b, mode=mode)
            torch.cuda.synchronize()
            err = (c.double() - ref).abs().max().item() / ref.abs().max().item()
            out[name] = err
        print(f"M={m:6d} N={n:5d} K={k:5d}  rel err " + "  ".join(f"{k_} {v:.2e}" for k_, v in out.items()), flush=True)
    # throughput at bench-like sizes (weights re-split each call in this test entry: time the kernel only via events
    # around a second call is not possible here, so this is an upper bound on time)
    E = 1 << 20
    # (M, N, K) of the pipeline's GEMMs (forward, then the adjoints) with E edges per chunk
    shapes = [(E, 128, 64), (E, 128, 128), (E, 1536, 128), (E, 640, 768), (E, 512, 1024), (E, 256, 512),
              (E, 384, 384), (E, 512, 512), (E, 256, 256), (E, 768, 640), (E, 1024, 512), (E, 512, 256),
              (E, 128, 1536), (E, 64, 128), (E, 384, 128), (E, 128, 384)]
    for (m, n, k) in shapes:
        a = torch.randn(m, k, device="cuda"); w = torch.randn(n, k, device="cuda") / k ** 0.5
        line = f"M={m} N={n} K={k}:"
        for mode, name in ((2, "tc"), (3, "tc2/64"), (5, "pair/64")):
            if mode in (3, 5) and (k % 64 or n % 32):
                continue
            ms = engine.gemm_bench(a, w, mode, iters=5)
            line += f"  {name} {ms:.3f} ms {2.0 * m * n * k / ms / 1e9:6.1f} TF/s"
        print(line, flush=True)
        del a, w

if __name__ == "__main__":
    main()
