"""-m gpu: the library's GEMM kernels against a float64 torch matmul on the shapes the model uses."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(1000, 128, 64), (4097, 1536, 128), (2500, 640, 768), (5000, 512, 512), (5000, 256, 256),
          (2500, 384, 384), (5000, 512, 256), (5000, 256, 128), (2500, 768, 640), (3000, 128, 1536), (777, 64, 128)]


@pytest.mark.parametrize("m,n,k", SHAPES)
def test_simt_gemm(m, n, k, built_lib):
    from pdb2reaction_b200 import engine
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) / k ** 0.5
    b = torch.randn(n, device="cuda", generator=g)
    c = engine.gemm(a, w, b, mode=0)
    ref = (a.double() @ w.double().T + b.double())
    assert (c.double() - ref).abs().max() < 2e-5 * ref.abs().max()


@pytest.mark.parametrize("m,n,k", [s for s in SHAPES if s[2] % 64 == 0 and s[1] % 32 == 0])
@pytest.mark.parametrize("accumulate_bias", [False, True])
def test_tensor_core_bf16x3_gemm(m, n, k, accumulate_bias, built_lib):
    """tcgen05 path: bf16 hi/lo split, 3 MMAs per product, fp32 accumulation in TMEM -> ~4e-6 relative."""
    from pdb2reaction_b200 import engine
    g = torch.Generator(device="cuda").manual_seed(m * 3 + n + k)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) / k ** 0.5
    b = torch.randn(n, device="cuda", generator=g) if accumulate_bias else None
    c = engine.gemm(a, w, b, mode=1)
    ref = a.double() @ w.double().T + (b.double() if b is not None else 0.0)
    assert (c.double() - ref).abs().max() < 2e-5 * ref.abs().max()


TC2_SHAPES = [s for s in SHAPES if s[2] % 64 == 0 and s[1] % 32 == 0] + [(5, 128, 128), (100, 256, 128), (129, 640, 768),
                                                                          (255, 1536, 128), (300001, 256, 256)]


@pytest.mark.parametrize("m,n,k", TC2_SHAPES)
@pytest.mark.parametrize("mode", [3, 5])
@pytest.mark.parametrize("with_bias", [False, True])
def test_tma_fed_tensor_core_gemm(m, n, k, mode, with_bias, built_lib):
    """gemm_tc2: both operands by TMA from bf16 hi/lo planes, TMA-store epilogue; mode 3 = single-CTA kernel,
    mode 5 = the CTA-pair (cta_group::2, 256-row tiles) kernel the engine uses by default.
    Covers row tails (M not a multiple of 128 / 256, M < 128) and every N tile width the model uses."""
    from pdb2reaction_b200 import engine
    g = torch.Generator(device="cuda").manual_seed(m * 5 + n + k)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) / k ** 0.5
    b = torch.randn(n, device="cuda", generator=g) if with_bias else None
    c = engine.gemm(a, w, b, mode=mode)
    ref = a.double() @ w.double().T + (b.double() if b is not None else 0.0)
    assert torch.isfinite(c).all()
    assert (c.double() - ref).abs().max() < 2e-5 * ref.abs().max()


@pytest.mark.parametrize("m,n,k", [(129, 256, 128), (70001, 1024, 512), (8192, 640, 768), (300, 128, 64)])
def test_cta_pair_gemm_is_bit_identical_to_single_cta(m, n, k, built_lib):
    """Same MMAs in the same k order per output element: the pair kernel changes which SM holds a tile, not the bits."""
    from pdb2reaction_b200 import engine
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) / k ** 0.5
    b = torch.randn(n, device="cuda", generator=g)
    assert torch.equal(engine.gemm(a, w, b, mode=3), engine.gemm(a, w, b, mode=5))


@pytest.mark.parametrize("m,n,k", [(368, 640, 768), (380, 512, 1024), (100, 128, 64), (777, 64, 128), (40, 384, 384),
                                   (3000, 128, 1536), (65, 1536, 128), (1, 128, 128)])
def test_simt_splitk_cluster_gemm(m, n, k, built_lib):
    """Latency variant for small molecules: 64 x 64 tiles, K split over a thread-block cluster, partial tiles summed
    through distributed shared memory in rank order.  fp32-exact-grade, and a row's bits never depend on M."""
    from pdb2reaction_b200 import engine
    g = torch.Generator(device="cuda").manual_seed(m + 2 * n + k)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) / k ** 0.5
    b = torch.randn(n, device="cuda", generator=g)
    c = engine.gemm(a, w, b, mode=7)
    ref = (a.double() @ w.double().T + b.double())
    assert (c.double() - ref).abs().max() < 2e-5 * ref.abs().max()
    assert torch.equal(engine.gemm(a, w, b, mode=7), c)                          # run-to-run
    big = torch.cat([torch.randn(131, k, device="cuda", generator=g), a, torch.randn(7, k, device="cuda", generator=g)])
    assert torch.equal(engine.gemm(big, w, b, mode=7)[131:131 + m], c)           # batch / tile position independent
