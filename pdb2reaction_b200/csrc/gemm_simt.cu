// fp32 SIMT GEMM  C[M,N] = A[M,K] . W[N,K]^T (+ bias) (+ C), batched over grid.z.
//
// This is the exact-fp32 path: it serves the node-level (small-M) contractions, is the
// numerical ground truth the tensor-core path (gemm_tc.cu) is validated against on the GPU, and
// the path used when UMAB_GEMM=simt.  128x128x16 CTA tile, 256 threads, 8x8 register tile per
// thread, operands staged transposed in shared memory with register double buffering.
#include <cooperative_groups.h>

#include "common.cuh"

namespace umab {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

// operand rounding of the precision study (GemmArgs::round_mode); never active on the product path
__device__ __forceinline__ float rnd_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ float rnd_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float rnd_bf16x2(float x) { const float h = rnd_bf16(x); return h + rnd_bf16(x - h); }
__device__ __forceinline__ float rnd_a(float x, int mode) {
    return mode == 1 ? rnd_tf32(x) : mode == 2 ? rnd_bf16x2(x) : (mode == 3 || mode == 4) ? rnd_bf16(x) : x;
}
__device__ __forceinline__ float rnd_w(float x, int mode) {
    return mode == 1 ? rnd_tf32(x) : mode == 3 ? rnd_bf16x2(x) : (mode == 2 || mode == 4) ? rnd_bf16(x) : x;
}

__global__ void __launch_bounds__(NT, 2)
gemm_simt_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int bz = blockIdx.z;
    const float* __restrict__ A = g.A + (long long)bz * g.strideA;
    const float* __restrict__ W = g.W + (long long)g.wsel[bz] * g.strideW;
    float* Cm = g.Cmat + (long long)bz * g.strideC;
    const int M = g.M, N = g.N, K = g.K;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;          // 16 x 16 threads, each 8 x 8 outputs

    // global -> register staging: each thread loads 2 float4 of A and 2 of W per k-tile
    // tile is [128 rows][16 k] = 512 float4; thread t loads float4 #t and #t+256
    const int lrow0 = tid / 4, lk = (tid % 4) * 4;   // rows lrow0 and lrow0+64
    float4 ra[2], rb[2];

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            long long r = m0 + lrow0 + i * 64;
            ra[i] = (r < M) ? ld4(A + r * g.lda + k0 + lk) : f4zero();
            int n = n0 + lrow0 + i * 64;
            rb[i] = (n < N) ? ld4(W + (long long)n * g.ldw + k0 + lk) : f4zero();
        }
        if (g.round_mode) {
            const int rm = g.round_mode;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                ra[i] = make_float4(rnd_a(ra[i].x, rm), rnd_a(ra[i].y, rm), rnd_a(ra[i].z, rm), rnd_a(ra[i].w, rm));
                rb[i] = make_float4(rnd_w(rb[i].x, rm), rnd_w(rb[i].y, rm), rnd_w(rb[i].z, rm), rnd_w(rb[i].w, rm));
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int r = lrow0 + i * 64;
            As[buf][lk + 0][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y;
            As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w;
            Bs[buf][lk + 0][r] = rb[i].x; Bs[buf][lk + 1][r] = rb[i].y;
            Bs[buf][lk + 2][r] = rb[i].z; Bs[buf][lk + 3][r] = rb[i].w;
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    const int nk = K / BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], b[8];
            float4 t;
            t = ld4(&As[buf][k][ty * 4]);        a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
            t = ld4(&As[buf][k][64 + ty * 4]);   a[4] = t.x; a[5] = t.y; a[6] = t.z; a[7] = t.w;
            t = ld4(&Bs[buf][k][tx * 4]);        b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w;
            t = ld4(&Bs[buf][k][64 + tx * 4]);   b[4] = t.x; b[5] = t.y; b[6] = t.z; b[7] = t.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    const bool use_bias = g.bias != nullptr && (!g.bias_first_batch_only || bz == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        long long r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            int n = n0 + jh * 64 + tx * 4;
            if (n >= N) continue;             // N is a multiple of 4 (checked on the host)
            float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
            if (use_bias) v = f4add(v, ld4(g.bias + n));
            float* p = Cm + r * g.ldc + n;
            if (g.accumulate) v = f4add(v, ld4(p));
            st4(p, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Latency variant for small molecules (images below 100 atoms: M = a few hundred edges, grids of 1-15 CTAs of the
// kernel above, each walking the whole K = 768 / 1024 loop: 44 us per launch, 85 % of a 20-atom evaluation).
// 64 x 64 tiles and SPLIT-K over a thread-block cluster: the S CTAs of a cluster (cluster dims 1 x 1 x S) each
// accumulate one K slice of the same output tile, park the partial tile in their shared memory, and after a cluster
// barrier every CTA sums its share of the rows over the S partials through distributed shared memory IN RANK ORDER
// (fixed summation order: deterministic, independent of M and of the tile position).  S is a function of K only.
namespace cg = cooperative_groups;
constexpr int SBM = 64, SBN = 64;

__global__ void __launch_bounds__(NT, 2)
gemm_simt_splitk_kernel(GemmArgs g, int S) {
    __shared__ __align__(16) float As[2][BK][SBM + 4];
    __shared__ __align__(16) float Bs[2][BK][SBN + 4];
    __shared__ __align__(16) float part[SBM][SBN + 4];

    cg::cluster_group cluster = cg::this_cluster();
    const int ks = (int)cluster.block_rank();                  // K slice of this CTA (cluster spans grid.z)
    const int bz = blockIdx.z / S;
    const float* __restrict__ A = g.A + (long long)bz * g.strideA;
    const float* __restrict__ W = g.W + (long long)g.wsel[bz] * g.strideW;
    float* Cm = g.Cmat + (long long)bz * g.strideC;
    const int M = g.M, N = g.N;
    const int kper = g.K / S;
    const int k_begin = ks * kper;
    const long long m0 = (long long)blockIdx.x * SBM;
    const int n0 = blockIdx.y * SBN;
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;          // 16 x 16 threads, each 4 x 4 outputs
    const int lrow = tid / 4, lk = (tid % 4) * 4;    // one float4 of A and one of W per thread and k tile
    float4 ra, rb;

    auto load_tiles = [&](int k0) {
        const long long r = m0 + lrow;
        ra = (r < M) ? ld4(A + r * g.lda + k0 + lk) : f4zero();
        const int n = n0 + lrow;
        rb = (n < N) ? ld4(W + (long long)n * g.ldw + k0 + lk) : f4zero();
        if (g.round_mode) {
            const int rm = g.round_mode;
            ra = make_float4(rnd_a(ra.x, rm), rnd_a(ra.y, rm), rnd_a(ra.z, rm), rnd_a(ra.w, rm));
            rb = make_float4(rnd_w(rb.x, rm), rnd_w(rb.y, rm), rnd_w(rb.z, rm), rnd_w(rb.w, rm));
        }
    };
    auto store_tiles = [&](int buf) {
        As[buf][lk + 0][lrow] = ra.x; As[buf][lk + 1][lrow] = ra.y; As[buf][lk + 2][lrow] = ra.z; As[buf][lk + 3][lrow] = ra.w;
        Bs[buf][lk + 0][lrow] = rb.x; Bs[buf][lk + 1][lrow] = rb.y; Bs[buf][lk + 2][lrow] = rb.z; Bs[buf][lk + 3][lrow] = rb.w;
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    load_tiles(k_begin);
    store_tiles(0);
    __syncthreads();
    const int nk = kper / BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles(k_begin + (kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a4 = ld4(&As[buf][k][ty * 4]);
            const float4 b4 = ld4(&Bs[buf][k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) st4(&part[ty * 4 + i][tx * 4], make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    cluster.sync();                                   // all S partial tiles are in shared memory

    // CTA `ks` finishes rows [ks * 64 / S, (ks + 1) * 64 / S) of the tile: sum over the partials in rank order
    const bool use_bias = g.bias != nullptr && (!g.bias_first_batch_only || bz == 0);
    const int rows_per = SBM / S;
    for (int idx = tid; idx < rows_per * (SBN / 4); idx += NT) {
        const int rl = ks * rows_per + idx / (SBN / 4);
        const int c4 = (idx % (SBN / 4)) * 4;
        const long long r = m0 + rl;
        const int n = n0 + c4;
        float4 v = f4zero();
        for (int p = 0; p < S; ++p) {
            const float* remote = cluster.map_shared_rank(&part[rl][c4], p);
            const float4 t = *reinterpret_cast<const float4*>(remote);
            v = make_float4(v.x + t.x, v.y + t.y, v.z + t.z, v.w + t.w);
        }
        if (r < M && n < N) {                         // N is a multiple of 4 (checked on the host)
            if (use_bias) v = f4add(v, ld4(g.bias + n));
            float* pc = Cm + r * g.ldc + n;
            if (g.accumulate) v = f4add(v, ld4(pc));
            st4(pc, v);
        }
    }
    cluster.sync();                                   // nobody leaves while its partial tile can still be read
}

}  // namespace

// K-slices of the split-K variant: a function of K only, so that a row's arithmetic never depends on the batch
int gemm_simt_splitk_slices(int K) {
    if (K >= 1024 && K % 128 == 0) return 8;
    if (K >= 384 && K % 64 == 0) return 4;
    if (K >= 128 && K % 32 == 0) return 2;
    return 1;
}

void gemm_simt_splitk(const GemmArgs& a, cudaStream_t st) {
    if (a.M <= 0) return;
    if (a.K % BK != 0 || a.N % 4 != 0 || a.lda % 4 != 0 || a.ldw % 4 != 0 || a.ldc % 4 != 0 || a.batch > 9)
        throw CudaError("gemm_simt_splitk: unsupported shape (K%16, N%4, ld%4, batch<=9 required)");
    const int S = gemm_simt_splitk_slices(a.K);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((a.M + SBM - 1) / SBM), (unsigned)((a.N + SBN - 1) / SBN), (unsigned)(a.batch * S));
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = (unsigned)S;
    cfg.attrs = at; cfg.numAttrs = 1;
    UMAB_CUDA(cudaLaunchKernelEx(&cfg, gemm_simt_splitk_kernel, a, S));
    UMAB_LAUNCH_CHECK();
}

void gemm_simt(const GemmArgs& a, cudaStream_t st) {
    if (a.M <= 0) return;
    if (a.K % BK != 0 || a.N % 4 != 0 || a.lda % 4 != 0 || a.ldw % 4 != 0 || a.ldc % 4 != 0 || a.batch > 9)
        throw CudaError("gemm_simt: unsupported shape (K%16, N%4, ld%4, batch<=9 required)");
    dim3 grid((unsigned)((a.M + BM - 1) / BM), (unsigned)((a.N + BN - 1) / BN), (unsigned)a.batch);
    gemm_simt_kernel<<<grid, NT, 0, st>>>(a);
    UMAB_LAUNCH_CHECK();
}

}  // namespace umab
