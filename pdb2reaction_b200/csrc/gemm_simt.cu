// fp32 SIMT GEMM  C[M,N] = A[M,K] . W[N,K]^T (+ bias) (+ C), batched over grid.z.
//
// This is the exact-fp32 path: it serves the node-level (small-M) contractions, is the
// numerical ground truth the tensor-core path (gemm_tc.cu) is validated against on the GPU, and
// the path used when UMAB_GEMM=simt.  128x128x16 CTA tile, 256 threads, 8x8 register tile per
// thread, operands staged transposed in shared memory with register double buffering.
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

}  // namespace

void gemm_simt(const GemmArgs& a, cudaStream_t st) {
    if (a.M <= 0) return;
    if (a.K % BK != 0 || a.N % 4 != 0 || a.lda % 4 != 0 || a.ldw % 4 != 0 || a.ldc % 4 != 0 || a.batch > 9)
        throw CudaError("gemm_simt: unsupported shape (K%16, N%4, ld%4, batch<=9 required)");
    dim3 grid((unsigned)((a.M + BM - 1) / BM), (unsigned)((a.N + BN - 1) / BN), (unsigned)a.batch);
    gemm_simt_kernel<<<grid, NT, 0, st>>>(a);
    UMAB_LAUNCH_CHECK();
}

}  // namespace umab
