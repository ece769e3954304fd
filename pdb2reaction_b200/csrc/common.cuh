// Shared device/host helpers for the umab sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace umab {

constexpr int C = 128;        // sphere channels
constexpr int H = 128;        // hidden channels
constexpr int NB = 64;        // Gaussian distance basis
constexpr int WIG = 36;       // per-edge Wigner record: D1 (9, row-major) | D2 (25, row-major) | 2 pad
constexpr int RAD1 = 12 * C;  // conv1 radial outputs: 3*2C + 2*2C + 1*2C = 1536

// m-primary row k  ->  l-primary coefficient index
__host__ __device__ constexpr int to_m(int k) {
    constexpr int t[9] = {0, 2, 6, 3, 7, 1, 5, 8, 4};
    return t[k];
}

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void check_cuda(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
        throw CudaError(buf);
    }
}
#define UMAB_CUDA(x) ::umab::check_cuda((x), #x, __FILE__, __LINE__)
extern std::atomic<long long> g_launch_count;   // kernels launched by this library (defined in engine.cu)
#define UMAB_LAUNCH_CHECK() (++::umab::g_launch_count, ::umab::check_cuda(cudaGetLastError(), "kernel launch", __FILE__, __LINE__))

// denominator in [1, inf): the fast reciprocal (MUFU.RCP, <= 2 ulp) needs none of the IEEE division's range fix-ups
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }
__device__ __forceinline__ float dsiluf_(float x) {
    float s = sigmoidf_(x);
    return s * (1.0f + x * (1.0f - s));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
// Blackwell packed fp32 (FFMA2 / FMUL2 / FADD2: two IEEE fp32 operations per instruction, same results as the scalar
// forms) halves the issue slots of the Wigner rotations in the latency-bound edge kernels.  UMAB_NO_F32X2 restores scalars.
#if !defined(UMAB_NO_F32X2)
__device__ __forceinline__ float4 f4mul(float4 a, float4 b) {
    const float2 lo = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    const float2 hi = __fmul2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 f4add(float4 a, float4 b) {
    const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    const float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
#else
__device__ __forceinline__ float4 f4mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
#endif
__device__ __forceinline__ float4 f4sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 f4scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ void f4fma(float4& acc, float s, float4 b) {
#if !defined(UMAB_NO_F32X2)
    const float2 s2 = make_float2(s, s);
    const float2 lo = __ffma2_rn(s2, make_float2(b.x, b.y), make_float2(acc.x, acc.y));
    const float2 hi = __ffma2_rn(s2, make_float2(b.z, b.w), make_float2(acc.z, acc.w));
    acc = make_float4(lo.x, lo.y, hi.x, hi.y);
#else
    acc.x = fmaf(s, b.x, acc.x); acc.y = fmaf(s, b.y, acc.y); acc.z = fmaf(s, b.z, acc.z); acc.w = fmaf(s, b.w, acc.w);
#endif
}
__device__ __forceinline__ float f4dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float f4hsum(float4 a) { return (a.x + a.y) + (a.z + a.w); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): the operand format of the bf16x3 tensor-core GEMMs
__device__ __forceinline__ void split4(float4 x, uint2& hi, uint2& lo) {
    // packed conversions (F2FP.BF16.F32.PACK_AB: two floats -> one bf16x2 register) and bit-level widening of hi
    hi.x = pack_bf16(x.x, x.y);
    hi.y = pack_bf16(x.z, x.w);
    const float h0 = __uint_as_float(hi.x << 16), h1 = __uint_as_float(hi.x & 0xffff0000u);
    const float h2 = __uint_as_float(hi.y << 16), h3 = __uint_as_float(hi.y & 0xffff0000u);
    lo.x = pack_bf16(x.x - h0, x.y - h1);
    lo.y = pack_bf16(x.z - h2, x.w - h3);
}

// ---------------------------------------------------------------- GEMM (C = A W^T [+bias] [+C])
struct GemmArgs {
    const float* A = nullptr; long long lda = 0; long long strideA = 0;
    // pre-split activation operand (bf16 hi / lo planes, row pitch K) written by the producing kernel:
    // when set, the tensor-core GEMM loads it by TMA and `A` is ignored
    const __nv_bfloat16* A_hi = nullptr; const __nv_bfloat16* A_lo = nullptr;
    const float* W = nullptr; long long ldw = 0; long long strideW = 0;   // W is [N, K] row-major
    float* Cmat = nullptr;    long long ldc = 0; long long strideC = 0;
    const float* bias = nullptr;    // [N] or null
    int bias_first_batch_only = 0;
    int M = 0, N = 0, K = 0;
    int batch = 1;
    int wsel[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // weight index per batch entry
    int accumulate = 0;
    // precision study only (tools/gpu_precision_study.py): operand rounding emulated in the fp32 SIMT kernel.
    // 0 exact fp32; 1 single-pass TF32 (both operands rounded to 10 mantissa bits); 2 bf16 x2 with the weights
    // rounded to bf16 (a_hi w_hi + a_lo w_hi); 3 bf16 x2 with the activations rounded to bf16 (a_hi w_hi + a_hi w_lo);
    // 4 single-pass bf16
    int round_mode = 0;
    // fused gate epilogue (CTA-pair tensor-core GEMM only; conv-1 m = +-1 / +-2 GEMMs of the forward): besides C the
    // kernel writes  C[r, n] * gate[r, gcol(n)]  as bf16 hi / lo planes (the A operand of the conv-2 GEMM), where
    // gate = sigmoid of the gate pre-activations [M, gate_ld] written by gate_b0_kernel and
    // gcol(n) = ((n >> 7) & 1) * 128 + (n & 127)  (gate_mode 1: m = +-1 rows alternate l = 1, 2)  or
    //           128 + (n & 127)                    (gate_mode 2: m = +-2 rows, l = 2 only)
    const float* gate = nullptr; int gate_ld = 0; int gate_mode = 0;
    __nv_bfloat16* out_hi = nullptr; __nv_bfloat16* out_lo = nullptr;
};

void gemm_simt(const GemmArgs& a, cudaStream_t st);
// latency variant for small molecules: 64 x 64 tiles, split-K over a thread-block cluster (gemm_simt.cu)
void gemm_simt_splitk(const GemmArgs& a, cudaStream_t st);

}  // namespace umab
