// Shared device/host helpers for the umab sm_100a kernels.
#pragma once
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

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
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
__device__ __forceinline__ float4 f4mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 f4scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ void f4fma(float4& acc, float s, float4 b) {
    acc.x = fmaf(s, b.x, acc.x); acc.y = fmaf(s, b.y, acc.y); acc.z = fmaf(s, b.z, acc.z); acc.w = fmaf(s, b.w, acc.w);
}
__device__ __forceinline__ float f4dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float f4hsum(float4 a) { return (a.x + a.y) + (a.z + a.w); }

// ---------------------------------------------------------------- GEMM (C = A W^T [+bias] [+C])
struct GemmArgs {
    const float* A = nullptr; long long lda = 0; long long strideA = 0;
    const float* W = nullptr; long long ldw = 0; long long strideW = 0;   // W is [N, K] row-major
    float* Cmat = nullptr;    long long ldc = 0; long long strideC = 0;
    const float* bias = nullptr;    // [N] or null
    int bias_first_batch_only = 0;
    int M = 0, N = 0, K = 0;
    int batch = 1;
    int wsel[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // weight index per batch entry
    int accumulate = 0;
};

void gemm_simt(const GemmArgs& a, cudaStream_t st);

// ---------------------------------------------------------------- kernel launchers (one per .cu)
void launch_neighbor_count(const float* pos, int n_img, int n_atoms, float cutoff, int cap, int* deg, float* thr, cudaStream_t st);
void launch_neighbor_fill(const float* pos, int n_img, int n_atoms, float cutoff, int cap, const float* thr,
                          const int* row_ptr, int* src, int* tgt, cudaStream_t st);
void launch_scan(const int* in, int* out, int n, cudaStream_t st);
void launch_source_csr(const int* src, int n_edges, int n_nodes, int* odeg, int* sptr, int* cursor, int* tmp,
                       int* sedge, cudaStream_t st);

void launch_geometry_fwd(const float* pos, const int* src, const int* tgt, int n_edges, float cutoff, float* vec,
                         float* dist, float* env, float* wig, float* gauss, cudaStream_t st);
void launch_geometry_bwd(const float* vec, const float* dist, const float* wig, const float* gauss,
                         const float* g_gauss, const float* g_env, const float* g_wig, int n_edges, float cutoff,
                         float* g_vec, cudaStream_t st);
void launch_force_reduce(const float* g_vec, const int* row_ptr, const int* sptr, const int* sedge, int n_nodes,
                         float* forces, cudaStream_t st);

void launch_ln_silu_fwd(float* u, float* h, const float* gamma, const float* beta, const float* bias,
                        const float* t_src, const float* t_tgt, const int* z, const int* src, const int* tgt,
                        int rows, cudaStream_t st);
void launch_ln_silu_bwd(const float* u, float* g, const float* gamma, const float* beta, int rows, cudaStream_t st);

void launch_gather_rotate_scale(const float* x, const int* src, const int* tgt, const float* wig, const float* rad,
                                long long e0, int n_e, float* A0, float* A1, float* A2, cudaStream_t st);
void launch_gather_rotate_bwd(const float* x, const int* row_ptr, const int* src, const float* wig, const float* rad,
                              long long e0, int node0, int n_nodes, const float* gA0, const float* gA1,
                              const float* gA2, float* g_rad, float* G, float* g_x, float* g_wig, cudaStream_t st);
void launch_source_reduce(const float* G, const int* sptr, const int* sedge, int n_nodes, float* g_x, cudaStream_t st);
void launch_combine_gate_fwd(const float* Y0, const float* Y1, const float* Y2, int n_e, float* B0, float* B1,
                             float* B2, cudaStream_t st);
void launch_combine_gate_bwd(const float* Y0, const float* Y1, const float* Y2, int n_e, const float* gB0,
                             const float* gB1, const float* gB2, float* gY0, float* gY1, float* gY2, cudaStream_t st);
void launch_rotate_back_reduce(int mode, const float* Z0, const float* Z1, const float* Z2, const int* row_ptr,
                               const float* wig, const float* env, float scale, long long e0, int node0,
                               int n_nodes, const float* base, float* out, cudaStream_t st);
void launch_rotate_back_bwd(int mode, const float* Z0, const float* Z1, const float* Z2, const int* tgt,
                            const float* wig, const float* env, float scale, long long e0, int n_e,
                            const float* g_out, float* gZ0, float* gZ1, float* gZ2, float* g_env, float* g_wig,
                            cudaStream_t st);

void launch_embed(const float* sphere_emb, const float* csd, const int* z, int n_nodes, float* x, cudaStream_t st);
void launch_rms_fwd(const float* x, const float* w_aff, const float* b_aff, const float* add0, int n_nodes, float* y,
                    cudaStream_t st);
void launch_rms_bwd(const float* x, const float* w_aff, const float* g_y, const float* g_add, int n_nodes,
                    float* g_x, cudaStream_t st);
void launch_ffn_gate_fwd(const float* y1, const float* gp, int n_nodes, float* a, cudaStream_t st);
void launch_ffn_gate_bwd(const float* y1, const float* gp, const float* g_a, int n_nodes, float* g_y1, float* g_gp,
                         cudaStream_t st);
void launch_eltwise(int mode, const float* a, const float* b, long long n, float* out, cudaStream_t st);
void launch_head_final(const float* p2, const float* w4, const float* b4, int n_nodes, float* node_e, float* g_p2,
                       cudaStream_t st);
void launch_energy_reduce(const float* node_e, int n_img, int n_atoms, double* energy, cudaStream_t st);
void launch_tile_int(const int* in, int n, int reps, int* out, cudaStream_t st);

}  // namespace umab
