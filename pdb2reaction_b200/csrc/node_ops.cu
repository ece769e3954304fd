// K11/K12 node-level kernels: embedding, equivariant RMS norm (fwd/adjoint), the FFN gate
// (fwd/adjoint), SiLU helpers for the energy head and the per-image energy reduction.
//
// fairchem: EquivariantRMSNormArraySphericalHarmonicsV2, SpectralAtomwise (GateActivation),
// MLP_EFS_Head -- reached from the reference through predict_unit.predict
// (pdb2reaction/uma_pysis.py:385).  One warp per atom, float4 per lane over the 128 channels.
// Twin: oracle/staged.py rms_fwd / rms_bwd / ffn_fwd / ffn_bwd.
#include "common.cuh"

namespace umab {

namespace {

constexpr float RMS_EPS = 1e-5f;
// balance weight of l-primary row r: 1 / ((2l+1) (lmax+1))
__device__ __forceinline__ constexpr float bal_w(int r) {
    return r == 0 ? (1.0f / 3.0f) : (r < 4 ? (1.0f / 9.0f) : (1.0f / 15.0f));
}
__device__ __forceinline__ constexpr int l_of(int r) { return r == 0 ? 0 : (r < 4 ? 1 : 2); }

__global__ void __launch_bounds__(256)
embed_kernel(const float* __restrict__ sphere_emb, const float* __restrict__ csd, const int* __restrict__ z,
             int n_nodes, float* __restrict__ x) {
    const int n = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (n >= n_nodes) return;
    float* xp = x + (long long)n * (9 * C) + lane * 4;
    st4(xp, f4add(ld4(sphere_emb + z[n] * C + lane * 4), ld4(csd + lane * 4)));
#pragma unroll
    for (int r = 1; r < 9; ++r) st4(xp + r * C, f4zero());
}

// y = rms_norm_sh(x) (+ add0 on the l=0 row)
__global__ void __launch_bounds__(256)
rms_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w_aff, const float* __restrict__ b_aff,
               const float* __restrict__ add0, int n_nodes, float* __restrict__ y) {
    const int n = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (n >= n_nodes) return;
    const float* xp = x + (long long)n * (9 * C) + lane * 4;
    float4 f[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) f[r] = ld4(xp + r * C);
    float mean0 = warp_sum(f4hsum(f[0])) * (1.0f / C);
    f[0] = make_float4(f[0].x - mean0, f[0].y - mean0, f[0].z - mean0, f[0].w - mean0);
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 9; ++r) s += bal_w(r) * f4dot(f[r], f[r]);
    s = warp_sum(s) * (1.0f / C);
    const float rs = rsqrtf(s + RMS_EPS);
    float* yp = y + (long long)n * (9 * C) + lane * 4;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        float4 o = f4mul(f4scale(f[r], rs), ld4(w_aff + l_of(r) * C + lane * 4));
        if (r == 0) {
            o = f4add(o, ld4(b_aff + lane * 4));
            if (add0) o = f4add(o, ld4(add0 + lane * 4));
        }
        st4(yp + r * C, o);
    }
}

// g_x = (g_add ? g_add : 0) + rms_bwd(x, g_y)
__global__ void __launch_bounds__(256)
rms_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w_aff, const float* __restrict__ g_y,
               const float* g_add, int n_nodes, float* g_x) {
    const int n = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (n >= n_nodes) return;
    const long long off = (long long)n * (9 * C) + lane * 4;
    float4 f[9], gw[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) f[r] = ld4(x + off + r * C);
    float mean0 = warp_sum(f4hsum(f[0])) * (1.0f / C);
    f[0] = make_float4(f[0].x - mean0, f[0].y - mean0, f[0].z - mean0, f[0].w - mean0);
    float s = 0.f, dot = 0.f;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        s += bal_w(r) * f4dot(f[r], f[r]);
        gw[r] = f4mul(ld4(g_y + off + r * C), ld4(w_aff + l_of(r) * C + lane * 4));
        dot += f4dot(gw[r], f[r]);
    }
    s = warp_sum(s) * (1.0f / C);
    dot = warp_sum(dot);
    const float rs = rsqrtf(s + RMS_EPS);
    const float k = rs * rs * rs * dot * (1.0f / C);
    float4 gf0;
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        float4 gf = f4sub(f4scale(gw[r], rs), f4scale(f[r], k * bal_w(r)));
        if (r == 0) { gf0 = gf; continue; }
        if (g_add) gf = f4add(gf, ld4(g_add + off + r * C));
        st4(g_x + off + r * C, gf);
    }
    float m = warp_sum(f4hsum(gf0)) * (1.0f / C);
    gf0 = make_float4(gf0.x - m, gf0.y - m, gf0.z - m, gf0.w - m);
    if (g_add) gf0 = f4add(gf0, ld4(g_add + off));
    st4(g_x + off, gf0);
}

// a[0] = silu(y1[0]);  a[r] = y1[r] * sigmoid(silu(gp[l_r - 1]))      (y1, a: [N,9,H]; gp: [N,2H])
__global__ void __launch_bounds__(256)
ffn_gate_fwd_kernel(const float* __restrict__ y1, const float* __restrict__ gp, int n_nodes, float* __restrict__ a) {
    const int n = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (n >= n_nodes) return;
    const long long off = (long long)n * (9 * H) + lane * 4;
    float4 gate[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        float4 v = ld4(gp + (long long)n * (2 * H) + l * H + lane * 4);
        gate[l] = make_float4(sigmoidf_(siluf_(v.x)), sigmoidf_(siluf_(v.y)), sigmoidf_(siluf_(v.z)), sigmoidf_(siluf_(v.w)));
    }
    float4 v0 = ld4(y1 + off);
    st4(a + off, make_float4(siluf_(v0.x), siluf_(v0.y), siluf_(v0.z), siluf_(v0.w)));
#pragma unroll
    for (int r = 1; r < 9; ++r) st4(a + off + r * H, f4mul(ld4(y1 + off + r * H), gate[l_of(r) - 1]));
}

// g_a -> (g_y1 [N,9,H] (may alias g_a), g_gp [N,2H])
__global__ void __launch_bounds__(256)
ffn_gate_bwd_kernel(const float* __restrict__ y1, const float* __restrict__ gp, const float* g_a, int n_nodes,
                    float* g_y1, float* __restrict__ g_gp) {
    const int n = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (n >= n_nodes) return;
    const long long off = (long long)n * (9 * H) + lane * 4;
    float4 gpv[2], sg[2], gs[2], gg[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        gpv[l] = ld4(gp + (long long)n * (2 * H) + l * H + lane * 4);
        gs[l] = make_float4(siluf_(gpv[l].x), siluf_(gpv[l].y), siluf_(gpv[l].z), siluf_(gpv[l].w));
        sg[l] = make_float4(sigmoidf_(gs[l].x), sigmoidf_(gs[l].y), sigmoidf_(gs[l].z), sigmoidf_(gs[l].w));
        gg[l] = f4zero();
    }
    float4 ga[9], yv[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) { ga[r] = ld4(g_a + off + r * H); yv[r] = ld4(y1 + off + r * H); }
    st4(g_y1 + off, make_float4(ga[0].x * dsiluf_(yv[0].x), ga[0].y * dsiluf_(yv[0].y),
                                ga[0].z * dsiluf_(yv[0].z), ga[0].w * dsiluf_(yv[0].w)));
#pragma unroll
    for (int r = 1; r < 9; ++r) {
        const int l = l_of(r) - 1;
        gg[l] = f4add(gg[l], f4mul(ga[r], yv[r]));
        st4(g_y1 + off + r * H, f4mul(ga[r], sg[l]));
    }
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        float4 o;
        o.x = gg[l].x * sg[l].x * (1.f - sg[l].x) * dsiluf_(gpv[l].x);
        o.y = gg[l].y * sg[l].y * (1.f - sg[l].y) * dsiluf_(gpv[l].y);
        o.z = gg[l].z * sg[l].z * (1.f - sg[l].z) * dsiluf_(gpv[l].z);
        o.w = gg[l].w * sg[l].w * (1.f - sg[l].w) * dsiluf_(gpv[l].w);
        st4(g_gp + (long long)n * (2 * H) + l * H + lane * 4, o);
    }
}

// elementwise over n4 float4s: mode 0: out = silu(in); mode 1: out = g * dsilu(p); mode 2: out = a + b
template <int MODE>
__global__ void eltwise_kernel(const float* a, const float* b, long long n4, float* out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 va = ld4(a + i * 4);
    float4 o;
    if (MODE == 0) {
        o = make_float4(siluf_(va.x), siluf_(va.y), siluf_(va.z), siluf_(va.w));
    } else if (MODE == 1) {
        float4 p = ld4(b + i * 4);
        o = make_float4(va.x * dsiluf_(p.x), va.y * dsiluf_(p.y), va.z * dsiluf_(p.z), va.w * dsiluf_(p.w));
    } else {
        o = f4add(va, ld4(b + i * 4));
    }
    st4(out + i * 4, o);
}

// node_e[n] = silu(p2[n]) . w4 + b4 ;   g_p2[n] = w4 * dsilu(p2[n])  (adjoint seed, dE_total/dnode_e = 1)
__global__ void __launch_bounds__(256)
head_final_kernel(const float* __restrict__ p2, const float* __restrict__ w4, const float* __restrict__ b4,
                  int n_nodes, float* __restrict__ node_e, float* __restrict__ g_p2) {
    const int n = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (n >= n_nodes) return;
    float4 p = ld4(p2 + (long long)n * H + lane * 4);
    float4 w = ld4(w4 + lane * 4);
    float4 s = make_float4(siluf_(p.x), siluf_(p.y), siluf_(p.z), siluf_(p.w));
    float e = warp_sum(f4dot(s, w));
    if (lane == 0) node_e[n] = e + b4[0];
    if (g_p2)
        st4(g_p2 + (long long)n * H + lane * 4,
            make_float4(w.x * dsiluf_(p.x), w.y * dsiluf_(p.y), w.z * dsiluf_(p.z), w.w * dsiluf_(p.w)));
}

// E[img] = sum of node energies, accumulated in double in a fixed order (reference quirk Q2:
// fp32 totals lose digits at 1e4 atoms; per-atom terms stay fp32)
__global__ void __launch_bounds__(256)
energy_reduce_kernel(const float* __restrict__ node_e, int n_atoms, double* __restrict__ energy) {
    __shared__ double part[256];
    const float* p = node_e + (long long)blockIdx.x * n_atoms;
    double s = 0.0;
    for (int i = threadIdx.x; i < n_atoms; i += 256) s += (double)p[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) energy[blockIdx.x] = part[0];
}

__global__ void tile_int_kernel(const int* __restrict__ in, int n, int reps, int* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long long)n * reps) out[i] = in[i % n];
}

}  // namespace

void launch_tile_int(const int* in, int n, int reps, int* out, cudaStream_t st) {
    long long tot = (long long)n * reps;
    if (tot <= 0) return;
    tile_int_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(in, n, reps, out);
    UMAB_LAUNCH_CHECK();
}

void launch_embed(const float* sphere_emb, const float* csd, const int* z, int n_nodes, float* x, cudaStream_t st) {
    if (n_nodes <= 0) return;
    embed_kernel<<<(n_nodes + 7) / 8, 256, 0, st>>>(sphere_emb, csd, z, n_nodes, x);
    UMAB_LAUNCH_CHECK();
}
void launch_rms_fwd(const float* x, const float* w_aff, const float* b_aff, const float* add0, int n_nodes,
                    float* y, cudaStream_t st) {
    if (n_nodes <= 0) return;
    rms_fwd_kernel<<<(n_nodes + 7) / 8, 256, 0, st>>>(x, w_aff, b_aff, add0, n_nodes, y);
    UMAB_LAUNCH_CHECK();
}
void launch_rms_bwd(const float* x, const float* w_aff, const float* g_y, const float* g_add, int n_nodes,
                    float* g_x, cudaStream_t st) {
    if (n_nodes <= 0) return;
    rms_bwd_kernel<<<(n_nodes + 7) / 8, 256, 0, st>>>(x, w_aff, g_y, g_add, n_nodes, g_x);
    UMAB_LAUNCH_CHECK();
}
void launch_ffn_gate_fwd(const float* y1, const float* gp, int n_nodes, float* a, cudaStream_t st) {
    if (n_nodes <= 0) return;
    ffn_gate_fwd_kernel<<<(n_nodes + 7) / 8, 256, 0, st>>>(y1, gp, n_nodes, a);
    UMAB_LAUNCH_CHECK();
}
void launch_ffn_gate_bwd(const float* y1, const float* gp, const float* g_a, int n_nodes, float* g_y1,
                         float* g_gp, cudaStream_t st) {
    if (n_nodes <= 0) return;
    ffn_gate_bwd_kernel<<<(n_nodes + 7) / 8, 256, 0, st>>>(y1, gp, g_a, n_nodes, g_y1, g_gp);
    UMAB_LAUNCH_CHECK();
}
void launch_eltwise(int mode, const float* a, const float* b, long long n, float* out, cudaStream_t st) {
    if (n <= 0) return;
    long long n4 = n / 4;
    unsigned grid = (unsigned)((n4 + 255) / 256);
    if (mode == 0) eltwise_kernel<0><<<grid, 256, 0, st>>>(a, b, n4, out);
    else if (mode == 1) eltwise_kernel<1><<<grid, 256, 0, st>>>(a, b, n4, out);
    else eltwise_kernel<2><<<grid, 256, 0, st>>>(a, b, n4, out);
    UMAB_LAUNCH_CHECK();
}
void launch_head_final(const float* p2, const float* w4, const float* b4, int n_nodes, float* node_e,
                       float* g_p2, cudaStream_t st) {
    if (n_nodes <= 0) return;
    head_final_kernel<<<(n_nodes + 7) / 8, 256, 0, st>>>(p2, w4, b4, n_nodes, node_e, g_p2);
    UMAB_LAUNCH_CHECK();
}
void launch_energy_reduce(const float* node_e, int n_img, int n_atoms, double* energy, cudaStream_t st) {
    if (n_img <= 0) return;
    energy_reduce_kernel<<<n_img, 256, 0, st>>>(node_e, n_atoms, energy);
    UMAB_LAUNCH_CHECK();
}

}  // namespace umab
