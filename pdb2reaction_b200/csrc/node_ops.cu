// K11/K12 node-level kernels: embedding, equivariant RMS norm (fwd/adjoint), the FFN gate
// (fwd/adjoint), SiLU helpers for the energy head and the per-image energy reduction.
//
// fairchem: EquivariantRMSNormArraySphericalHarmonicsV2, SpectralAtomwise (GateActivation),
// MLP_EFS_Head -- reached from the reference through predict_unit.predict
// (pdb2reaction/uma_pysis.py:385).  One warp per atom, float4 per lane over the 128 channels.
// Twin: oracle/staged.py rms_fwd / rms_bwd / ffn_fwd / ffn_bwd.
#include "dual.cuh"

namespace umab {

namespace {

constexpr float RMS_EPS = 1e-5f;
// balance weight of l-primary row r: 1 / ((2l+1) (lmax+1))
__device__ __forceinline__ constexpr float bal_w(int r) {
    return r == 0 ? (1.0f / 3.0f) : (r < 4 ? (1.0f / 9.0f) : (1.0f / 15.0f));
}
__device__ __forceinline__ constexpr int l_of(int r) { return r == 0 ? 0 : (r < 4 ? 1 : 2); }

// multiply / add a constant (weight) float4 into a value-or-dual vector
__device__ __forceinline__ float4 cmul(float4 a, float4 w) { return f4mul(a, w); }
__device__ __forceinline__ D4 cmul(D4 a, float4 w) { return {f4mul(a.v, w), f4mul(a.d, w)}; }
__device__ __forceinline__ float4 cadd(float4 a, float4 w) { return f4add(a, w); }
__device__ __forceinline__ D4 cadd(D4 a, float4 w) { return {f4add(a.v, w), a.d}; }

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
template <class S>
__global__ void __launch_bounds__(256)
rms_fwd_kernel(GP<S> x, const float* __restrict__ w_aff, const float* __restrict__ b_aff,
               const float* __restrict__ add0, int n_nodes, GP<S> y) {
    using V = typename VecOf<S>::type;
    const int n = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (n >= n_nodes) return;
    const long long off = (long long)n * (9 * C) + lane * 4;
    V f[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) f[r] = x.ld4(off + r * C);
    S mean0 = warp_sum(vhsum(f[0])) * (1.0f / C);
    f[0] = vsubs(f[0], mean0);
    S s = cst<S>(0.f);
#pragma unroll
    for (int r = 0; r < 9; ++r) s = s + bal_w(r) * vdot(f[r], f[r]);
    s = warp_sum(s) * (1.0f / C);
    const S rs = s_rsqrt(s + RMS_EPS);
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        V o = cmul(vscale(f[r], rs), ld4(w_aff + l_of(r) * C + lane * 4));
        if (r == 0) {
            o = cadd(o, ld4(b_aff + lane * 4));
            if (add0) o = cadd(o, ld4(add0 + lane * 4));
        }
        y.st4(off + r * C, o);
    }
}

// g_x = (g_add ? g_add : 0) + rms_bwd(x, g_y)
template <class S>
__global__ void __launch_bounds__(256)
rms_bwd_kernel(GP<S> x, const float* __restrict__ w_aff, GP<S> g_y, GP<S> g_add, int n_nodes, GP<S> g_x) {
    using V = typename VecOf<S>::type;
    const int n = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (n >= n_nodes) return;
    const long long off = (long long)n * (9 * C) + lane * 4;
    V f[9], gw[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) f[r] = x.ld4(off + r * C);
    S mean0 = warp_sum(vhsum(f[0])) * (1.0f / C);
    f[0] = vsubs(f[0], mean0);
    S s = cst<S>(0.f), dot = cst<S>(0.f);
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        s = s + bal_w(r) * vdot(f[r], f[r]);
        gw[r] = cmul(g_y.ld4(off + r * C), ld4(w_aff + l_of(r) * C + lane * 4));
        dot = dot + vdot(gw[r], f[r]);
    }
    s = warp_sum(s) * (1.0f / C);
    dot = warp_sum(dot);
    const S rs = s_rsqrt(s + RMS_EPS);
    const S k = rs * rs * rs * dot * (1.0f / C);
    V gf0 = vzero<V>();
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        V gf = vsub(vscale(gw[r], rs), vscale(f[r], k * bal_w(r)));
        if (r == 0) { gf0 = gf; continue; }
        if (g_add) gf = vadd(gf, g_add.ld4(off + r * C));
        g_x.st4(off + r * C, gf);
    }
    S m = warp_sum(vhsum(gf0)) * (1.0f / C);
    gf0 = vsubs(gf0, m);
    if (g_add) gf0 = vadd(gf0, g_add.ld4(off));
    g_x.st4(off, gf0);
}

// a[0] = silu(y1[0]);  a[r] = y1[r] * sigmoid(silu(gp[l_r - 1]))      (y1, a: [N,9,H]; gp: [N,2H])
template <class S>
__global__ void __launch_bounds__(256)
ffn_gate_fwd_kernel(GP<S> y1, GP<S> gp, int n_nodes, GP<S> a) {
    using V = typename VecOf<S>::type;
    const int n = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (n >= n_nodes) return;
    const long long off = (long long)n * (9 * H) + lane * 4;
    V gate[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) gate[l] = vsigmoid(vsilu(gp.ld4((long long)n * (2 * H) + l * H + lane * 4)));
    a.st4(off, vsilu(y1.ld4(off)));
#pragma unroll
    for (int r = 1; r < 9; ++r) a.st4(off + r * H, vmul(y1.ld4(off + r * H), gate[l_of(r) - 1]));
}

// g_a -> (g_y1 [N,9,H] (may alias g_a), g_gp [N,2H])
template <class S>
__global__ void __launch_bounds__(256)
ffn_gate_bwd_kernel(GP<S> y1, GP<S> gp, GP<S> g_a, int n_nodes, GP<S> g_y1, GP<S> g_gp) {
    using V = typename VecOf<S>::type;
    const int n = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (n >= n_nodes) return;
    const long long off = (long long)n * (9 * H) + lane * 4;
    V gpv[2], sg[2], gg[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        gpv[l] = gp.ld4((long long)n * (2 * H) + l * H + lane * 4);
        sg[l] = vsigmoid(vsilu(gpv[l]));
        gg[l] = vzero<V>();
    }
    V ga[9], yv[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) { ga[r] = g_a.ld4(off + r * H); yv[r] = y1.ld4(off + r * H); }
    g_y1.st4(off, vmul(ga[0], vdsilu(yv[0])));
#pragma unroll
    for (int r = 1; r < 9; ++r) {
        const int l = l_of(r) - 1;
        gg[l] = vadd(gg[l], vmul(ga[r], yv[r]));
        g_y1.st4(off + r * H, vmul(ga[r], sg[l]));
    }
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        V ds = vsub(sg[l], vmul(sg[l], sg[l]));                       // sigma' = s (1 - s)
        g_gp.st4((long long)n * (2 * H) + l * H + lane * 4, vmul(vmul(gg[l], ds), vdsilu(gpv[l])));
    }
}

// elementwise over n4 float4s: mode 0: out = silu(in); mode 1: out = g * dsilu(p); mode 2: out = a + b
template <int MODE, class S>
__global__ void eltwise_kernel(GP<S> a, GP<S> b, long long n4, GP<S> out) {
    using V = typename VecOf<S>::type;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    V va = a.ld4(i * 4);
    V o;
    if (MODE == 0) o = vsilu(va);
    else if (MODE == 1) o = vmul(va, vdsilu(b.ld4(i * 4)));
    else o = vadd(va, b.ld4(i * 4));
    out.st4(i * 4, o);
}

// node_e[n] = silu(p2[n]) . w4 + b4 ;   g_p2[n] = w4 * dsilu(p2[n])  (adjoint seed, dE_total/dnode_e = 1)
template <class S>
__global__ void __launch_bounds__(256)
head_final_kernel(GP<S> p2, const float* __restrict__ w4, const float* __restrict__ b4, int n_nodes,
                  float* __restrict__ node_e, GP<S> g_p2) {
    using V = typename VecOf<S>::type;
    const int n = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (n >= n_nodes) return;
    V p = p2.ld4((long long)n * H + lane * 4);
    float4 w = ld4(w4 + lane * 4);
    S e = warp_sum(vhsum(cmul(vsilu(p), w)));
    if (lane == 0) node_e[n] = val(e) + b4[0];
    if (g_p2) g_p2.st4((long long)n * H + lane * 4, cmul(vdsilu(p), w));
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

// block 0 of `reps` consecutive blocks of `block4` float4 copied over blocks 1 .. reps-1 (read once, written reps-1 times)
__global__ void __launch_bounds__(256)
replicate_block_kernel(float4* __restrict__ base, long long block4, int reps) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < block4; i += stride) {
        const float4 v = base[i];
        for (int r = 1; r < reps; ++r) base[r * block4 + i] = v;
    }
}
// flag |= 1 when any image of pos [n_img, n3] differs (bitwise) from image 0
__global__ void same_images_kernel(const float* __restrict__ pos, long long n3, int n_img, int* __restrict__ flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3 * (n_img - 1)) return;
    if (__float_as_uint(pos[n3 + i]) != __float_as_uint(pos[i % n3])) atomicOr(flag, 1);
}

}  // namespace

void launch_replicate_block(float* base, long long block_floats, int reps, cudaStream_t st) {
    if (reps <= 1 || block_floats <= 0) return;
    if (block_floats % 4 != 0 || (reinterpret_cast<uintptr_t>(base) & 15)) throw CudaError("replicate_block: unaligned block");
    const long long block4 = block_floats / 4;
    const long long want = (block4 + 255) / 256;
    replicate_block_kernel<<<(unsigned)std::min<long long>(want, 148LL * 16), 256, 0, st>>>(reinterpret_cast<float4*>(base), block4, reps);
    UMAB_LAUNCH_CHECK();
}
void launch_same_images(const float* pos, long long n3, int n_img, int* flag, cudaStream_t st) {
    UMAB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
    const long long tot = n3 * (n_img - 1);
    if (tot <= 0) return;
    same_images_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(pos, n3, n_img, flag);
    UMAB_LAUNCH_CHECK();
}
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
template <class S>
void launch_rms_fwd_t(GP<S> x, const float* w_aff, const float* b_aff, const float* add0, int n_nodes, GP<S> y,
                      cudaStream_t st) {
    if (n_nodes <= 0) return;
    rms_fwd_kernel<S><<<(n_nodes + 7) / 8, 256, 0, st>>>(x, w_aff, b_aff, add0, n_nodes, y);
    UMAB_LAUNCH_CHECK();
}
template <class S>
void launch_rms_bwd_t(GP<S> x, const float* w_aff, GP<S> g_y, GP<S> g_add, int n_nodes, GP<S> g_x, cudaStream_t st) {
    if (n_nodes <= 0) return;
    rms_bwd_kernel<S><<<(n_nodes + 7) / 8, 256, 0, st>>>(x, w_aff, g_y, g_add, n_nodes, g_x);
    UMAB_LAUNCH_CHECK();
}
template <class S>
void launch_ffn_gate_fwd_t(GP<S> y1, GP<S> gp, int n_nodes, GP<S> a, cudaStream_t st) {
    if (n_nodes <= 0) return;
    ffn_gate_fwd_kernel<S><<<(n_nodes + 7) / 8, 256, 0, st>>>(y1, gp, n_nodes, a);
    UMAB_LAUNCH_CHECK();
}
template <class S>
void launch_ffn_gate_bwd_t(GP<S> y1, GP<S> gp, GP<S> g_a, int n_nodes, GP<S> g_y1, GP<S> g_gp, cudaStream_t st) {
    if (n_nodes <= 0) return;
    ffn_gate_bwd_kernel<S><<<(n_nodes + 7) / 8, 256, 0, st>>>(y1, gp, g_a, n_nodes, g_y1, g_gp);
    UMAB_LAUNCH_CHECK();
}
template <class S>
void launch_eltwise_t(int mode, GP<S> a, GP<S> b, long long n, GP<S> out, cudaStream_t st) {
    if (n <= 0) return;
    long long n4 = n / 4;
    unsigned grid = (unsigned)((n4 + 255) / 256);
    if (mode == 0) eltwise_kernel<0, S><<<grid, 256, 0, st>>>(a, b, n4, out);
    else if (mode == 1) eltwise_kernel<1, S><<<grid, 256, 0, st>>>(a, b, n4, out);
    else eltwise_kernel<2, S><<<grid, 256, 0, st>>>(a, b, n4, out);
    UMAB_LAUNCH_CHECK();
}
template <class S>
void launch_head_final_t(GP<S> p2, const float* w4, const float* b4, int n_nodes, float* node_e, GP<S> g_p2,
                         cudaStream_t st) {
    if (n_nodes <= 0) return;
    head_final_kernel<S><<<(n_nodes + 7) / 8, 256, 0, st>>>(p2, w4, b4, n_nodes, node_e, g_p2);
    UMAB_LAUNCH_CHECK();
}
void launch_energy_reduce(const float* node_e, int n_img, int n_atoms, double* energy, cudaStream_t st) {
    if (n_img <= 0) return;
    energy_reduce_kernel<<<n_img, 256, 0, st>>>(node_e, n_atoms, energy);
    UMAB_LAUNCH_CHECK();
}

#define UMAB_INST(S)                                                                                              \
    template void launch_rms_fwd_t<S>(GP<S>, const float*, const float*, const float*, int, GP<S>, cudaStream_t); \
    template void launch_rms_bwd_t<S>(GP<S>, const float*, GP<S>, GP<S>, int, GP<S>, cudaStream_t);               \
    template void launch_ffn_gate_fwd_t<S>(GP<S>, GP<S>, int, GP<S>, cudaStream_t);                               \
    template void launch_ffn_gate_bwd_t<S>(GP<S>, GP<S>, GP<S>, int, GP<S>, GP<S>, cudaStream_t);                 \
    template void launch_eltwise_t<S>(int, GP<S>, GP<S>, long long, GP<S>, cudaStream_t);                         \
    template void launch_head_final_t<S>(GP<S>, const float*, const float*, int, float*, GP<S>, cudaStream_t);
UMAB_INST(float)
UMAB_INST(D1)
#undef UMAB_INST

}  // namespace umab
