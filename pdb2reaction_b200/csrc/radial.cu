// K4 (elementwise half): LayerNorm + SiLU of the radial MLPs, forward and adjoint.
//
// fairchem RadialMLP = Linear -> LayerNorm -> SiLU (x2) -> Linear, reached from the reference via
// predict_unit.predict (pdb2reaction/uma_pysis.py:385).  The three Linears are GEMMs (gemm_*.cu);
// the first one only contracts the 64 Gaussian columns of x_edge -- the source/target element
// embedding columns are folded into per-element tables T_src/T_tgt = emb . W1_part^T that are
// added here, together with the bias, before the normalisation.  One warp per edge row of 128.
// Twin: oracle/staged.py ln_silu_fwd / ln_silu_bwd / radial_fwd.
#include "common.cuh"

namespace umab {

namespace {

constexpr float LN_EPS = 1e-5f;

__global__ void __launch_bounds__(256)
ln_silu_fwd_kernel(float* u, float* __restrict__ h, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ bias,
                   const float* __restrict__ t_src, const float* __restrict__ t_tgt,
                   const int* __restrict__ z, const int* __restrict__ src, const int* __restrict__ tgt,
                   int rows) {
    const int row = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (row >= rows) return;
    float* up = u + (long long)row * 128 + lane * 4;
    float4 v = ld4(up);
    if (bias) v = f4add(v, ld4(bias + lane * 4));
    if (t_src) {
        int zs = z[src[row]], zt = z[tgt[row]];
        v = f4add(v, f4add(ld4(t_src + zs * 128 + lane * 4), ld4(t_tgt + zt * 128 + lane * 4)));
    }
    if (bias || t_src) st4(up, v);
    float mean = warp_sum(f4hsum(v)) * (1.0f / 128.0f);
    float4 c = make_float4(v.x - mean, v.y - mean, v.z - mean, v.w - mean);
    float var = warp_sum(f4dot(c, c)) * (1.0f / 128.0f);
    float rstd = rsqrtf(var + LN_EPS);
    float4 g = ld4(gamma + lane * 4), b = ld4(beta + lane * 4);
    float4 o;
    o.x = siluf_(c.x * rstd * g.x + b.x);
    o.y = siluf_(c.y * rstd * g.y + b.y);
    o.z = siluf_(c.z * rstd * g.z + b.z);
    o.w = siluf_(c.w * rstd * g.w + b.w);
    st4(h + (long long)row * 128 + lane * 4, o);
}

// g: in = dL/dh, out = dL/du (in place)
__global__ void __launch_bounds__(256)
ln_silu_bwd_kernel(const float* __restrict__ u, float* g, const float* __restrict__ gamma,
                   const float* __restrict__ beta, int rows) {
    const int row = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (row >= rows) return;
    float4 v = ld4(u + (long long)row * 128 + lane * 4);
    float mean = warp_sum(f4hsum(v)) * (1.0f / 128.0f);
    float4 c = make_float4(v.x - mean, v.y - mean, v.z - mean, v.w - mean);
    float var = warp_sum(f4dot(c, c)) * (1.0f / 128.0f);
    float rstd = rsqrtf(var + LN_EPS);
    float4 xh = f4scale(c, rstd);
    float4 ga = ld4(gamma + lane * 4), be = ld4(beta + lane * 4);
    float* gp = g + (long long)row * 128 + lane * 4;
    float4 go = ld4(gp);
    float4 gx;
    gx.x = go.x * dsiluf_(xh.x * ga.x + be.x) * ga.x;
    gx.y = go.y * dsiluf_(xh.y * ga.y + be.y) * ga.y;
    gx.z = go.z * dsiluf_(xh.z * ga.z + be.z) * ga.z;
    gx.w = go.w * dsiluf_(xh.w * ga.w + be.w) * ga.w;
    float m1 = warp_sum(f4hsum(gx)) * (1.0f / 128.0f);
    float m2 = warp_sum(f4dot(gx, xh)) * (1.0f / 128.0f);
    float4 o;
    o.x = rstd * (gx.x - m1 - xh.x * m2);
    o.y = rstd * (gx.y - m1 - xh.y * m2);
    o.z = rstd * (gx.z - m1 - xh.z * m2);
    o.w = rstd * (gx.w - m1 - xh.w * m2);
    st4(gp, o);
}

}  // namespace

void launch_ln_silu_fwd(float* u, float* h, const float* gamma, const float* beta, const float* bias,
                        const float* t_src, const float* t_tgt, const int* z, const int* src,
                        const int* tgt, int rows, cudaStream_t st) {
    if (rows <= 0) return;
    ln_silu_fwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(u, h, gamma, beta, bias, t_src, t_tgt, z, src, tgt, rows);
    UMAB_LAUNCH_CHECK();
}

void launch_ln_silu_bwd(const float* u, float* g, const float* gamma, const float* beta, int rows,
                        cudaStream_t st) {
    if (rows <= 0) return;
    ln_silu_bwd_kernel<<<(rows + 7) / 8, 256, 0, st>>>(u, g, gamma, beta, rows);
    UMAB_LAUNCH_CHECK();
}

}  // namespace umab
