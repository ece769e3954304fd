// K4 (elementwise half): LayerNorm + SiLU of the radial MLPs, forward and adjoint.
//
// fairchem RadialMLP = Linear -> LayerNorm -> SiLU (x2) -> Linear, reached from the reference via
// predict_unit.predict (pdb2reaction/uma_pysis.py:385).  The three Linears are GEMMs (gemm_*.cu);
// the first one only contracts the 64 Gaussian columns of x_edge -- the source/target element
// embedding columns are folded into per-element tables T_src/T_tgt = emb . W1_part^T that are
// added here, together with the bias, before the normalisation.  One warp per edge row of 128.
// Twin: oracle/staged.py ln_silu_fwd / ln_silu_bwd / radial_fwd.
#include "dual.cuh"

namespace umab {

namespace {

constexpr float LN_EPS = 1e-5f;

// S = float (energy/forces) or D1 (value + tangent, analytic Hessian columns)
// RPW rows per warp: the loads of all RPW rows are requested before the first row is reduced (one row per warp left
// 16 short-lived warps per SM with one 512 B request each in flight; ncu: issue-active 60 %, 0.69 of the HBM peak)
constexpr int RPW = 4;

template <class S>
__global__ void __launch_bounds__(256)
ln_silu_fwd_kernel(GP<S> u, AP<S> h, const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ bias, const float* __restrict__ t_src, const float* __restrict__ t_tgt,
                   const int* __restrict__ z, const int* __restrict__ src, const int* __restrict__ tgt, int rows_all,
                   ImgShare sh) {
    using V = typename VecOf<S>::type;
    // a warp takes RPW consecutive rows of ONE image (sh: groups of RPW rows per image; u = GEMM output of image 0)
    long long grp;
    int img;
    ImgShare shg = sh;
    shg.rows = (sh.rows + RPW - 1) / RPW;
    if (!share_row<S>(shg, 8, ((long long)rows_all + RPW - 1) / RPW, grp, img)) return;
    const int lane = threadIdx.x % 32;
    int row0, rows;                                     // this warp's first row and the end of its image (or of the launch)
    if (sh.n_img > 1) {
        row0 = img * sh.rows + (int)(grp - (long long)img * shg.rows) * RPW;
        rows = (img + 1) * sh.rows;
        u = u.vback((long long)img * sh.rows * 128);
    } else {
        row0 = (int)grp * RPW;
        rows = rows_all;
    }
    if (row0 >= rows) return;
    V v[RPW];
    float4 add[RPW];
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
        const int row = min(row0 + i, rows - 1);
        v[i] = u.ld4((long long)row * 128 + lane * 4);
        add[i] = f4zero();
        if (t_src) {
            const int zs = z[src[row]], zt = z[tgt[row]];
            add[i] = f4add(ld4(t_src + zs * 128 + lane * 4), ld4(t_tgt + zt * 128 + lane * 4));
        }
    }
    const float4 g = ld4(gamma + lane * 4), b = ld4(beta + lane * 4);
    float4 bs = f4zero();
    if (bias) bs = ld4(bias + lane * 4);
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
        const int row = row0 + i;
        if (row >= rows) break;
        const long long off = (long long)row * 128 + lane * 4;
        V x = v[i];
        if (bias || t_src) {
            // constants (bias, per-element tables) only touch the value plane
            const float4 a = f4add(bs, add[i]);
            if constexpr (std::is_same<S, float>::value) x = f4add(x, a); else x.v = f4add(x.v, a);
            u.st4(off, x);
        }
        S mean = warp_sum(vhsum(x)) * (1.0f / 128.0f);
        V c = vsubs(x, mean);
        S var = warp_sum(vdot(c, c)) * (1.0f / 128.0f);
        S rstd = s_rsqrt(var + LN_EPS);
        V y = vscale(c, rstd);
        if constexpr (std::is_same<S, float>::value) {
            y = f4add(f4mul(y, g), b);
        } else {
            y.v = f4add(f4mul(y.v, g), b);
            y.d = f4mul(y.d, g);
        }
        h.st4(off, vsilu(y));
    }
}

// g = dL/dh  ->  out = dL/du (the A operand of the next adjoint GEMM; never aliases g)
template <class S>
__global__ void __launch_bounds__(256)
ln_silu_bwd_kernel(GP<S> u, GP<S> g, AP<S> out, const float* __restrict__ gamma, const float* __restrict__ beta, int rows_all,
                   ImgShare sh, bool share_u) {
    using V = typename VecOf<S>::type;
    long long grp;
    int img;
    ImgShare shg = sh;
    shg.rows = (sh.rows + RPW - 1) / RPW;
    if (!share_row<S>(shg, 8, ((long long)rows_all + RPW - 1) / RPW, grp, img)) return;
    const int lane = threadIdx.x % 32;
    int row0, rows;
    if (sh.n_img > 1) {
        row0 = img * sh.rows + (int)(grp - (long long)img * shg.rows) * RPW;
        rows = (img + 1) * sh.rows;
        if (share_u) u = u.vback((long long)img * sh.rows * 128);        // u of the table-adding first layer is per image
        g = g.vback((long long)img * sh.rows * 128);
    } else {
        row0 = (int)grp * RPW;
        rows = rows_all;
    }
    if (row0 >= rows) return;
    V v[RPW], gg[RPW];
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
        const long long off = (long long)min(row0 + i, rows - 1) * 128 + lane * 4;
        v[i] = u.ld4(off);
        gg[i] = g.ldg4(off);
    }
    const float4 ga = ld4(gamma + lane * 4), be = ld4(beta + lane * 4);
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
        const int row = row0 + i;
        if (row >= rows) break;
        const long long off = (long long)row * 128 + lane * 4;
        S mean = warp_sum(vhsum(v[i])) * (1.0f / 128.0f);
        V c = vsubs(v[i], mean);
        S var = warp_sum(vdot(c, c)) * (1.0f / 128.0f);
        S rstd = s_rsqrt(var + LN_EPS);
        V xh = vscale(c, rstd);
        V y = xh;
        if constexpr (std::is_same<S, float>::value) {
            y = f4add(f4mul(y, ga), be);
        } else {
            y.v = f4add(f4mul(y.v, ga), be);
            y.d = f4mul(y.d, ga);
        }
        V gx = vmul(gg[i], vdsilu(y));
        if constexpr (std::is_same<S, float>::value) {
            gx = f4mul(gx, ga);
        } else {
            gx.v = f4mul(gx.v, ga);
            gx.d = f4mul(gx.d, ga);
        }
        S m1 = warp_sum(vhsum(gx)) * (1.0f / 128.0f);
        S m2 = warp_sum(vdot(gx, xh)) * (1.0f / 128.0f);
        V o = vscale(vsub(vsubs(gx, m1), vscale(xh, m2)), rstd);
        out.st4(off, o);
    }
}

}  // namespace

// groups of RPW rows: per image when the launch shares value planes across images (a group never spans two images)
static inline ImgShare group_share(ImgShare sh) { return ImgShare{sh.n_img, (sh.rows + RPW - 1) / RPW}; }

template <class S>
void launch_ln_silu_fwd_t(GP<S> u, AP<S> h, const float* gamma, const float* beta, const float* bias,
                          const float* t_src, const float* t_tgt, const int* z, const int* src, const int* tgt,
                          int rows, cudaStream_t st, ImgShare sh) {
    if (rows <= 0) return;
    if (bias || t_src) sh = ImgShare{0, 0};           // this variant updates u in place: every image owns its rows
    const unsigned grid = share_grid(group_share(sh), 8, ((long long)rows + RPW - 1) / RPW);
    ln_silu_fwd_kernel<S><<<grid, 256, 0, st>>>(u, h, gamma, beta, bias, t_src, t_tgt, z, src, tgt, rows, sh);
    UMAB_LAUNCH_CHECK();
}
template <class S>
void launch_ln_silu_bwd_t(GP<S> u, GP<S> g, AP<S> out, const float* gamma, const float* beta, int rows, cudaStream_t st,
                          ImgShare sh, bool share_u) {
    if (rows <= 0) return;
    const unsigned grid = share_grid(group_share(sh), 8, ((long long)rows + RPW - 1) / RPW);
    ln_silu_bwd_kernel<S><<<grid, 256, 0, st>>>(u, g, out, gamma, beta, rows, sh, share_u);
    UMAB_LAUNCH_CHECK();
}
template void launch_ln_silu_fwd_t<float>(GP<float>, AP<float>, const float*, const float*, const float*, const float*,
                                          const float*, const int*, const int*, const int*, int, cudaStream_t, ImgShare);
template void launch_ln_silu_fwd_t<D1>(GP<D1>, AP<D1>, const float*, const float*, const float*, const float*,
                                       const float*, const int*, const int*, const int*, int, cudaStream_t, ImgShare);
template void launch_ln_silu_bwd_t<float>(GP<float>, GP<float>, AP<float>, const float*, const float*, int, cudaStream_t,
                                          ImgShare, bool);
template void launch_ln_silu_bwd_t<D1>(GP<D1>, GP<D1>, AP<D1>, const float*, const float*, int, cudaStream_t, ImgShare, bool);

}  // namespace umab
