// K6/K8/K10 and their adjoints: the memory-bound halves of the Edgewise block.
//
//   gather_rotate_scale   x[src], x[tgt] -> Wigner rotate (to m-primary) -> x radial weights
//                         -> A0 [E,768] A1 [E,2,512] A2 [E,2,256]   (inputs of the SO(2) conv-1 GEMMs)
//   combine_gate          conv-1 outputs Y0/Y1/Y2 -> complex combine + gate activation
//                         -> B0 [E,384] B1 [E,2,256] B2 [E,2,128]   (inputs of the conv-2 GEMMs)
//   rotate_back_reduce    conv-2 outputs Z0/Z1/Z2 -> combine -> x envelope -> rotate back
//                         -> segmented sum over the CSR row of each target (no atomics)
// and the matching backward kernels.  fairchem: Edgewise.forward / SO2_Convolution /
// GateActivation / EdgeDegreeEmbedding, reached from the reference through predict_unit.predict
// (pdb2reaction/uma_pysis.py:385).  All kernels: one warp per edge (or per target node for the
// segmented reductions), one float4 = 4 of the 128 channels per lane, fully coalesced 512 B rows.
// Twin: oracle/staged.py (same function names).
#include "common.cuh"

namespace umab {

namespace {

struct WigReg { float d1[3][3]; float d2[5][5]; };

__device__ __forceinline__ WigReg load_wig(const float* __restrict__ wig, long long e) {
    WigReg w;
    const float4* p = reinterpret_cast<const float4*>(wig + e * WIG);
    float t[36];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        float4 v = __ldg(p + i);
        t[i * 4 + 0] = v.x; t[i * 4 + 1] = v.y; t[i * 4 + 2] = v.z; t[i * 4 + 3] = v.w;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) w.d1[a][b] = t[a * 3 + b];
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int b = 0; b < 5; ++b) w.d2[a][b] = t[9 + a * 5 + b];
    return w;
}

// y = D x  (l-primary rows, block diagonal)
__device__ __forceinline__ void rot_fwd(const WigReg& w, const float4* x, float4* y) {
    y[0] = x[0];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float4 s = f4zero();
#pragma unroll
        for (int b = 0; b < 3; ++b) f4fma(s, w.d1[a][b], x[1 + b]);
        y[1 + a] = s;
    }
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        float4 s = f4zero();
#pragma unroll
        for (int b = 0; b < 5; ++b) f4fma(s, w.d2[a][b], x[4 + b]);
        y[4 + a] = s;
    }
}
// y = D^T x
__device__ __forceinline__ void rot_bwd(const WigReg& w, const float4* x, float4* y) {
    y[0] = x[0];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float4 s = f4zero();
#pragma unroll
        for (int b = 0; b < 3; ++b) f4fma(s, w.d1[b][a], x[1 + b]);
        y[1 + a] = s;
    }
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        float4 s = f4zero();
#pragma unroll
        for (int b = 0; b < 5; ++b) f4fma(s, w.d2[b][a], x[4 + b]);
        y[4 + a] = s;
    }
}

// layout of m-primary row k inside the conv-1 input buffers / radial vector (floats)
__device__ __forceinline__ constexpr int a_buf(int k) { return k < 3 ? 0 : (k < 7 ? 1 : 2); }
__device__ __forceinline__ constexpr int a_off(int k) {   // offset inside the edge record of its buffer
    return k < 3 ? k * 256 : (k < 7 ? (k - 3) * 256 : (k - 7) * 256);
}
__device__ __forceinline__ constexpr int r_off(int k) {
    return k < 3 ? k * 256 : (k == 3 || k == 5 ? 768 : (k == 4 || k == 6 ? 1024 : 1280));
}

// totals of r[t] over the warp land in lane t (31 shuffles instead of 160)
__device__ __forceinline__ float warp_transpose_sum32(float (&r)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            float send = upper ? r[i] : r[i + o];
            float keep = upper ? r[i + o] : r[i];
            r[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return r[0];
}

// partial (this lane's 4 channels) outer products  acc[(b,a)] += sum_c zl[b][c] g[a][c]  restricted to
// the l=1 and l=2 blocks; slot order = Wigner record order (D1 row-major, then D2 row-major)
__device__ __forceinline__ void wig_outer_acc(float (&acc)[34], const float4* zl, const float4* g) {
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int a = 0; a < 3; ++a) acc[b * 3 + a] += f4dot(zl[1 + b], g[1 + a]);
#pragma unroll
    for (int b = 0; b < 5; ++b)
#pragma unroll
        for (int a = 0; a < 5; ++a) acc[9 + b * 5 + a] += f4dot(zl[4 + b], g[4 + a]);
}

// warp-reduce the 34 partials and add them (times `scale`) into g_wig[e]
__device__ __forceinline__ void wig_grad_commit(float (&acc)[34], float scale, float* g_wig, long long e, int lane) {
    float r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = acc[i];
    float tot = warp_transpose_sum32(r, lane);
    float t32 = warp_sum(acc[32]);
    float t33 = warp_sum(acc[33]);
    float* p = g_wig + e * WIG;
    p[lane] += scale * tot;
    if (lane == 0) p[32] += scale * t32;
    if (lane == 1) p[33] += scale * t33;
}

// ------------------------------------------------------------------ gather + rotate + radial scale
__global__ void __launch_bounds__(256)
gather_rotate_scale_kernel(const float* __restrict__ x, const int* __restrict__ src, const int* __restrict__ tgt,
                           const float* __restrict__ wig, const float* __restrict__ rad, long long e0, int n_e,
                           float* __restrict__ A0, float* __restrict__ A1, float* __restrict__ A2) {
    const int el = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (el >= n_e) return;
    const long long e = e0 + el;
    const WigReg w = load_wig(wig, e);
    const float* rp = rad + (long long)el * RAD1 + lane * 4;
    float* const bufs[3] = {A0 + (long long)el * 768, A1 + (long long)el * 1024, A2 + (long long)el * 512};
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int node = half == 0 ? src[e] : tgt[e];
        const float* xp = x + (long long)node * (9 * C) + lane * 4;
        float4 xr[9], yl[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) xr[r] = ld4(xp + r * C);
        rot_fwd(w, xr, yl);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            float4 rv = ld4(rp + r_off(k) + half * C);
            st4(bufs[a_buf(k)] + a_off(k) + half * C + lane * 4, f4mul(yl[to_m(k)], rv));
        }
    }
}

// adjoint: one warp per TARGET node of the chunk, looping over its CSR row.
//   g_rad (may alias rad) [E,1536];  G[e] = dL/dx[src] contribution [9,128] (l-primary);
//   g_x[i] = sum over the row of the target-half contributions;  g_wig[e] += ...
__global__ void __launch_bounds__(256)
gather_rotate_bwd_kernel(const float* __restrict__ x, const int* __restrict__ row_ptr, const int* __restrict__ src,
                         const float* __restrict__ wig, const float* rad, long long e0, int node0, int n_nodes,
                         const float* __restrict__ gA0, const float* __restrict__ gA1, const float* __restrict__ gA2,
                         float* g_rad, float* __restrict__ G, float* __restrict__ g_x, float* __restrict__ g_wig) {
    const int nl = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (nl >= n_nodes) return;
    const int i = node0 + nl;
    float4 acc_i[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) acc_i[r] = f4zero();
    const float* xi_p = x + (long long)i * (9 * C) + lane * 4;

    for (long long e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
        const long long el = e - e0;
        const WigReg w = load_wig(wig, e);
        const float* rp = rad + el * RAD1 + lane * 4;
        float* grp = g_rad + el * RAD1 + lane * 4;
        const float* const gbufs[3] = {gA0 + el * 768, gA1 + el * 1024, gA2 + el * 512};
        float wacc[34];
#pragma unroll
        for (int q = 0; q < 34; ++q) wacc[q] = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const float* xp = half == 0 ? x + (long long)src[e] * (9 * C) + lane * 4 : xi_p;
            float4 xr[9], yl[9], gml[9];
            float4 g_rad_v[6];      // radial groups 0,1,2 (m=0 rows), 3,4 (m=1: l=1,2), 5 (m=2)
#pragma unroll
            for (int r = 0; r < 9; ++r) xr[r] = ld4(xp + r * C);
            rot_fwd(w, xr, yl);
#pragma unroll
            for (int q = 0; q < 6; ++q) g_rad_v[q] = f4zero();
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                float4 ga = ld4(gbufs[a_buf(k)] + a_off(k) + half * C + lane * 4);
                float4 rv = ld4(rp + r_off(k) + half * C);
                const int grp_id = k < 3 ? k : (k == 3 || k == 5 ? 3 : (k == 4 || k == 6 ? 4 : 5));
                g_rad_v[grp_id] = f4add(g_rad_v[grp_id], f4mul(ga, yl[to_m(k)]));
                gml[to_m(k)] = f4mul(ga, rv);
            }
            // every read of this lane's rad[e] columns of this half is done: g_rad may alias rad
#pragma unroll
            for (int q = 0; q < 3; ++q) st4(grp + q * 256 + half * C, g_rad_v[q]);
            st4(grp + 768 + half * C, g_rad_v[3]);
            st4(grp + 1024 + half * C, g_rad_v[4]);
            st4(grp + 1280 + half * C, g_rad_v[5]);
            // dL/dD[a][b] += sum_c gml[a][c] x[b][c]
            wig_outer_acc(wacc, gml, xr);
            float4 gx[9];
            rot_bwd(w, gml, gx);
            if (half == 0) {
                float* gp = G + e * (9 * C) + lane * 4;
#pragma unroll
                for (int r = 0; r < 9; ++r) st4(gp + r * C, gx[r]);
            } else {
#pragma unroll
                for (int r = 0; r < 9; ++r) acc_i[r] = f4add(acc_i[r], gx[r]);
            }
        }
        wig_grad_commit(wacc, 1.0f, g_wig, e, lane);
    }
    float* op = g_x + (long long)i * (9 * C) + lane * 4;
#pragma unroll
    for (int r = 0; r < 9; ++r) st4(op + r * C, acc_i[r]);
}

// g_x[j] += sum over out-edges of j of G[e]
__global__ void __launch_bounds__(256)
source_reduce_kernel(const float* __restrict__ G, const int* __restrict__ sptr, const int* __restrict__ sedge,
                     int n_nodes, float* __restrict__ g_x) {
    const int j = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (j >= n_nodes) return;
    float4 acc[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) acc[r] = f4zero();
    for (int k = sptr[j]; k < sptr[j + 1]; ++k) {
        const float* gp = G + (long long)sedge[k] * (9 * C) + lane * 4;
#pragma unroll
        for (int r = 0; r < 9; ++r) acc[r] = f4add(acc[r], ld4(gp + r * C));
    }
    float* op = g_x + (long long)j * (9 * C) + lane * 4;
#pragma unroll
    for (int r = 0; r < 9; ++r) st4(op + r * C, f4add(ld4(op + r * C), acc[r]));
}

// ------------------------------------------------------------------ combine + gate (between the convs)
__global__ void __launch_bounds__(256)
combine_gate_fwd_kernel(const float* __restrict__ Y0, const float* __restrict__ Y1, const float* __restrict__ Y2,
                        int n_e, float* __restrict__ B0, float* __restrict__ B1, float* __restrict__ B2) {
    const int el = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (el >= n_e) return;
    const float* y0 = Y0 + (long long)el * 640 + lane * 4;
    const float* y1 = Y1 + (long long)el * 1024 + lane * 4;
    const float* y2 = Y2 + (long long)el * 512 + lane * 4;
    float4 g[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        float4 v = ld4(y0 + l * 128);
        g[l] = make_float4(sigmoidf_(v.x), sigmoidf_(v.y), sigmoidf_(v.z), sigmoidf_(v.w));
    }
    float* b0 = B0 + (long long)el * 384 + lane * 4;
    float4 t0 = ld4(y0 + 256);
    st4(b0, make_float4(siluf_(t0.x), siluf_(t0.y), siluf_(t0.z), siluf_(t0.w)));
    st4(b0 + 128, f4mul(ld4(y0 + 384), g[0]));
    st4(b0 + 256, f4mul(ld4(y0 + 512), g[1]));
    float* b1 = B1 + (long long)el * 512 + lane * 4;
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        float4 o_r = f4sub(ld4(y1 + l * 128), ld4(y1 + 512 + 256 + l * 128));
        float4 o_i = f4add(ld4(y1 + 512 + l * 128), ld4(y1 + 256 + l * 128));
        st4(b1 + l * 128, f4mul(o_r, g[l]));
        st4(b1 + 256 + l * 128, f4mul(o_i, g[l]));
    }
    float* b2 = B2 + (long long)el * 256 + lane * 4;
    float4 p_r = f4sub(ld4(y2), ld4(y2 + 256 + 128));
    float4 p_i = f4add(ld4(y2 + 256), ld4(y2 + 128));
    st4(b2, f4mul(p_r, g[1]));
    st4(b2 + 128, f4mul(p_i, g[1]));
}

// gY* may alias Y*, (gB* are read-only)
__global__ void __launch_bounds__(256)
combine_gate_bwd_kernel(const float* Y0, const float* Y1, const float* Y2, int n_e,
                        const float* __restrict__ gB0, const float* __restrict__ gB1, const float* __restrict__ gB2,
                        float* gY0, float* gY1, float* gY2) {
    const int el = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (el >= n_e) return;
    const float* y0 = Y0 + (long long)el * 640 + lane * 4;
    const float* y1 = Y1 + (long long)el * 1024 + lane * 4;
    const float* y2 = Y2 + (long long)el * 512 + lane * 4;
    const float* gb0 = gB0 + (long long)el * 384 + lane * 4;
    const float* gb1 = gB1 + (long long)el * 512 + lane * 4;
    const float* gb2 = gB2 + (long long)el * 256 + lane * 4;
    float4 sg[2], g_gate[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        float4 v = ld4(y0 + l * 128);
        sg[l] = make_float4(sigmoidf_(v.x), sigmoidf_(v.y), sigmoidf_(v.z), sigmoidf_(v.w));
    }
    float4 t0 = ld4(y0 + 256), t1 = ld4(y0 + 384), t2 = ld4(y0 + 512);
    float4 gb00 = ld4(gb0), gb01 = ld4(gb0 + 128), gb02 = ld4(gb0 + 256);
    g_gate[0] = f4mul(gb01, t1);
    g_gate[1] = f4mul(gb02, t2);
    float4 g_or[2], g_oi[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        float4 o_r = f4sub(ld4(y1 + l * 128), ld4(y1 + 512 + 256 + l * 128));
        float4 o_i = f4add(ld4(y1 + 512 + l * 128), ld4(y1 + 256 + l * 128));
        float4 br = ld4(gb1 + l * 128), bi = ld4(gb1 + 256 + l * 128);
        g_gate[l] = f4add(g_gate[l], f4add(f4mul(br, o_r), f4mul(bi, o_i)));
        g_or[l] = f4mul(br, sg[l]);
        g_oi[l] = f4mul(bi, sg[l]);
    }
    float4 p_r = f4sub(ld4(y2), ld4(y2 + 256 + 128));
    float4 p_i = f4add(ld4(y2 + 256), ld4(y2 + 128));
    float4 b2r = ld4(gb2), b2i = ld4(gb2 + 128);
    g_gate[1] = f4add(g_gate[1], f4add(f4mul(b2r, p_r), f4mul(b2i, p_i)));
    float4 g_pr = f4mul(b2r, sg[1]), g_pi = f4mul(b2i, sg[1]);

    // ---- all reads done; writes (possibly in place)
    float* o0 = gY0 + (long long)el * 640 + lane * 4;
    float* o1 = gY1 + (long long)el * 1024 + lane * 4;
    float* o2 = gY2 + (long long)el * 512 + lane * 4;
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        float4 s = sg[l], gg = g_gate[l];
        st4(o0 + l * 128, make_float4(gg.x * s.x * (1.f - s.x), gg.y * s.y * (1.f - s.y),
                                      gg.z * s.z * (1.f - s.z), gg.w * s.w * (1.f - s.w)));
    }
    st4(o0 + 256, make_float4(gb00.x * dsiluf_(t0.x), gb00.y * dsiluf_(t0.y), gb00.z * dsiluf_(t0.z), gb00.w * dsiluf_(t0.w)));
    st4(o0 + 384, f4mul(gb01, sg[0]));
    st4(o0 + 512, f4mul(gb02, sg[1]));
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        st4(o1 + l * 128, g_or[l]);
        st4(o1 + 256 + l * 128, g_oi[l]);
        st4(o1 + 512 + l * 128, g_oi[l]);
        st4(o1 + 512 + 256 + l * 128, f4scale(g_or[l], -1.f));
    }
    st4(o2, g_pr);
    st4(o2 + 128, g_pi);
    st4(o2 + 256, g_pi);
    st4(o2 + 256 + 128, f4scale(g_pr, -1.f));
}

// ------------------------------------------------------------------ rotate back + segmented reduce
// MODE 0: message rows from the conv-2 outputs Z0/Z1/Z2.  MODE 1: edge-degree embedding, rows 0..2
// from Z0 (= radial output [E,384]), rows 3..8 zero.
template <int MODE>
__device__ __forceinline__ void load_zl(const float* Z0, const float* Z1, const float* Z2, long long el, int lane, float4* zl) {
    const float* z0 = Z0 + el * 384 + lane * 4;
#pragma unroll
    for (int k = 0; k < 3; ++k) zl[to_m(k)] = ld4(z0 + k * 128);
    if (MODE == 0) {
        const float* z1 = Z1 + el * 1024 + lane * 4;
        const float* z2 = Z2 + el * 512 + lane * 4;
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            zl[to_m(3 + l)] = f4sub(ld4(z1 + l * 128), ld4(z1 + 512 + 256 + l * 128));
            zl[to_m(5 + l)] = f4add(ld4(z1 + 512 + l * 128), ld4(z1 + 256 + l * 128));
        }
        zl[to_m(7)] = f4sub(ld4(z2), ld4(z2 + 256 + 128));
        zl[to_m(8)] = f4add(ld4(z2 + 256), ld4(z2 + 128));
    } else {
#pragma unroll
        for (int k = 3; k < 9; ++k) zl[to_m(k)] = f4zero();
    }
}

template <int MODE>
__global__ void __launch_bounds__(256)
rotate_back_reduce_kernel(const float* __restrict__ Z0, const float* __restrict__ Z1, const float* __restrict__ Z2,
                          const int* __restrict__ row_ptr, const float* __restrict__ wig,
                          const float* __restrict__ env, float scale, long long e0, int node0, int n_nodes,
                          const float* base, float* out) {   // base may alias out
    const int nl = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (nl >= n_nodes) return;
    const int i = node0 + nl;
    float4 acc[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) acc[r] = f4zero();
    for (long long e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
        const WigReg w = load_wig(wig, e);
        float4 zl[9], y[9];
        load_zl<MODE>(Z0, Z1, Z2, e - e0, lane, zl);
        rot_bwd(w, zl, y);
        const float s = __ldg(env + e) * scale;
#pragma unroll
        for (int r = 0; r < 9; ++r) f4fma(acc[r], s, y[r]);
    }
    float* op = out + (long long)i * (9 * C) + lane * 4;
    if (base) {
        const float* bp = base + (long long)i * (9 * C) + lane * 4;
#pragma unroll
        for (int r = 0; r < 9; ++r) st4(op + r * C, f4add(ld4(bp + r * C), acc[r]));
    } else {
#pragma unroll
        for (int r = 0; r < 9; ++r) st4(op + r * C, acc[r]);
    }
}

// adjoint, one warp per edge.  gZ* may alias Z*.
template <int MODE>
__global__ void __launch_bounds__(256)
rotate_back_bwd_kernel(const float* Z0, const float* Z1, const float* Z2, const int* __restrict__ tgt,
                       const float* __restrict__ wig, const float* __restrict__ env, float scale,
                       long long e0, int n_e, const float* __restrict__ g_out,
                       float* gZ0, float* gZ1, float* gZ2, float* __restrict__ g_env, float* __restrict__ g_wig) {
    const int el = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (el >= n_e) return;
    const long long e = e0 + el;
    const WigReg w = load_wig(wig, e);
    float4 zl[9], g[9], t[9];
    load_zl<MODE>(Z0, Z1, Z2, el, lane, zl);
    const float* gp = g_out + (long long)tgt[e] * (9 * C) + lane * 4;
#pragma unroll
    for (int r = 0; r < 9; ++r) g[r] = ld4(gp + r * C);
    const float s = __ldg(env + e) * scale;
    // d/denv
    rot_bwd(w, zl, t);
    float part = 0.f;
#pragma unroll
    for (int r = 0; r < 9; ++r) part += f4dot(t[r], g[r]);
    part = warp_sum(part);
    if (lane == 0) g_env[e] += scale * part;
    // d/dD[b][a] = s * sum_c zl[b][c] g[a][c]
    float wacc[34];
#pragma unroll
    for (int q = 0; q < 34; ++q) wacc[q] = 0.f;
    wig_outer_acc(wacc, zl, g);
    wig_grad_commit(wacc, s, g_wig, e, lane);
    // d/dz (m-primary rows) = s * (D g)[to_m(k)]
    rot_fwd(w, g, t);
    float* o0 = gZ0 + (long long)el * 384 + lane * 4;
#pragma unroll
    for (int k = 0; k < 3; ++k) st4(o0 + k * 128, f4scale(t[to_m(k)], s));
    if (MODE == 0) {
        float* o1 = gZ1 + (long long)el * 1024 + lane * 4;
        float* o2 = gZ2 + (long long)el * 512 + lane * 4;
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            float4 g_or = f4scale(t[to_m(3 + l)], s), g_oi = f4scale(t[to_m(5 + l)], s);
            st4(o1 + l * 128, g_or);
            st4(o1 + 256 + l * 128, g_oi);
            st4(o1 + 512 + l * 128, g_oi);
            st4(o1 + 512 + 256 + l * 128, f4scale(g_or, -1.f));
        }
        float4 g_pr = f4scale(t[to_m(7)], s), g_pi = f4scale(t[to_m(8)], s);
        st4(o2, g_pr);
        st4(o2 + 128, g_pi);
        st4(o2 + 256, g_pi);
        st4(o2 + 256 + 128, f4scale(g_pr, -1.f));
    }
}

}  // namespace

void launch_gather_rotate_scale(const float* x, const int* src, const int* tgt, const float* wig, const float* rad,
                                long long e0, int n_e, float* A0, float* A1, float* A2, cudaStream_t st) {
    if (n_e <= 0) return;
    gather_rotate_scale_kernel<<<(n_e + 7) / 8, 256, 0, st>>>(x, src, tgt, wig, rad, e0, n_e, A0, A1, A2);
    UMAB_LAUNCH_CHECK();
}

void launch_gather_rotate_bwd(const float* x, const int* row_ptr, const int* src, const float* wig, const float* rad,
                              long long e0, int node0, int n_nodes, const float* gA0, const float* gA1,
                              const float* gA2, float* g_rad, float* G, float* g_x, float* g_wig, cudaStream_t st) {
    if (n_nodes <= 0) return;
    gather_rotate_bwd_kernel<<<(n_nodes + 7) / 8, 256, 0, st>>>(x, row_ptr, src, wig, rad, e0, node0, n_nodes,
                                                                gA0, gA1, gA2, g_rad, G, g_x, g_wig);
    UMAB_LAUNCH_CHECK();
}

void launch_source_reduce(const float* G, const int* sptr, const int* sedge, int n_nodes, float* g_x, cudaStream_t st) {
    if (n_nodes <= 0) return;
    source_reduce_kernel<<<(n_nodes + 7) / 8, 256, 0, st>>>(G, sptr, sedge, n_nodes, g_x);
    UMAB_LAUNCH_CHECK();
}

void launch_combine_gate_fwd(const float* Y0, const float* Y1, const float* Y2, int n_e, float* B0, float* B1,
                             float* B2, cudaStream_t st) {
    if (n_e <= 0) return;
    combine_gate_fwd_kernel<<<(n_e + 7) / 8, 256, 0, st>>>(Y0, Y1, Y2, n_e, B0, B1, B2);
    UMAB_LAUNCH_CHECK();
}

void launch_combine_gate_bwd(const float* Y0, const float* Y1, const float* Y2, int n_e, const float* gB0,
                             const float* gB1, const float* gB2, float* gY0, float* gY1, float* gY2, cudaStream_t st) {
    if (n_e <= 0) return;
    combine_gate_bwd_kernel<<<(n_e + 7) / 8, 256, 0, st>>>(Y0, Y1, Y2, n_e, gB0, gB1, gB2, gY0, gY1, gY2);
    UMAB_LAUNCH_CHECK();
}

void launch_rotate_back_reduce(int mode, const float* Z0, const float* Z1, const float* Z2, const int* row_ptr,
                               const float* wig, const float* env, float scale, long long e0, int node0,
                               int n_nodes, const float* base, float* out, cudaStream_t st) {
    if (n_nodes <= 0) return;
    dim3 grid((n_nodes + 7) / 8);
    if (mode == 0)
        rotate_back_reduce_kernel<0><<<grid, 256, 0, st>>>(Z0, Z1, Z2, row_ptr, wig, env, scale, e0, node0, n_nodes, base, out);
    else
        rotate_back_reduce_kernel<1><<<grid, 256, 0, st>>>(Z0, Z1, Z2, row_ptr, wig, env, scale, e0, node0, n_nodes, base, out);
    UMAB_LAUNCH_CHECK();
}

void launch_rotate_back_bwd(int mode, const float* Z0, const float* Z1, const float* Z2, const int* tgt,
                            const float* wig, const float* env, float scale, long long e0, int n_e,
                            const float* g_out, float* gZ0, float* gZ1, float* gZ2, float* g_env, float* g_wig,
                            cudaStream_t st) {
    if (n_e <= 0) return;
    dim3 grid((n_e + 7) / 8);
    if (mode == 0)
        rotate_back_bwd_kernel<0><<<grid, 256, 0, st>>>(Z0, Z1, Z2, tgt, wig, env, scale, e0, n_e, g_out, gZ0, gZ1, gZ2, g_env, g_wig);
    else
        rotate_back_bwd_kernel<1><<<grid, 256, 0, st>>>(Z0, Z1, Z2, tgt, wig, env, scale, e0, n_e, g_out, gZ0, gZ1, gZ2, g_env, g_wig);
    UMAB_LAUNCH_CHECK();
}

}  // namespace umab
