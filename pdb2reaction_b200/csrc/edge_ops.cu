// K6/K8/K10 and their adjoints: the memory-bound halves of the Edgewise block.
//
//   gather_rotate_scale   x[src], x[tgt] -> Wigner rotate (to m-primary) -> x radial weights
//                         -> A0 [E,768] A1 [E,(+m|-m) 2x512] A2 [E,2x256]   (inputs of the SO(2) conv-1 GEMMs)
//   combine_gate          conv-1 outputs Y0 [E,640] Y1 [E,(o_r|o_i) 2x256] Y2 [E,2x128] -> gate activation
//                         -> B0 [E,384] B1 [E,2x256] B2 [E,2x128]   (inputs of the conv-2 GEMMs)
//   rotate_back_reduce    conv-2 outputs Z0 [E,384] Z1 [E,2x256] Z2 [E,2x128] -> x envelope -> rotate back
//                         -> segmented sum over the CSR row of each target (no atomics)
// The (+m, -m) -> (real, imaginary) combination of SO2_m_Conv is folded into the m > 0 weights (complex block
// form, engine.prepare_engine_weights), so the m > 0 GEMMs emit [o_r | o_i] directly: half the conv output.
// and the matching backward kernels.  fairchem: Edgewise.forward / SO2_Convolution /
// GateActivation / EdgeDegreeEmbedding, reached from the reference through predict_unit.predict
// (pdb2reaction/uma_pysis.py:385).  All kernels: one warp per edge (or per target node for the
// segmented reductions), one float4 = 4 of the 128 channels per lane, fully coalesced 512 B rows.
// Twin: oracle/staged.py (same function names).
#include "dual.cuh"

namespace umab {

namespace {

// this warp's row and image: the float kernels keep exactly their plain one-warp-per-row mapping (and code)
#define UMAB_SHARE_ROW(WPC, N, ROW)                                                   \
    int ROW;                                                                          \
    int img = 0;                                                                      \
    if constexpr (std::is_same<S, float>::value) {                                    \
        ROW = blockIdx.x * (WPC) + threadIdx.x / 32;                                  \
        if (ROW >= (N)) return;                                                       \
    } else {                                                                          \
        long long row_;                                                               \
        if (!share_row<S>(sh, (WPC), (N), row_, img)) return;                         \
        ROW = (int)row_;                                                              \
    }

// every kernel is a template on the scalar type S: float (energy/forces) or D1 (value + tangent)
template <class S> struct WigReg { S d1[3][3]; S d2[5][5]; };

template <class S>
__device__ __forceinline__ WigReg<S> load_wig(GP<S> wig, long long e) {
    using V = typename VecOf<S>::type;
    WigReg<S> w;
    S t[36];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        V v = wig.ldg4(e * WIG + 4 * i);
        if constexpr (std::is_same<S, float>::value) {
            t[i * 4 + 0] = v.x; t[i * 4 + 1] = v.y; t[i * 4 + 2] = v.z; t[i * 4 + 3] = v.w;
        } else {
            t[i * 4 + 0] = D1{v.v.x, v.d.x}; t[i * 4 + 1] = D1{v.v.y, v.d.y};
            t[i * 4 + 2] = D1{v.v.z, v.d.z}; t[i * 4 + 3] = D1{v.v.w, v.d.w};
        }
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
template <class S, class V>
__device__ __forceinline__ void rot_fwd(const WigReg<S>& w, const V* x, V* y) {
    y[0] = x[0];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        V s = vzero<V>();
#pragma unroll
        for (int b = 0; b < 3; ++b) vfma(s, w.d1[a][b], x[1 + b]);
        y[1 + a] = s;
    }
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        V s = vzero<V>();
#pragma unroll
        for (int b = 0; b < 5; ++b) vfma(s, w.d2[a][b], x[4 + b]);
        y[4 + a] = s;
    }
}
// y = D^T x
template <class S, class V>
__device__ __forceinline__ void rot_bwd(const WigReg<S>& w, const V* x, V* y) {
    y[0] = x[0];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        V s = vzero<V>();
#pragma unroll
        for (int b = 0; b < 3; ++b) vfma(s, w.d1[b][a], x[1 + b]);
        y[1 + a] = s;
    }
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        V s = vzero<V>();
#pragma unroll
        for (int b = 0; b < 5; ++b) vfma(s, w.d2[b][a], x[4 + b]);
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

// ---- gradient with respect to the edge ROTATION as a torque (round 2; replaces the 34 outer-product entries dL/dD).
// The Wigner blocks depend on the edge direction only through the rotation R_e (D = D(R_e), D1 = R_e), and a change of
// R_e is an infinitesimal rotation  D -> (1 + sum_k dw_k J_k) D  with the so(3) generators J_k of the l = 1, 2 real
// harmonic blocks (basis of oracle/wigner.py: l = 1 -> (x, y, z); edge axis = y).  So for y = D x:  dL = sum_k dw_k
// <g_y, J_k y>, and for the rotate-back  out = s D^T z:  dL = - sum_k dw_k <g_z, J_k z>  with g_z = s D g_out.  The
// energy is invariant to the rotation about the edge axis (k = y: the roll angle), so only t_x and t_z are kept; the
// geometry adjoint turns them into dE/d(edge vector) = (t_z R_e[0] - t_x R_e[2]) / d  (geometry.cu).  Two warp
// reductions per edge and kernel instead of 34, 8 bytes of read-modify-write instead of 288.
//   J_x:  l=1 (1,2) = -1;  l=2 (0,1) = 1, (2,3) = -sqrt3, (3,4) = -1        (antisymmetric; (a,b) listed, (b,a) = -)
//   J_z:  l=1 (0,1) = -1;  l=2 (0,3) = -1, (1,2) = -sqrt3, (1,4) = -1
template <class V> struct TorqueAcc { V x1, xs, z1, zs; };      // unit-weight and sqrt3-weight terms of t_x, t_z
template <class V>
__device__ __forceinline__ TorqueAcc<V> torque_zero() { return {vzero<V>(), vzero<V>(), vzero<V>(), vzero<V>()}; }
// acc += (this lane's 4 channels of)  g^T J_k v ;  g, v: 9 l-primary rows
template <class V>
__device__ __forceinline__ void torque_acc(TorqueAcc<V>& t, const V* g, const V* v) {
    const V* g1 = g + 1; const V* v1 = v + 1; const V* g2 = g + 4; const V* v2 = v + 4;
    // t_x
    vfmav(t.x1, g1[2], v1[1]); vfnmav(t.x1, g1[1], v1[2]);
    vfmav(t.x1, g2[0], v2[1]); vfnmav(t.x1, g2[1], v2[0]);
    vfmav(t.x1, g2[4], v2[3]); vfnmav(t.x1, g2[3], v2[4]);
    vfmav(t.xs, g2[3], v2[2]); vfnmav(t.xs, g2[2], v2[3]);
    // t_z
    vfmav(t.z1, g1[1], v1[0]); vfnmav(t.z1, g1[0], v1[1]);
    vfmav(t.z1, g2[3], v2[0]); vfnmav(t.z1, g2[0], v2[3]);
    vfmav(t.z1, g2[4], v2[1]); vfnmav(t.z1, g2[1], v2[4]);
    vfmav(t.zs, g2[2], v2[1]); vfnmav(t.zs, g2[1], v2[2]);
}
// old values of the per-edge torque record [E, 4] = (t_x, t_z, 0, 0), requested at the top of the edge's work
template <class S>
__device__ __forceinline__ S torque_load(GP<S> g_tau, long long e, int lane) {
    return lane < 2 ? g_tau.ld(e * 4 + lane) : cst<S>(0.f);
}
// warp-reduce and add `scale` times the torque into g_tau[e]
template <class S, class V>
__device__ __forceinline__ void torque_commit(const TorqueAcc<V>& t, S scale, S old, GP<S> g_tau, long long e, int lane) {
    constexpr float SQ3 = 1.7320508075688772f;
    // explicit fused multiply-adds: the open-chunk and the closed-chunk kernels must round identically whatever the
    // compiler would contract on its own (test_closed_chunks_equal_open_chunks compares them bit for bit)
    const S tx = warp_sum(s_fma(cst<S>(SQ3), vhsum(t.xs), vhsum(t.x1)));
    const S tz = warp_sum(s_fma(cst<S>(SQ3), vhsum(t.zs), vhsum(t.z1)));
    if (lane < 2) g_tau.st(e * 4 + lane, s_fma(scale, lane == 0 ? tx : tz, old));
}

// scalars [LO, LO+N) of the Wigner record of edge e (D1 = 0..8, D2 = 9..33)
template <class S, int LO, int N>
__device__ __forceinline__ void load_wig_part(GP<S> wig, long long e, S* out) {
    using V = typename VecOf<S>::type;
    constexpr int Q0 = LO / 4, Q1 = (LO + N + 3) / 4;
#pragma unroll
    for (int q = Q0; q < Q1; ++q) {
        const V v = wig.ldg4(e * WIG + 4 * q);
        S t[4];
        if constexpr (std::is_same<S, float>::value) {
            t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
        } else {
            t[0] = D1{v.v.x, v.d.x}; t[1] = D1{v.v.y, v.d.y}; t[2] = D1{v.v.z, v.d.z}; t[3] = D1{v.v.w, v.d.w};
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int idx = 4 * q + jj - LO;
            if (idx >= 0 && idx < N) out[idx] = t[jj];
        }
    }
}
// l-primary coefficient index -> m-primary row (inverse of to_m)
__device__ __forceinline__ constexpr int from_l(int l) {
    constexpr int t[9] = {0, 5, 1, 3, 8, 6, 2, 4, 7};
    return t[l];
}

// ------------------------------------------------------------------ gather + rotate + radial scale
// Processed one l block at a time (l = 0, then the 3x3, then the 5x5 Wigner block) so that only one block of
// Wigner scalars and node rows is live: the kernel stays inside 80 registers (3 CTAs per SM) with the bf16
// hi/lo operand stores.
// (a 128-register cap made the dual-number rotate_back_* / combine_gate_bwd slower; for this kernel see GRS_D1_* below)
#ifndef GRS_D1_WARPS
// dual-number instantiation of gather_rotate_scale: warps per CTA, CTAs per SM.  Measured in a shared-base batch of 40
// Hessian columns: 8 x 1 (194 registers) 32.0 ms, 8 x 2 (128 registers, 248 B of spills) 27.9-29.4 ms, 4 x 3 (168 registers,
// 48 B of spills, 12 warps per SM) 25.1 ms
#define GRS_D1_WARPS 4
#define GRS_D1_MINB 3
#endif
template <class S> constexpr int grs_warps() { return std::is_same<S, float>::value ? 8 : GRS_D1_WARPS; }
template <class S>
__global__ void __launch_bounds__(grs_warps<S>() * 32, std::is_same<S, float>::value ? 3 : GRS_D1_MINB)
gather_rotate_scale_kernel(GP<S> x, const int* __restrict__ src, const int* __restrict__ tgt, GP<S> wig, GP<S> rad,
                           long long e0, int n_e, AP<S> A0, AP<S> A1, AP<S> A2, ImgShare sh) {
    using V = typename VecOf<S>::type;
    UMAB_SHARE_ROW(grs_warps<S>(), n_e, el);
    const int lane = threadIdx.x % 32;
    if (img) rad = rad.vback((long long)img * sh.rows * RAD1);
    const long long e = e0 + el;
    const long long rp = (long long)el * RAD1 + lane * 4;
    const long long i0 = (long long)el * 768 + lane * 4, i1 = (long long)el * 1024 + lane * 4,
                    i2 = (long long)el * 512 + lane * 4;
    // the l blocks below are latency-serialised phases: pull every row they will read into L2 up front
    const int node_s = src[e], node_t = tgt[e];
#pragma unroll
    for (int q = 0; q < RAD1 / 128; ++q) rad.prefetch(rp + q * 128);
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        x.prefetch((long long)node_s * (9 * C) + r * C + lane * 4);
        x.prefetch((long long)node_t * (9 * C) + r * C + lane * 4);
    }
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        const int node = half == 0 ? node_s : node_t;
        const long long xp = (long long)node * (9 * C) + lane * 4;
        // m-primary row k of the rotated message: times its radial weight, into the operand buffer of its m block
        auto put = [&](int k, V y) {
            const V o = vmul(y, rad.ldg4(rp + r_off(k) + half * C));
            if (a_buf(k) == 0) A0.st4(i0 + a_off(k) + half * C, o);
            else if (a_buf(k) == 1) A1.st4(i1 + a_off(k) + half * C, o);
            else A2.st4(i2 + a_off(k) + half * C, o);
        };
        put(from_l(0), x.ldg4(xp));
        {
            S d[9];
            load_wig_part<S, 0, 9>(wig, e, d);
            V xr[3];
#pragma unroll
            for (int b = 0; b < 3; ++b) xr[b] = x.ldg4(xp + (1 + b) * C);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                V y = vzero<V>();
#pragma unroll
                for (int b = 0; b < 3; ++b) vfma(y, d[a * 3 + b], xr[b]);
                put(from_l(1 + a), y);
            }
        }
        {
            S d[25];
            load_wig_part<S, 9, 25>(wig, e, d);
            V xr[5];
#pragma unroll
            for (int b = 0; b < 5; ++b) xr[b] = x.ldg4(xp + (4 + b) * C);
#pragma unroll
            for (int a = 0; a < 5; ++a) {
                V y = vzero<V>();
#pragma unroll
                for (int b = 0; b < 5; ++b) vfma(y, d[a * 5 + b], xr[b]);
                put(from_l(4 + a), y);
            }
        }
    }
}

// adjoint: one warp per TARGET node of the chunk, looping over its CSR row.
//   g_rad [E,1536] (A operand of the radial adjoint GEMM; never aliases rad);  G[e] = dL/dx[src] contribution [9,128] (l-primary);
//   g_x[i] = sum over the row of the target-half contributions;  g_wig[e] += ...
template <class S, bool PL>
__global__ void __launch_bounds__(256)
gather_rotate_bwd_kernel(GP<S> x, const int* __restrict__ row_ptr, const int* __restrict__ src, GP<S> wig, GP<S> rad,
                         long long e0, int node0, int n_nodes, GP<S> gA0, GP<S> gA1, GP<S> gA2, AP<S> g_rad, GP<S> G,
                         GP<S> g_x, GP<S> g_wig) {
    using V = typename VecOf<S>::type;
    const int nl = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (nl >= n_nodes) return;
    const int i = node0 + nl;
    V acc_i[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) acc_i[r] = vzero<V>();
    const long long xi_p = (long long)i * (9 * C) + lane * 4;

    const long long e_end = row_ptr[i + 1];
    for (long long e = row_ptr[i]; e < e_end; ++e) {
        const long long el = e - e0;
        if (e + 1 < e_end) {
            // the loop is a chain of dependent DRAM round trips: pull the next edge's rows into L2 meanwhile
            const long long en = el + 1;
            if (lane < 9) wig.prefetch((e + 1) * WIG + lane * 4);
#pragma unroll
            for (int q = 0; q < RAD1 / 128; ++q) rad.prefetch(en * RAD1 + q * 128 + lane * 4);
#pragma unroll
            for (int q = 0; q < 6; ++q) gA0.prefetch(en * 768 + q * 128 + lane * 4);
#pragma unroll
            for (int q = 0; q < 8; ++q) gA1.prefetch(en * 1024 + q * 128 + lane * 4);
#pragma unroll
            for (int q = 0; q < 4; ++q) gA2.prefetch(en * 512 + q * 128 + lane * 4);
            const long long xn = (long long)src[e + 1] * (9 * C) + lane * 4;
#pragma unroll
            for (int r = 0; r < 9; ++r) x.prefetch(xn + r * C);
        }
        const WigReg<S> w = load_wig<S>(wig, e);
        const long long rp = el * RAD1 + lane * 4;
        const GP<S> gbufs[3] = {gA0 + el * 768, gA1 + el * 1024, gA2 + el * 512};
        // target half first, then the source half, each with its own g_wig commit: the same order of additions as
        // the split (closed-chunk) kernels below, so both forms give identical bits
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int half = 1 - hh;
            TorqueAcc<V> tq = torque_zero<V>();
            const S tau_old = torque_load<S>(g_wig, e, lane);
            const long long xp = half == 0 ? (long long)src[e] * (9 * C) + lane * 4 : xi_p;
            V xr[9], yl[9], gml[9];
            V g_rad_v[6];      // radial groups 0,1,2 (m=0 rows), 3,4 (m=1: l=1,2), 5 (m=2)
#pragma unroll
            for (int r = 0; r < 9; ++r) xr[r] = x.ldg4(xp + r * C);
            rot_fwd(w, xr, yl);
#pragma unroll
            for (int q = 0; q < 6; ++q) g_rad_v[q] = vzero<V>();
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                V ga = gbufs[a_buf(k)].ldg4(a_off(k) + half * C + lane * 4);
                V rv = rad.ldg4(rp + r_off(k) + half * C);
                const int grp_id = k < 3 ? k : (k == 3 || k == 5 ? 3 : (k == 4 || k == 6 ? 4 : 5));
                vfmav(g_rad_v[grp_id], ga, yl[to_m(k)]);           // explicit FMA (never left to the contraction heuristics)
                gml[to_m(k)] = k == 0 ? vmul_unfused(ga, rv) : vmul(ga, rv);   // row 0 is added as it is: see vmul_unfused
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) g_rad.template st4t<PL>(rp + q * 256 + half * C, g_rad_v[q]);
            g_rad.template st4t<PL>(rp + 768 + half * C, g_rad_v[3]);
            g_rad.template st4t<PL>(rp + 1024 + half * C, g_rad_v[4]);
            g_rad.template st4t<PL>(rp + 1280 + half * C, g_rad_v[5]);
            torque_acc(tq, gml, yl);                 // <g_y, J_k y>,  y = D x
            V gx[9];
            rot_bwd(w, gml, gx);
            if (half == 0) {
                const long long gp = e * (9 * C) + lane * 4;
#pragma unroll
                for (int r = 0; r < 9; ++r) G.st4(gp + r * C, gx[r]);
            } else {
#pragma unroll
                for (int r = 0; r < 9; ++r) acc_i[r] = vadd(acc_i[r], gx[r]);
            }
            torque_commit(tq, cst<S>(1.0f), tau_old, g_wig, e, lane);
        }
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) g_x.st4(xi_p + r * C, acc_i[r]);
}

// One edge of gather_rotate_bwd_half, one l block at a time: the SAME operations in the SAME order per accumulator as the
// all-at-once body below (identical bits), but only one Wigner block, its rows of y and of the gradient are live at a
// time.  Used by the dual-number instantiation, whose all-at-once body needs ~360 registers (1 KB of spills).
// The node's own rows are re-read per block (L1 hits) and the nine row accumulators live in shared memory
// (acc_sm: this lane's slots, row r at [(2 r) * 32] (value) and [(2 r + 1) * 32] (tangent)).
__device__ __forceinline__ void acc_add(float4* acc_sm, int r, D4 g) {
    acc_sm[(2 * r) * 32] = f4add(acc_sm[(2 * r) * 32], g.v);
    acc_sm[(2 * r + 1) * 32] = f4add(acc_sm[(2 * r + 1) * 32], g.d);
}
[[maybe_unused]] __device__ __forceinline__ void acc_add(float4* acc_sm, int r, float4 g) { acc_sm[r * 32] = f4add(acc_sm[r * 32], g); }
__device__ __forceinline__ void acc_get(const float4* acc_sm, int r, D4& out) { out = D4{acc_sm[(2 * r) * 32], acc_sm[(2 * r + 1) * 32]}; }
[[maybe_unused]] __device__ __forceinline__ void acc_get(const float4* acc_sm, int r, float4& out) { out = acc_sm[r * 32]; }
#ifndef UMAB_HALF_BLOCKED_FLOAT
// 1: the float instantiation uses the blocked body too.  Measured (same box, C4 step, gather_rotate_bwd family): plain
// 59.8 / 59.1 ms; blocked at 172 / 202 registers, 8 warps per SM 63.2 / 62.9 ms; blocked at 128 registers (4-24 B of
// spills), 16 warps per SM 62.1 / 61.5 ms -- the float kernel is HBM-bound and gains nothing from the occupancy, the
// shared-memory accumulators cost it 3 ms.  Same bits either way (test_gpu_parity green with the blocked build).
#define UMAB_HALF_BLOCKED_FLOAT 0
#endif
template <class S> constexpr bool half_blocked() { return !std::is_same<S, float>::value || UMAB_HALF_BLOCKED_FLOAT; }
template <class S> constexpr int half_planes() { return std::is_same<S, float>::value ? 1 : 2; }
template <int HALF, class S, bool PL, class V>
__device__ __forceinline__ void half_edge_blocked(GP<S> x, long long xi_p, GP<S> wig, long long e, const GP<S>* gbufs, GP<S> rad,
                                                  long long rp, int lane, AP<S> g_rad, float4* acc_sm, TorqueAcc<V>& tq) {
    auto ga_of = [&](int kk) { return gbufs[a_buf(kk)].ldg4(a_off(kk) + HALF * C + lane * 4); };
    auto rv_of = [&](int kk) { return rad.ldg4(rp + r_off(kk) + HALF * C); };
    asm volatile("" : "+l"(xi_p));       // the rows are re-read per edge on purpose: do not hoist them out of the edge loop
    {   // l = 0 (m-primary row 0): y = x, the product is added as it is (vmul_unfused)
        const V ga = ga_of(0), rv = rv_of(0);
        V g0 = vzero<V>();
        vfmav(g0, ga, x.ldg4(xi_p));
        g_rad.template st4t<PL>(rp + HALF * C, g0);
        acc_add(acc_sm, 0, vmul_unfused(ga, rv));
    }
    asm volatile("" ::: "memory");     // phase boundary: keeps the loads of the next block out of this one's live range
    {   // l = 1: m-primary rows 1, 3, 5 (ascending, as the all-at-once loop visits them)
        S d[9];
        load_wig_part<S, 0, 9>(wig, e, d);
        V y[3], gm[3];
        {
            V xr[3];
#pragma unroll
            for (int b = 0; b < 3; ++b) xr[b] = x.ldg4(xi_p + (1 + b) * C);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
            y[a] = vzero<V>();
#pragma unroll
            for (int b = 0; b < 3; ++b) vfma(y[a], d[a * 3 + b], xr[b]);
            }
        }
        V g1 = vzero<V>(), g3 = vzero<V>();
        {
            constexpr int ks[3] = {1, 3, 5};
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int kk = ks[j], a = to_m(kk) - 1;
                const V ga = ga_of(kk), rv = rv_of(kk);
                if (kk == 1) vfmav(g1, ga, y[a]); else vfmav(g3, ga, y[a]);
                gm[a] = vmul(ga, rv);
            }
        }
        g_rad.template st4t<PL>(rp + 256 + HALF * C, g1);
        g_rad.template st4t<PL>(rp + 768 + HALF * C, g3);
        vfmav(tq.x1, gm[2], y[1]); vfnmav(tq.x1, gm[1], y[2]);
        vfmav(tq.z1, gm[1], y[0]); vfnmav(tq.z1, gm[0], y[1]);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            V gx = vzero<V>();
#pragma unroll
            for (int b = 0; b < 3; ++b) vfma(gx, d[b * 3 + a], gm[b]);
            acc_add(acc_sm, 1 + a, gx);
        }
    }
    asm volatile("" ::: "memory");
    {   // l = 2: m-primary rows 2, 4, 6, 7, 8
        S d[25];
        load_wig_part<S, 9, 25>(wig, e, d);
        V y[5], gm[5];
        {
            V xr[5];
#pragma unroll
            for (int b = 0; b < 5; ++b) xr[b] = x.ldg4(xi_p + (4 + b) * C);
#pragma unroll
            for (int a = 0; a < 5; ++a) {
                y[a] = vzero<V>();
#pragma unroll
                for (int b = 0; b < 5; ++b) vfma(y[a], d[a * 5 + b], xr[b]);
            }
        }
        V g2 = vzero<V>(), g4 = vzero<V>(), g5 = vzero<V>();
        {
            constexpr int ks[5] = {2, 4, 6, 7, 8};
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int kk = ks[j], a = to_m(kk) - 4;
                const V ga = ga_of(kk), rv = rv_of(kk);
                if (kk == 2) vfmav(g2, ga, y[a]); else if (kk < 7) vfmav(g4, ga, y[a]); else vfmav(g5, ga, y[a]);
                gm[a] = vmul(ga, rv);
            }
        }
        g_rad.template st4t<PL>(rp + 512 + HALF * C, g2);
        g_rad.template st4t<PL>(rp + 1024 + HALF * C, g4);
        g_rad.template st4t<PL>(rp + 1280 + HALF * C, g5);
        vfmav(tq.x1, gm[0], y[1]); vfnmav(tq.x1, gm[1], y[0]);
        vfmav(tq.x1, gm[4], y[3]); vfnmav(tq.x1, gm[3], y[4]);
        vfmav(tq.xs, gm[3], y[2]); vfnmav(tq.xs, gm[2], y[3]);
        vfmav(tq.z1, gm[3], y[0]); vfnmav(tq.z1, gm[0], y[3]);
        vfmav(tq.z1, gm[4], y[1]); vfnmav(tq.z1, gm[1], y[4]);
        vfmav(tq.zs, gm[2], y[1]); vfnmav(tq.zs, gm[1], y[2]);
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            V gx = vzero<V>();
#pragma unroll
            for (int b = 0; b < 5; ++b) vfma(gx, d[b * 5 + a], gm[b]);
            acc_add(acc_sm, 4 + a, gx);
        }
    }
}

// ---- split form for CLOSED chunks (every out-edge of a node of the chunk lies inside the chunk: whole images).
// The source half of the adjoint is then reduced by SOURCE node as well, so the per-edge buffer G [E,9,128] and the
// source_reduce pass disappear (9.2 KB less HBM traffic per edge and layer, 4.6 KB less memory per edge), and the
// source rows need no gather: all out-edges of node j share x[j].
//   HALF = 1  one warp per target node i, over its in-edges  (CSR row):       g_x[i]  = sum of target-half terms
//   HALF = 0  one warp per source node j, over its out-edges (sedge list):    g_x[j] += sum of source-half terms
// Both add their share of dL/dD into g_wig[e]; HALF = 1 runs first, HALF = 0 second (fixed order: deterministic).
#ifndef UMAB_HALF_NW
#define UMAB_HALF_NW 8          // warps (nodes) per CTA of the half kernels
#define UMAB_HALF_MINB 1        // CTAs per SM the register budget is sized for
// measured on one box (tools/gpu_ab_libs.sh): 8 warps x 1 CTA (204 / 238 registers, no spills) 60.0 / 60.7 ms per C4 step;
// 6 x 2 and 4 x 3 (168 registers, 200-280 B of spills, 12 resident warps) 70.1 / 69.9 and 69.6 / 68.4 ms
#endif
#ifndef UMAB_HALF_MINB_D1
#define UMAB_HALF_MINB_D1 1     // the dual-number instantiation (Hessian columns)
#endif
template <int HALF, class S, bool PL>
__global__ void __launch_bounds__(UMAB_HALF_NW * 32, std::is_same<S, float>::value ? UMAB_HALF_MINB : UMAB_HALF_MINB_D1)
gather_rotate_bwd_half_kernel(GP<S> x, const int* __restrict__ ptr, const int* __restrict__ elist, GP<S> wig, GP<S> rad,
                              long long e0, int node0, int n_nodes, GP<S> gA0, GP<S> gA1, GP<S> gA2, AP<S> g_rad,
                              GP<S> g_x, GP<S> g_wig, ImgShare sh, int e_img) {
    using V = typename VecOf<S>::type;
    UMAB_SHARE_ROW(UMAB_HALF_NW, n_nodes, nl);       // sh.rows = nodes per image; e_img = edges per image
    const int lane = threadIdx.x % 32;
    if (img) {
        rad = rad.vback((long long)img * e_img * RAD1);
        gA0 = gA0.vback((long long)img * e_img * 768);
        gA1 = gA1.vback((long long)img * e_img * 1024);
        gA2 = gA2.vback((long long)img * e_img * 512);
    }
    const int i = node0 + nl;
    V acc_i[9], xr[9];
    const long long xi_p = (long long)i * (9 * C) + lane * 4;
    // dual numbers: row accumulators in shared memory, the node's rows re-read per block (half_edge_blocked)
    constexpr bool kBlocked = half_blocked<S>();
    extern __shared__ float4 half_acc_sm[];
    float4* const acc_sm = half_acc_sm + (threadIdx.x / 32) * (9 * half_planes<S>() * 32) + lane;
    if (kBlocked) {
#pragma unroll
        for (int q = 0; q < 9 * half_planes<S>(); ++q) acc_sm[q * 32] = f4zero();
    } else {
#pragma unroll
        for (int r = 0; r < 9; ++r) { acc_i[r] = vzero<V>(); xr[r] = x.ldg4(xi_p + r * C); }
    }

    const int k_end = ptr[i + 1];
    for (int k = ptr[i]; k < k_end; ++k) {
        const long long e = HALF == 1 ? (long long)k : (long long)elist[k];
        const long long el = e - e0;
        if (k + 1 < k_end) {
            const long long e_n = HALF == 1 ? (long long)k + 1 : (long long)elist[k + 1];
            const long long en = e_n - e0;
            if (lane < 9) wig.prefetch(e_n * WIG + lane * 4);
            else if (lane == 9) g_wig.prefetch(e_n * 4);
#pragma unroll
            for (int q = 0; q < RAD1 / 256; ++q) rad.prefetch(en * RAD1 + q * 256 + HALF * C + lane * 4);
#pragma unroll
            for (int q = 0; q < 3; ++q) gA0.prefetch(en * 768 + q * 256 + HALF * C + lane * 4);
#pragma unroll
            for (int q = 0; q < 4; ++q) gA1.prefetch(en * 1024 + q * 256 + HALF * C + lane * 4);
#pragma unroll
            for (int q = 0; q < 2; ++q) gA2.prefetch(en * 512 + q * 256 + HALF * C + lane * 4);
        }
        const S tau_old = torque_load<S>(g_wig, e, lane);
        const long long rp = el * RAD1 + lane * 4;
        const GP<S> gbufs[3] = {gA0 + el * 768, gA1 + el * 1024, gA2 + el * 512};
        TorqueAcc<V> tq = torque_zero<V>();
        if constexpr (kBlocked) {
            half_edge_blocked<HALF, S, PL, V>(x, xi_p, wig, e, gbufs, rad, rp, lane, g_rad, acc_sm, tq);
            torque_commit(tq, cst<S>(1.0f), tau_old, g_wig, e, lane);
            continue;
        }
        const WigReg<S> w = load_wig<S>(wig, e);
        V yl[9], gml[9];
        V g_rad_v[6];
        rot_fwd(w, xr, yl);
#pragma unroll
        for (int q = 0; q < 6; ++q) g_rad_v[q] = vzero<V>();
#pragma unroll
        for (int kk = 0; kk < 9; ++kk) {
            V ga = gbufs[a_buf(kk)].ldg4(a_off(kk) + HALF * C + lane * 4);
            V rv = rad.ldg4(rp + r_off(kk) + HALF * C);
            const int grp_id = kk < 3 ? kk : (kk == 3 || kk == 5 ? 3 : (kk == 4 || kk == 6 ? 4 : 5));
            vfmav(g_rad_v[grp_id], ga, yl[to_m(kk)]);          // explicit FMA (never left to the contraction heuristics)
            gml[to_m(kk)] = kk == 0 ? vmul_unfused(ga, rv) : vmul(ga, rv);   // row 0 is added as it is: see vmul_unfused
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) g_rad.template st4t<PL>(rp + q * 256 + HALF * C, g_rad_v[q]);
        g_rad.template st4t<PL>(rp + 768 + HALF * C, g_rad_v[3]);
        g_rad.template st4t<PL>(rp + 1024 + HALF * C, g_rad_v[4]);
        g_rad.template st4t<PL>(rp + 1280 + HALF * C, g_rad_v[5]);
        torque_acc(tq, gml, yl);                     // <g_y, J_k y>,  y = D x
        V gx[9];
        rot_bwd(w, gml, gx);
#pragma unroll
        for (int r = 0; r < 9; ++r) acc_i[r] = vadd(acc_i[r], gx[r]);
        torque_commit(tq, cst<S>(1.0f), tau_old, g_wig, e, lane);
    }
    if constexpr (kBlocked) {
#pragma unroll
        for (int r = 0; r < 9; ++r) acc_get(acc_sm, r, acc_i[r]);
    }
    if (HALF == 1) {
#pragma unroll
        for (int r = 0; r < 9; ++r) g_x.st4(xi_p + r * C, acc_i[r]);
    } else {
#pragma unroll
        for (int r = 0; r < 9; ++r) g_x.st4(xi_p + r * C, vadd(g_x.ldg4(xi_p + r * C), acc_i[r]));
    }
}

// g_x[j] += sum over out-edges of j of G[e]      (linear: the Hessian path runs it once per plane)
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
template <class S>
__global__ void __launch_bounds__(256, min_blocks<S>(4))
combine_gate_fwd_kernel(GP<S> Y0, GP<S> Y1, GP<S> Y2, int n_e, AP<S> B0, AP<S> B1, AP<S> B2, ImgShare sh) {
    using V = typename VecOf<S>::type;
    UMAB_SHARE_ROW(8, n_e, el);
    const int lane = threadIdx.x % 32;
    if (img) {
        Y0 = Y0.vback((long long)img * sh.rows * 640);
        Y1 = Y1.vback((long long)img * sh.rows * 512);
        Y2 = Y2.vback((long long)img * sh.rows * 256);
    }
    const long long y0 = (long long)el * 640 + lane * 4;
    const long long y1 = (long long)el * 512 + lane * 4;
    const long long y2 = (long long)el * 256 + lane * 4;
    V g[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) g[l] = vsigmoid(Y0.ldg4(y0 + l * 128));
    const long long b0 = (long long)el * 384 + lane * 4;
    B0.st4(b0, vsilu(Y0.ldg4(y0 + 256)));
    B0.st4(b0 + 128, vmul(Y0.ldg4(y0 + 384), g[0]));
    B0.st4(b0 + 256, vmul(Y0.ldg4(y0 + 512), g[1]));
    const long long b1 = (long long)el * 512 + lane * 4;
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        B1.st4(b1 + l * 128, vmul(Y1.ldg4(y1 + l * 128), g[l]));                  // o_r
        B1.st4(b1 + 256 + l * 128, vmul(Y1.ldg4(y1 + 256 + l * 128), g[l]));      // o_i
    }
    const long long b2 = (long long)el * 256 + lane * 4;
    B2.st4(b2, vmul(Y2.ldg4(y2), g[1]));
    B2.st4(b2 + 128, vmul(Y2.ldg4(y2 + 128), g[1]));
}

// Fused-gate path (float, tensor-core GEMMs): only the m = 0 part of the combine runs as a kernel -- it also leaves
// sigmoid(gate pre-activations) [n_e, 256] for the conv-1 m = +-1 / +-2 GEMMs, whose epilogues apply the gate and write
// B1 / B2 themselves (gemm_tc2.cu).  Same arithmetic, same bits as combine_gate_fwd_kernel.
__global__ void __launch_bounds__(256, 4)
gate_b0_kernel(GP<float> Y0, int n_e, AP<float> B0, float* __restrict__ sg) {
    const int el = blockIdx.x * 8 + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (el >= n_e) return;
    const long long y0 = (long long)el * 640 + lane * 4;
    float4 g[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        g[l] = vsigmoid(Y0.ldg4(y0 + l * 128));
        st4(sg + (long long)el * 256 + l * 128 + lane * 4, g[l]);
    }
    const long long b0 = (long long)el * 384 + lane * 4;
    B0.st4(b0, vsilu(Y0.ldg4(y0 + 256)));
    B0.st4(b0 + 128, vmul(Y0.ldg4(y0 + 384), g[0]));
    B0.st4(b0 + 256, vmul(Y0.ldg4(y0 + 512), g[1]));
}

// gY* (A operands of the conv-1 adjoint GEMMs) never alias Y*
template <class S>
__global__ void __launch_bounds__(256, min_blocks<S>(3))
combine_gate_bwd_kernel(GP<S> Y0, GP<S> Y1, GP<S> Y2, int n_e, GP<S> gB0, GP<S> gB1, GP<S> gB2, AP<S> gY0, AP<S> gY1,
                        AP<S> gY2, ImgShare sh) {
    using V = typename VecOf<S>::type;
    UMAB_SHARE_ROW(8, n_e, el);
    const int lane = threadIdx.x % 32;
    if (img) {
        Y0 = Y0.vback((long long)img * sh.rows * 640);
        Y1 = Y1.vback((long long)img * sh.rows * 512);
        Y2 = Y2.vback((long long)img * sh.rows * 256);
        gB0 = gB0.vback((long long)img * sh.rows * 384);
        gB1 = gB1.vback((long long)img * sh.rows * 512);
        gB2 = gB2.vback((long long)img * sh.rows * 256);
    }
    const long long y0 = (long long)el * 640 + lane * 4;
    const long long y1 = (long long)el * 512 + lane * 4;
    const long long y2 = (long long)el * 256 + lane * 4;
    const long long gb0 = (long long)el * 384 + lane * 4;
    const long long gb1 = (long long)el * 512 + lane * 4;
    const long long gb2 = (long long)el * 256 + lane * 4;
    V sg[2], g_gate[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) sg[l] = vsigmoid(Y0.ldg4(y0 + l * 128));
    V t0 = Y0.ldg4(y0 + 256), t1 = Y0.ldg4(y0 + 384), t2 = Y0.ldg4(y0 + 512);
    V gb00 = gB0.ldg4(gb0), gb01 = gB0.ldg4(gb0 + 128), gb02 = gB0.ldg4(gb0 + 256);
    g_gate[0] = vmul(gb01, t1);
    g_gate[1] = vmul(gb02, t2);
    V g_or[2], g_oi[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        V o_r = Y1.ldg4(y1 + l * 128);
        V o_i = Y1.ldg4(y1 + 256 + l * 128);
        V br = gB1.ldg4(gb1 + l * 128), bi = gB1.ldg4(gb1 + 256 + l * 128);
        g_gate[l] = vadd(g_gate[l], vadd(vmul(br, o_r), vmul(bi, o_i)));
        g_or[l] = vmul(br, sg[l]);
        g_oi[l] = vmul(bi, sg[l]);
    }
    V p_r = Y2.ldg4(y2);
    V p_i = Y2.ldg4(y2 + 128);
    V b2r = gB2.ldg4(gb2), b2i = gB2.ldg4(gb2 + 128);
    g_gate[1] = vadd(g_gate[1], vadd(vmul(b2r, p_r), vmul(b2i, p_i)));
    V g_pr = vmul(b2r, sg[1]), g_pi = vmul(b2i, sg[1]);

#pragma unroll
    for (int l = 0; l < 2; ++l) {
        // sigma'(y) = s (1 - s)
        V s = sg[l];
        V ds = vsub(s, vmul(s, s));
        gY0.st4(y0 + l * 128, vmul(g_gate[l], ds));
    }
    gY0.st4(y0 + 256, vmul(gb00, vdsilu(t0)));
    gY0.st4(y0 + 384, vmul(gb01, sg[0]));
    gY0.st4(y0 + 512, vmul(gb02, sg[1]));
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        gY1.st4(y1 + l * 128, g_or[l]);
        gY1.st4(y1 + 256 + l * 128, g_oi[l]);
    }
    gY2.st4(y2, g_pr);
    gY2.st4(y2 + 128, g_pi);
}

// ------------------------------------------------------------------ rotate back + segmented reduce
// MODE 0: message rows from the conv-2 outputs Z0/Z1/Z2.  MODE 1: edge-degree embedding, rows 0..2
// from Z0 (= radial output [E,384]), rows 3..8 zero.
template <int MODE, class S, class V>
__device__ __forceinline__ void load_zl(GP<S> Z0, GP<S> Z1, GP<S> Z2, long long el, int lane, V* zl) {
    const long long z0 = el * 384 + lane * 4;
#pragma unroll
    for (int k = 0; k < 3; ++k) zl[to_m(k)] = Z0.ldg4(z0 + k * 128);
    if (MODE == 0) {
        const long long z1 = el * 512 + lane * 4;
        const long long z2 = el * 256 + lane * 4;
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            zl[to_m(3 + l)] = Z1.ldg4(z1 + l * 128);
            zl[to_m(5 + l)] = Z1.ldg4(z1 + 256 + l * 128);
        }
        zl[to_m(7)] = Z2.ldg4(z2);
        zl[to_m(8)] = Z2.ldg4(z2 + 128);
    } else {
#pragma unroll
        for (int k = 3; k < 9; ++k) zl[to_m(k)] = vzero<V>();
    }
}

template <int MODE, class S>
__global__ void __launch_bounds__(256, min_blocks<S>(2))
rotate_back_reduce_kernel(GP<S> Z0, GP<S> Z1, GP<S> Z2, const int* __restrict__ row_ptr, GP<S> wig, GP<S> env,
                          float scale, long long e0, int node0, int n_nodes, GP<S> base, GP<S> out,   // base may alias out
                          ImgShare sh, int e_img) {
    using V = typename VecOf<S>::type;
    UMAB_SHARE_ROW(8, n_nodes, nl);                  // sh.rows = nodes per image; e_img = edges per image
    const int lane = threadIdx.x % 32;
    if (img) {
        Z0 = Z0.vback((long long)img * e_img * 384);
        if (MODE == 0) {
            Z1 = Z1.vback((long long)img * e_img * 512);
            Z2 = Z2.vback((long long)img * e_img * 256);
        }
    }
    const int i = node0 + nl;
    V acc[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) acc[r] = vzero<V>();
    const long long e_end = row_ptr[i + 1];
    for (long long e = row_ptr[i]; e < e_end; ++e) {
        if (e + 1 < e_end) {
            const long long en = e + 1 - e0;
            if (lane < 9) wig.prefetch((e + 1) * WIG + lane * 4);
#pragma unroll
            for (int q = 0; q < 3; ++q) Z0.prefetch(en * 384 + q * 128 + lane * 4);
            if (MODE == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) Z1.prefetch(en * 512 + q * 128 + lane * 4);
#pragma unroll
                for (int q = 0; q < 2; ++q) Z2.prefetch(en * 256 + q * 128 + lane * 4);
            }
        }
        const WigReg<S> w = load_wig<S>(wig, e);
        V zl[9], y[9];
        load_zl<MODE, S, V>(Z0, Z1, Z2, e - e0, lane, zl);
        rot_bwd(w, zl, y);
        const S s = env.ldg(e) * scale;
#pragma unroll
        for (int r = 0; r < 9; ++r) vfma(acc[r], s, y[r]);
    }
    const long long op = (long long)i * (9 * C) + lane * 4;
    if (base) {
#pragma unroll
        for (int r = 0; r < 9; ++r) out.st4(op + r * C, vadd(base.ld4(op + r * C), acc[r]));
    } else {
#pragma unroll
        for (int r = 0; r < 9; ++r) out.st4(op + r * C, acc[r]);
    }
}

// adjoint, one warp per edge.  gZ* (A operands of the conv-2 adjoint GEMMs) never alias Z*.
template <int MODE, class S, bool PL>
__global__ void __launch_bounds__(256, min_blocks<S>(2))
rotate_back_bwd_kernel(GP<S> Z0, GP<S> Z1, GP<S> Z2, const int* __restrict__ tgt, GP<S> wig, GP<S> env, float scale,
                       long long e0, int n_e, GP<S> g_out, AP<S> gZ0, AP<S> gZ1, AP<S> gZ2, GP<S> g_env, GP<S> g_wig,
                       ImgShare sh) {
    using V = typename VecOf<S>::type;
    UMAB_SHARE_ROW(8, n_e, el);
    const int lane = threadIdx.x % 32;
    if (img) {
        Z0 = Z0.vback((long long)img * sh.rows * 384);
        if (MODE == 0) {
            Z1 = Z1.vback((long long)img * sh.rows * 512);
            Z2 = Z2.vback((long long)img * sh.rows * 256);
        }
    }
    const long long e = e0 + el;
    // the read-modify-write operands first: their DRAM round trips overlap everything below
    const S tau_old = torque_load<S>(g_wig, e, lane);
    S genv_old = cst<S>(0.f);
    if (lane == 0) genv_old = g_env.ld(e);
    const WigReg<S> w = load_wig<S>(wig, e);
    V zl[9], g[9], t[9];
    load_zl<MODE, S, V>(Z0, Z1, Z2, el, lane, zl);
    const long long gp = (long long)tgt[e] * (9 * C) + lane * 4;
#pragma unroll
    for (int r = 0; r < 9; ++r) g[r] = g_out.ldg4(gp + r * C);
    const S s = env.ldg(e) * scale;
    // t = D g serves all three outputs:  d/dz = s t (m-primary rows),  d/denv = scale <D^T z, g> = scale <z, t>,
    // torque = - s <t, J_k z>
    rot_fwd(w, g, t);
    V pacc = vzero<V>();
#pragma unroll
    for (int r = 0; r < 9; ++r) vfmav(pacc, t[r], zl[r]);
    S part = warp_sum(vhsum(pacc));
    if (lane == 0) g_env.st(e, genv_old + part * scale);
    const long long o0 = (long long)el * 384 + lane * 4;
#pragma unroll
    for (int k = 0; k < 3; ++k) gZ0.template st4t<PL>(o0 + k * 128, vscale(t[to_m(k)], s));
    if (MODE == 0) {
        const long long o1 = (long long)el * 512 + lane * 4;
        const long long o2 = (long long)el * 256 + lane * 4;
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            gZ1.template st4t<PL>(o1 + l * 128, vscale(t[to_m(3 + l)], s));
            gZ1.template st4t<PL>(o1 + 256 + l * 128, vscale(t[to_m(5 + l)], s));
        }
        gZ2.template st4t<PL>(o2, vscale(t[to_m(7)], s));
        gZ2.template st4t<PL>(o2 + 128, vscale(t[to_m(8)], s));
    }
    // torque of the rotate-back: - s <D g, J_k z>  (after the stores: they do not wait for the reductions)
    TorqueAcc<V> tq = torque_zero<V>();
    torque_acc(tq, t, zl);
    torque_commit(tq, cst<S>(0.f) - s, tau_old, g_wig, e, lane);
}

}  // namespace

template <class S>
void launch_gather_rotate_scale_t(GP<S> x, const int* src, const int* tgt, GP<S> wig, GP<S> rad, long long e0, int n_e,
                                  AP<S> A0, AP<S> A1, AP<S> A2, cudaStream_t st, ImgShare sh) {
    if (n_e <= 0) return;
    gather_rotate_scale_kernel<S><<<share_grid(sh, grs_warps<S>(), n_e), grs_warps<S>() * 32, 0, st>>>(x, src, tgt, wig, rad, e0, n_e, A0, A1, A2, sh);
    UMAB_LAUNCH_CHECK();
}
template <class S>
void launch_gather_rotate_bwd_t(GP<S> x, const int* row_ptr, const int* src, GP<S> wig, GP<S> rad, long long e0,
                                int node0, int n_nodes, GP<S> gA0, GP<S> gA1, GP<S> gA2, AP<S> g_rad, GP<S> G,
                                GP<S> g_x, GP<S> g_wig, cudaStream_t st) {
    if (n_nodes <= 0) return;
    if (g_rad.planes())
        gather_rotate_bwd_kernel<S, true><<<(n_nodes + 7) / 8, 256, 0, st>>>(x, row_ptr, src, wig, rad, e0, node0, n_nodes,
                                                                             gA0, gA1, gA2, g_rad, G, g_x, g_wig);
    else
        gather_rotate_bwd_kernel<S, false><<<(n_nodes + 7) / 8, 256, 0, st>>>(x, row_ptr, src, wig, rad, e0, node0, n_nodes,
                                                                              gA0, gA1, gA2, g_rad, G, g_x, g_wig);
    UMAB_LAUNCH_CHECK();
}
// closed chunks: target halves over the CSR rows, then source halves over the out-edge lists (no G, no source_reduce)
template <class S>
void launch_gather_rotate_bwd_closed_t(GP<S> x, const int* row_ptr, const int* sptr, const int* sedge, GP<S> wig, GP<S> rad,
                                       long long e0, int node0, int n_nodes, GP<S> gA0, GP<S> gA1, GP<S> gA2, AP<S> g_rad,
                                       GP<S> g_x, GP<S> g_wig, cudaStream_t st, ImgShare sh, int e_img) {
    if (n_nodes <= 0) return;
    // (a 128-register build of these kernels -- 2 CTAs per SM, ~500 B of spills -- was measured: 111 ms instead of 74)
    const dim3 grid(share_grid(sh, UMAB_HALF_NW, n_nodes));
    // dual numbers: the nine row accumulators of every warp live in shared memory (value + tangent: 9216 B per warp)
    const size_t smem = half_blocked<S>() ? (size_t)UMAB_HALF_NW * 9 * half_planes<S>() * 32 * sizeof(float4) : 0;
    if (smem > 48 * 1024) {
        UMAB_CUDA(cudaFuncSetAttribute(gather_rotate_bwd_half_kernel<1, S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        UMAB_CUDA(cudaFuncSetAttribute(gather_rotate_bwd_half_kernel<0, S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        UMAB_CUDA(cudaFuncSetAttribute(gather_rotate_bwd_half_kernel<1, S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        UMAB_CUDA(cudaFuncSetAttribute(gather_rotate_bwd_half_kernel<0, S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (g_rad.planes()) {
        gather_rotate_bwd_half_kernel<1, S, true><<<grid, UMAB_HALF_NW * 32, smem, st>>>(x, row_ptr, nullptr, wig, rad, e0, node0, n_nodes, gA0,
                                                                        gA1, gA2, g_rad, g_x, g_wig, sh, e_img);
        UMAB_LAUNCH_CHECK();
        gather_rotate_bwd_half_kernel<0, S, true><<<grid, UMAB_HALF_NW * 32, smem, st>>>(x, sptr, sedge, wig, rad, e0, node0, n_nodes, gA0, gA1,
                                                                        gA2, g_rad, g_x, g_wig, sh, e_img);
    } else {
        gather_rotate_bwd_half_kernel<1, S, false><<<grid, UMAB_HALF_NW * 32, smem, st>>>(x, row_ptr, nullptr, wig, rad, e0, node0, n_nodes, gA0,
                                                                         gA1, gA2, g_rad, g_x, g_wig, sh, e_img);
        UMAB_LAUNCH_CHECK();
        gather_rotate_bwd_half_kernel<0, S, false><<<grid, UMAB_HALF_NW * 32, smem, st>>>(x, sptr, sedge, wig, rad, e0, node0, n_nodes, gA0, gA1,
                                                                         gA2, g_rad, g_x, g_wig, sh, e_img);
    }
    UMAB_LAUNCH_CHECK();
}
void launch_source_reduce(const float* G, const int* sptr, const int* sedge, int n_nodes, float* g_x, cudaStream_t st) {
    if (n_nodes <= 0) return;
    source_reduce_kernel<<<(n_nodes + 7) / 8, 256, 0, st>>>(G, sptr, sedge, n_nodes, g_x);
    UMAB_LAUNCH_CHECK();
}
template <class S>
void launch_combine_gate_fwd_t(GP<S> Y0, GP<S> Y1, GP<S> Y2, int n_e, AP<S> B0, AP<S> B1, AP<S> B2, cudaStream_t st,
                               ImgShare sh) {
    if (n_e <= 0) return;
    combine_gate_fwd_kernel<S><<<share_grid(sh, 8, n_e), 256, 0, st>>>(Y0, Y1, Y2, n_e, B0, B1, B2, sh);
    UMAB_LAUNCH_CHECK();
}
void launch_gate_b0(GP<float> Y0, int n_e, AP<float> B0, float* sg, cudaStream_t st) {
    if (n_e <= 0) return;
    gate_b0_kernel<<<(n_e + 7) / 8, 256, 0, st>>>(Y0, n_e, B0, sg);
    UMAB_LAUNCH_CHECK();
}

template <class S>
void launch_combine_gate_bwd_t(GP<S> Y0, GP<S> Y1, GP<S> Y2, int n_e, GP<S> gB0, GP<S> gB1, GP<S> gB2, AP<S> gY0,
                               AP<S> gY1, AP<S> gY2, cudaStream_t st, ImgShare sh) {
    if (n_e <= 0) return;
    combine_gate_bwd_kernel<S><<<share_grid(sh, 8, n_e), 256, 0, st>>>(Y0, Y1, Y2, n_e, gB0, gB1, gB2, gY0, gY1, gY2, sh);
    UMAB_LAUNCH_CHECK();
}
template <class S>
void launch_rotate_back_reduce_t(int mode, GP<S> Z0, GP<S> Z1, GP<S> Z2, const int* row_ptr, GP<S> wig, GP<S> env,
                                 float scale, long long e0, int node0, int n_nodes, GP<S> base, GP<S> out,
                                 cudaStream_t st, ImgShare sh, int e_img) {
    if (n_nodes <= 0) return;
    dim3 grid(share_grid(sh, 8, n_nodes));
    if (mode == 0)
        rotate_back_reduce_kernel<0, S><<<grid, 256, 0, st>>>(Z0, Z1, Z2, row_ptr, wig, env, scale, e0, node0, n_nodes, base, out, sh, e_img);
    else
        rotate_back_reduce_kernel<1, S><<<grid, 256, 0, st>>>(Z0, Z1, Z2, row_ptr, wig, env, scale, e0, node0, n_nodes, base, out, sh, e_img);
    UMAB_LAUNCH_CHECK();
}
template <class S>
void launch_rotate_back_bwd_t(int mode, GP<S> Z0, GP<S> Z1, GP<S> Z2, const int* tgt, GP<S> wig, GP<S> env, float scale,
                              long long e0, int n_e, GP<S> g_out, AP<S> gZ0, AP<S> gZ1, AP<S> gZ2, GP<S> g_env,
                              GP<S> g_wig, cudaStream_t st, ImgShare sh) {
    if (n_e <= 0) return;
    dim3 grid(share_grid(sh, 8, n_e));
#define UMAB_RBB(MODE, PL)                                                                                           \
    rotate_back_bwd_kernel<MODE, S, PL><<<grid, 256, 0, st>>>(Z0, Z1, Z2, tgt, wig, env, scale, e0, n_e, g_out, gZ0, gZ1, \
                                                               gZ2, g_env, g_wig, sh)
    const bool pl = gZ0.planes();
    if (mode == 0) {
        if (pl) UMAB_RBB(0, true); else UMAB_RBB(0, false);
    } else {
        if (pl) UMAB_RBB(1, true); else UMAB_RBB(1, false);
    }
#undef UMAB_RBB
    UMAB_LAUNCH_CHECK();
}

#define UMAB_INST(S)                                                                                                  \
    template void launch_gather_rotate_scale_t<S>(GP<S>, const int*, const int*, GP<S>, GP<S>, long long, int, AP<S>,  \
                                                  AP<S>, AP<S>, cudaStream_t, ImgShare);                              \
    template void launch_gather_rotate_bwd_t<S>(GP<S>, const int*, const int*, GP<S>, GP<S>, long long, int, int,      \
                                                GP<S>, GP<S>, GP<S>, AP<S>, GP<S>, GP<S>, GP<S>, cudaStream_t);       \
    template void launch_gather_rotate_bwd_closed_t<S>(GP<S>, const int*, const int*, const int*, GP<S>, GP<S>,       \
                                                       long long, int, int, GP<S>, GP<S>, GP<S>, AP<S>, GP<S>, GP<S>,  \
                                                       cudaStream_t, ImgShare, int);                                   \
    template void launch_combine_gate_fwd_t<S>(GP<S>, GP<S>, GP<S>, int, AP<S>, AP<S>, AP<S>, cudaStream_t, ImgShare); \
    template void launch_combine_gate_bwd_t<S>(GP<S>, GP<S>, GP<S>, int, GP<S>, GP<S>, GP<S>, AP<S>, AP<S>, AP<S>,     \
                                               cudaStream_t, ImgShare);                                               \
    template void launch_rotate_back_reduce_t<S>(int, GP<S>, GP<S>, GP<S>, const int*, GP<S>, GP<S>, float, long long, \
                                                 int, int, GP<S>, GP<S>, cudaStream_t, ImgShare, int);                \
    template void launch_rotate_back_bwd_t<S>(int, GP<S>, GP<S>, GP<S>, const int*, GP<S>, GP<S>, float, long long,    \
                                              int, GP<S>, AP<S>, AP<S>, AP<S>, GP<S>, GP<S>, cudaStream_t, ImgShare);
UMAB_INST(float)
UMAB_INST(D1)
#undef UMAB_INST

}  // namespace umab
