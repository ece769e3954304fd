// K2/K3: per-edge geometry -- edge vector, distance, polynomial envelope, Gaussian distance
// basis and the closed-form l<=2 Wigner blocks -- and the adjoint that turns the accumulated
// per-edge gradients (d/dgauss, d/denv, and the torque of the edge frame) into dE/d(edge vector).
//
// Replaces fairchem's GaussianSmearing / PolynomialEnvelope / init_edge_rot_euler_angles +
// eulers_to_wigner, reached from the reference through predict_unit.predict
// (pdb2reaction/uma_pysis.py:385).  Instead of Euler angles (acos/atan2 + Z J Z J Z products) the
// blocks are polynomials of the rotation R_e = Rx(-beta) Ry(-alpha) written directly in the unit
// vector n = (x,y,z):  rows (z/s, 0, -x/s), (x, y, z), (xy/s, -s, zy/s),  s = sqrt(x^2+z^2);
// D1 = R_e, D2[m][n] = (2/3) <R_e^T Q_m R_e, Q_n> with Q_m the quadratic forms of the l=2 real
// harmonics.  The energy is invariant to the roll angle gamma (mmax = lmax), so gamma = 0.
// Twin: oracle/staged.py geometry_fwd / geometry_bwd.
#include "dual.cuh"

namespace umab {

namespace {

constexpr float SQ3H = 0.86602540378443864676f;   // sqrt(3)/2

template <class S> struct M3 { S v[3][3]; };

// Q_m X
template <class S>
__device__ __forceinline__ M3<S> qleft(int m, const M3<S>& x) {
    M3<S> o;
    const S zero = cst<S>(0.f);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        S r0 = x.v[0][j], r1 = x.v[1][j], r2 = x.v[2][j];
        S o0, o1, o2;
        switch (m) {
            case 0: o0 = SQ3H * r2; o1 = zero; o2 = SQ3H * r0; break;
            case 1: o0 = SQ3H * r1; o1 = SQ3H * r0; o2 = zero; break;
            case 2: o0 = -0.5f * r0; o1 = r1; o2 = -0.5f * r2; break;
            case 3: o0 = zero; o1 = SQ3H * r2; o2 = SQ3H * r1; break;
            default: o0 = -SQ3H * r0; o1 = zero; o2 = SQ3H * r2; break;
        }
        o.v[0][j] = o0; o.v[1][j] = o1; o.v[2][j] = o2;
    }
    return o;
}
// <M, Q_n>, n = 0..4
template <class S>
__device__ __forceinline__ void qdot(const M3<S>& m, S* out) {
    out[0] = SQ3H * (m.v[0][2] + m.v[2][0]);
    out[1] = SQ3H * (m.v[0][1] + m.v[1][0]);
    out[2] = -0.5f * m.v[0][0] + m.v[1][1] - 0.5f * m.v[2][2];
    out[3] = SQ3H * (m.v[1][2] + m.v[2][1]);
    out[4] = SQ3H * (m.v[2][2] - m.v[0][0]);
}
// sum_n g[n] Q_n
template <class S>
__device__ __forceinline__ M3<S> qcomb(const S* g) {
    M3<S> t;
    t.v[0][0] = -0.5f * g[2] - SQ3H * g[4];
    t.v[1][1] = g[2];
    t.v[2][2] = -0.5f * g[2] + SQ3H * g[4];
    t.v[0][1] = t.v[1][0] = SQ3H * g[1];
    t.v[0][2] = t.v[2][0] = SQ3H * g[0];
    t.v[1][2] = t.v[2][1] = SQ3H * g[3];
    return t;
}
template <class S>
__device__ __forceinline__ M3<S> matmul(const M3<S>& a, const M3<S>& b) {
    M3<S> o;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            o.v[i][j] = a.v[i][0] * b.v[0][j] + a.v[i][1] * b.v[1][j] + a.v[i][2] * b.v[2][j];
    return o;
}
template <class S>
__device__ __forceinline__ M3<S> matmul_tn(const M3<S>& a, const M3<S>& b) {   // a^T b
    M3<S> o;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            o.v[i][j] = a.v[0][i] * b.v[0][j] + a.v[1][i] * b.v[1][j] + a.v[2][i] * b.v[2][j];
    return o;
}

template <class S> struct EdgeFrame { S x, y, z, s, ca, sa, inv_s; bool pole; };

template <class S>
__device__ __forceinline__ EdgeFrame<S> make_frame(S nx, S ny, S nz) {
    EdgeFrame<S> f;
    f.x = nx; f.y = ny; f.z = nz;
    const S s2 = nx * nx + nz * nz;
    f.pole = val(s2) < 1e-24f;
    f.s = f.pole ? cst<S>(0.f) : s_sqrt(s2);
    f.inv_s = f.pole ? cst<S>(0.f) : 1.0f / f.s;
    f.ca = f.pole ? cst<S>(1.f) : nz * f.inv_s;
    f.sa = f.pole ? cst<S>(0.f) : nx * f.inv_s;
    return f;
}
template <class S>
__device__ __forceinline__ M3<S> frame_rot(const EdgeFrame<S>& f) {
    M3<S> r;
    const S zero = cst<S>(0.f);
    r.v[0][0] = f.ca;        r.v[0][1] = zero;  r.v[0][2] = -f.sa;
    r.v[1][0] = f.x;         r.v[1][1] = f.y;   r.v[1][2] = f.z;
    r.v[2][0] = f.y * f.sa;  r.v[2][1] = -f.s;  r.v[2][2] = f.y * f.ca;
    return r;
}

// pos is dual in the Hessian path (tangent = displacement direction); everything derived is dual
template <class S>
__global__ void __launch_bounds__(256)
geometry_fwd_kernel(GP<S> pos, const int* __restrict__ src, const int* __restrict__ tgt, int n_edges, float cutoff,
                    GP<S> vec, GP<S> dist, GP<S> env, GP<S> wig, GP<S> gauss) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x % 32;
    S d = cst<S>(0.f);
    if (e < n_edges) {
        int j = src[e], i = tgt[e];
        S vx = pos.ld(j * 3 + 0) - pos.ld(i * 3 + 0);
        S vy = pos.ld(j * 3 + 1) - pos.ld(i * 3 + 1);
        S vz = pos.ld(j * 3 + 2) - pos.ld(i * 3 + 2);
        d = s_sqrt(vx * vx + vy * vy + vz * vz);
        S inv = 1.0f / d;
        EdgeFrame<S> f = make_frame<S>(vx * inv, vy * inv, vz * inv);
        M3<S> r = frame_rot(f);
        vec.st(e * 3 + 0, vx); vec.st(e * 3 + 1, vy); vec.st(e * 3 + 2, vz);
        dist.st(e, d);
        S u = d / cutoff;
        S u2 = u * u, u4 = u2 * u2, u5 = u4 * u;
        S ev = 1.0f + u5 * (-21.0f + u * (35.0f - 15.0f * u));
        env.st(e, (val(u) < 1.0f) ? ev : cst<S>(0.f));
        const long long wo = (long long)e * WIG;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) wig.st(wo + a * 3 + b, r.v[a][b]);
#pragma unroll
        for (int m = 0; m < 5; ++m) {
            M3<S> mm = matmul_tn(r, qleft(m, r));
            S o[5];
            qdot(mm, o);
#pragma unroll
            for (int n = 0; n < 5; ++n) wig.st(wo + 9 + m * 5 + n, (2.0f / 3.0f) * o[n]);
        }
        wig.st(wo + 34, cst<S>(0.f)); wig.st(wo + 35, cst<S>(0.f));
    }
    // Gaussian basis, written warp-cooperatively so that each row is one coalesced 256 B store
    const float delta = cutoff / (NB - 1);
    const float coeff = -0.5f / ((2.0f * delta) * (2.0f * delta));
    const int e0 = e - lane;
#pragma unroll 4
    for (int q = 0; q < 32; ++q) {
        S dq = s_shfl(d, q);
        int eq = e0 + q;
        if (eq < n_edges) {
            S t0 = dq - (float)(2 * lane) * delta, t1 = dq - (float)(2 * lane + 1) * delta;
            gauss.st((long long)eq * NB + 2 * lane, s_exp(coeff * (t0 * t0)));
            gauss.st((long long)eq * NB + 2 * lane + 1, s_exp(coeff * (t1 * t1)));
        }
    }
}

template <class S>
__global__ void __launch_bounds__(256)
geometry_bwd_kernel(GP<S> vec, GP<S> dist, GP<S> wig, GP<S> gauss, GP<S> g_gauss, GP<S> g_env, GP<S> g_wig,
                    int n_edges, float cutoff, GP<S> g_vec) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x % 32;
    const float delta = cutoff / (NB - 1);
    const float coeff = -0.5f / ((2.0f * delta) * (2.0f * delta));
    S d = (e < n_edges) ? dist.ld(e) : cst<S>(1.f);
    // d/dd of the Gaussian basis, reduced warp-cooperatively (coalesced row reads)
    S g_d = cst<S>(0.f);
    const int e0 = e - lane;
#pragma unroll 4
    for (int q = 0; q < 32; ++q) {
        S dq = s_shfl(d, q);
        int eq = e0 + q;
        S part = cst<S>(0.f);
        if (eq < n_edges) {
            const long long o = (long long)eq * NB + 2 * lane;
            S t0 = dq - (float)(2 * lane) * delta, t1 = dq - (float)(2 * lane + 1) * delta;
            part = g_gauss.ld(o) * gauss.ld(o) * (2.0f * coeff) * t0 + g_gauss.ld(o + 1) * gauss.ld(o + 1) * (2.0f * coeff) * t1;
        }
        part = warp_sum(part);
        if (lane == q) g_d = part;
    }
    if (e >= n_edges) return;
    S u = d / cutoff;
    S u2 = u * u, u4 = u2 * u2;
    S denv = (val(u) < 1.0f) ? (u4 * (-105.0f + u * (210.0f - 105.0f * u))) / cutoff : cst<S>(0.f);
    g_d = g_d + g_env.ld(e) * denv;

    // rotation part: the edge kernels accumulated the torque (t_x, t_z) of the edge frame (edge_ops.cu: D -> (1 + dw.J) D);
    // keeping R_e n = y_hat under a change dn of the direction needs  R_e dn = y_hat x dw = (dw_z, 0, -dw_x), and the
    // energy does not depend on dw_y (roll angle), so  dL = t_z (R_e dn)_x - t_x (R_e dn)_z  and
    // dL/dn = t_z R_e[0] - t_x R_e[2]  -- already perpendicular to n = R_e[1]; no pole: any orthonormal frame works
    const S inv = 1.0f / d;
    const long long wo = (long long)e * WIG;
    const S tx = g_wig.ld((long long)e * 4 + 0), tz = g_wig.ld((long long)e * 4 + 1);
    S gn[3], n[3];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        gn[b] = tz * wig.ld(wo + 0 * 3 + b) - tx * wig.ld(wo + 2 * 3 + b);
        n[b] = vec.ld(e * 3 + b) * inv;
    }
#pragma unroll
    for (int b = 0; b < 3; ++b) g_vec.st(e * 3 + b, gn[b] * inv + g_d * n[b]);
}

// F[i] = sum_{e into i} g_vec[e] - sum_{e out of i} g_vec[e]   (vec = pos[src] - pos[tgt]); linear:
// the Hessian path runs it once per plane
__global__ void force_reduce_kernel(const float* __restrict__ g_vec, const int* __restrict__ row_ptr,
                                    const int* __restrict__ sptr, const int* __restrict__ sedge,
                                    int n_nodes, float* __restrict__ forces) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
        fx += g_vec[e * 3 + 0]; fy += g_vec[e * 3 + 1]; fz += g_vec[e * 3 + 2];
    }
    float ox = 0.f, oy = 0.f, oz = 0.f;
    for (int k = sptr[i]; k < sptr[i + 1]; ++k) {
        int e = sedge[k];
        ox += g_vec[e * 3 + 0]; oy += g_vec[e * 3 + 1]; oz += g_vec[e * 3 + 2];
    }
    forces[i * 3 + 0] = fx - ox; forces[i * 3 + 1] = fy - oy; forces[i * 3 + 2] = fz - oz;
}

}  // namespace

template <class S>
void launch_geometry_fwd_t(GP<S> pos, const int* src, const int* tgt, int n_edges, float cutoff, GP<S> vec,
                           GP<S> dist, GP<S> env, GP<S> wig, GP<S> gauss, cudaStream_t st) {
    if (n_edges <= 0) return;
    geometry_fwd_kernel<S><<<(n_edges + 255) / 256, 256, 0, st>>>(pos, src, tgt, n_edges, cutoff, vec, dist, env, wig, gauss);
    UMAB_LAUNCH_CHECK();
}
template <class S>
void launch_geometry_bwd_t(GP<S> vec, GP<S> dist, GP<S> wig, GP<S> gauss, GP<S> g_gauss, GP<S> g_env, GP<S> g_wig,
                           int n_edges, float cutoff, GP<S> g_vec, cudaStream_t st) {
    if (n_edges <= 0) return;
    geometry_bwd_kernel<S><<<(n_edges + 255) / 256, 256, 0, st>>>(vec, dist, wig, gauss, g_gauss, g_env, g_wig, n_edges,
                                                                  cutoff, g_vec);
    UMAB_LAUNCH_CHECK();
}
#define UMAB_INST(S)                                                                                               \
    template void launch_geometry_fwd_t<S>(GP<S>, const int*, const int*, int, float, GP<S>, GP<S>, GP<S>, GP<S>,  \
                                           GP<S>, cudaStream_t);                                                  \
    template void launch_geometry_bwd_t<S>(GP<S>, GP<S>, GP<S>, GP<S>, GP<S>, GP<S>, GP<S>, int, float, GP<S>, cudaStream_t);
UMAB_INST(float)
UMAB_INST(D1)
#undef UMAB_INST

void launch_force_reduce(const float* g_vec, const int* row_ptr, const int* sptr, const int* sedge,
                         int n_nodes, float* forces, cudaStream_t st) {
    if (n_nodes <= 0) return;
    force_reduce_kernel<<<(n_nodes + 255) / 256, 256, 0, st>>>(g_vec, row_ptr, sptr, sedge, n_nodes, forces);
    UMAB_LAUNCH_CHECK();
}

}  // namespace umab
