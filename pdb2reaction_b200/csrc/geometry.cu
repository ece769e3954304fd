// K2/K3: per-edge geometry -- edge vector, distance, polynomial envelope, Gaussian distance
// basis and the closed-form l<=2 Wigner blocks -- and the adjoint that turns the accumulated
// per-edge gradients (d/dgauss, d/denv, d/dwigner) into dE/d(edge vector).
//
// Replaces fairchem's GaussianSmearing / PolynomialEnvelope / init_edge_rot_euler_angles +
// eulers_to_wigner, reached from the reference through predict_unit.predict
// (pdb2reaction/uma_pysis.py:385).  Instead of Euler angles (acos/atan2 + Z J Z J Z products) the
// blocks are polynomials of the rotation R_e = Rx(-beta) Ry(-alpha) written directly in the unit
// vector n = (x,y,z):  rows (z/s, 0, -x/s), (x, y, z), (xy/s, -s, zy/s),  s = sqrt(x^2+z^2);
// D1 = R_e, D2[m][n] = (2/3) <R_e^T Q_m R_e, Q_n> with Q_m the quadratic forms of the l=2 real
// harmonics.  The energy is invariant to the roll angle gamma (mmax = lmax), so gamma = 0.
// Twin: oracle/staged.py geometry_fwd / geometry_bwd.
#include "common.cuh"

namespace umab {

namespace {

constexpr float SQ3H = 0.86602540378443864676f;   // sqrt(3)/2

struct M3 { float v[3][3]; };

// Q_m X
__device__ __forceinline__ M3 qleft(int m, const M3& x) {
    M3 o;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float r0 = x.v[0][j], r1 = x.v[1][j], r2 = x.v[2][j];
        float o0, o1, o2;
        switch (m) {
            case 0: o0 = SQ3H * r2; o1 = 0.f; o2 = SQ3H * r0; break;
            case 1: o0 = SQ3H * r1; o1 = SQ3H * r0; o2 = 0.f; break;
            case 2: o0 = -0.5f * r0; o1 = r1; o2 = -0.5f * r2; break;
            case 3: o0 = 0.f; o1 = SQ3H * r2; o2 = SQ3H * r1; break;
            default: o0 = -SQ3H * r0; o1 = 0.f; o2 = SQ3H * r2; break;
        }
        o.v[0][j] = o0; o.v[1][j] = o1; o.v[2][j] = o2;
    }
    return o;
}
// <M, Q_n>, n = 0..4
__device__ __forceinline__ void qdot(const M3& m, float* out) {
    out[0] = SQ3H * (m.v[0][2] + m.v[2][0]);
    out[1] = SQ3H * (m.v[0][1] + m.v[1][0]);
    out[2] = -0.5f * m.v[0][0] + m.v[1][1] - 0.5f * m.v[2][2];
    out[3] = SQ3H * (m.v[1][2] + m.v[2][1]);
    out[4] = SQ3H * (m.v[2][2] - m.v[0][0]);
}
// sum_n g[n] Q_n
__device__ __forceinline__ M3 qcomb(const float* g) {
    M3 t;
    t.v[0][0] = -0.5f * g[2] - SQ3H * g[4];
    t.v[1][1] = g[2];
    t.v[2][2] = -0.5f * g[2] + SQ3H * g[4];
    t.v[0][1] = t.v[1][0] = SQ3H * g[1];
    t.v[0][2] = t.v[2][0] = SQ3H * g[0];
    t.v[1][2] = t.v[2][1] = SQ3H * g[3];
    return t;
}
__device__ __forceinline__ M3 matmul(const M3& a, const M3& b) {
    M3 o;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            o.v[i][j] = a.v[i][0] * b.v[0][j] + a.v[i][1] * b.v[1][j] + a.v[i][2] * b.v[2][j];
    return o;
}
__device__ __forceinline__ M3 matmul_tn(const M3& a, const M3& b) {   // a^T b
    M3 o;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            o.v[i][j] = a.v[0][i] * b.v[0][j] + a.v[1][i] * b.v[1][j] + a.v[2][i] * b.v[2][j];
    return o;
}

struct EdgeFrame { float x, y, z, s, ca, sa, inv_s; bool pole; };

__device__ __forceinline__ EdgeFrame make_frame(float nx, float ny, float nz) {
    EdgeFrame f;
    f.x = nx; f.y = ny; f.z = nz;
    f.s = sqrtf(nx * nx + nz * nz);
    f.pole = f.s < 1e-12f;
    f.inv_s = f.pole ? 0.f : 1.0f / f.s;
    f.ca = f.pole ? 1.f : nz * f.inv_s;
    f.sa = f.pole ? 0.f : nx * f.inv_s;
    return f;
}
__device__ __forceinline__ M3 frame_rot(const EdgeFrame& f) {
    M3 r;
    r.v[0][0] = f.ca;        r.v[0][1] = 0.f;   r.v[0][2] = -f.sa;
    r.v[1][0] = f.x;         r.v[1][1] = f.y;   r.v[1][2] = f.z;
    r.v[2][0] = f.y * f.sa;  r.v[2][1] = -f.s;  r.v[2][2] = f.y * f.ca;
    return r;
}

__global__ void __launch_bounds__(256)
geometry_fwd_kernel(const float* __restrict__ pos, const int* __restrict__ src, const int* __restrict__ tgt,
                    int n_edges, float cutoff, float* __restrict__ vec, float* __restrict__ dist,
                    float* __restrict__ env, float* __restrict__ wig, float* __restrict__ gauss) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x % 32;
    float d = 0.f;
    if (e < n_edges) {
        int j = src[e], i = tgt[e];
        float vx = pos[j * 3 + 0] - pos[i * 3 + 0];
        float vy = pos[j * 3 + 1] - pos[i * 3 + 1];
        float vz = pos[j * 3 + 2] - pos[i * 3 + 2];
        d = sqrtf(vx * vx + vy * vy + vz * vz);
        float inv = 1.0f / d;
        EdgeFrame f = make_frame(vx * inv, vy * inv, vz * inv);
        M3 r = frame_rot(f);
        vec[e * 3 + 0] = vx; vec[e * 3 + 1] = vy; vec[e * 3 + 2] = vz;
        dist[e] = d;
        float u = d / cutoff;
        float u2 = u * u, u4 = u2 * u2, u5 = u4 * u;
        float ev = 1.0f + u5 * (-21.0f + u * (35.0f - 15.0f * u));
        env[e] = (u < 1.0f) ? ev : 0.f;
        float* w = wig + (long long)e * WIG;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) w[a * 3 + b] = r.v[a][b];
#pragma unroll
        for (int m = 0; m < 5; ++m) {
            M3 mm = matmul_tn(r, qleft(m, r));
            float o[5];
            qdot(mm, o);
#pragma unroll
            for (int n = 0; n < 5; ++n) w[9 + m * 5 + n] = (2.0f / 3.0f) * o[n];
        }
        w[34] = 0.f; w[35] = 0.f;
    }
    // Gaussian basis, written warp-cooperatively so that each row is one coalesced 256 B store
    const float delta = cutoff / (NB - 1);
    const float coeff = -0.5f / ((2.0f * delta) * (2.0f * delta));
    const int e0 = e - lane;
#pragma unroll 4
    for (int q = 0; q < 32; ++q) {
        float dq = __shfl_sync(0xffffffffu, d, q);
        int eq = e0 + q;
        if (eq < n_edges) {
            float t0 = dq - (float)(2 * lane) * delta, t1 = dq - (float)(2 * lane + 1) * delta;
            float2 g = make_float2(expf(coeff * t0 * t0), expf(coeff * t1 * t1));
            *reinterpret_cast<float2*>(gauss + (long long)eq * NB + 2 * lane) = g;
        }
    }
}

__global__ void __launch_bounds__(256)
geometry_bwd_kernel(const float* __restrict__ vec, const float* __restrict__ dist, const float* __restrict__ wig,
                    const float* __restrict__ gauss, const float* __restrict__ g_gauss,
                    const float* __restrict__ g_env, const float* __restrict__ g_wig, int n_edges,
                    float cutoff, float* __restrict__ g_vec) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x % 32;
    const float delta = cutoff / (NB - 1);
    const float coeff = -0.5f / ((2.0f * delta) * (2.0f * delta));
    float d = (e < n_edges) ? dist[e] : 1.f;
    // d/dd of the Gaussian basis, reduced warp-cooperatively (coalesced row reads)
    float g_d = 0.f;
    const int e0 = e - lane;
#pragma unroll 4
    for (int q = 0; q < 32; ++q) {
        float dq = __shfl_sync(0xffffffffu, d, q);
        int eq = e0 + q;
        float part = 0.f;
        if (eq < n_edges) {
            float2 gg = *reinterpret_cast<const float2*>(g_gauss + (long long)eq * NB + 2 * lane);
            float2 gv = *reinterpret_cast<const float2*>(gauss + (long long)eq * NB + 2 * lane);
            float t0 = dq - (float)(2 * lane) * delta, t1 = dq - (float)(2 * lane + 1) * delta;
            part = gg.x * gv.x * (2.0f * coeff) * t0 + gg.y * gv.y * (2.0f * coeff) * t1;
        }
        part = warp_sum(part);
        if (lane == q) g_d = part;
    }
    if (e >= n_edges) return;
    float u = d / cutoff;
    float u2 = u * u, u4 = u2 * u2;
    float denv = (u < 1.0f) ? u4 * (-105.0f + u * (210.0f - 105.0f * u)) / cutoff : 0.f;
    g_d += g_env[e] * denv;

    float vx = vec[e * 3 + 0], vy = vec[e * 3 + 1], vz = vec[e * 3 + 2];
    float inv = 1.0f / d;
    EdgeFrame f = make_frame(vx * inv, vy * inv, vz * inv);
    M3 r;
    const float* w = wig + (long long)e * WIG;
    const float* gw = g_wig + (long long)e * WIG;
    M3 g_rot;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) { r.v[a][b] = w[a * 3 + b]; g_rot.v[a][b] = gw[a * 3 + b]; }
    // dL/dR += (4/3) sum_m Q_m R (sum_n G2[m][n] Q_n)
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        float g5[5];
#pragma unroll
        for (int n = 0; n < 5; ++n) g5[n] = gw[9 + m * 5 + n];
        M3 t = qleft(m, matmul(r, qcomb(g5)));
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) g_rot.v[a][b] += (4.0f / 3.0f) * t.v[a][b];
    }
    float g_ca = g_rot.v[0][0] + f.y * g_rot.v[2][2];
    float g_sa = -g_rot.v[0][2] + f.y * g_rot.v[2][0];
    float g_s = -g_rot.v[2][1];
    float gx = g_rot.v[1][0];
    float gy = g_rot.v[1][1] + f.sa * g_rot.v[2][0] + f.ca * g_rot.v[2][2];
    float gz = g_rot.v[1][2];
    float g_s_tot = g_s - (g_ca * f.ca + g_sa * f.sa) * f.inv_s;
    gx += g_sa * f.inv_s + g_s_tot * f.sa;
    gz += g_ca * f.inv_s + g_s_tot * f.ca;
    float dotn = gx * f.x + gy * f.y + gz * f.z;
    g_vec[e * 3 + 0] = (gx - f.x * dotn) * inv + g_d * f.x;
    g_vec[e * 3 + 1] = (gy - f.y * dotn) * inv + g_d * f.y;
    g_vec[e * 3 + 2] = (gz - f.z * dotn) * inv + g_d * f.z;
}

// F[i] = sum_{e into i} g_vec[e] - sum_{e out of i} g_vec[e]   (vec = pos[src] - pos[tgt])
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

void launch_geometry_fwd(const float* pos, const int* src, const int* tgt, int n_edges, float cutoff,
                         float* vec, float* dist, float* env, float* wig, float* gauss, cudaStream_t st) {
    if (n_edges <= 0) return;
    geometry_fwd_kernel<<<(n_edges + 255) / 256, 256, 0, st>>>(pos, src, tgt, n_edges, cutoff, vec, dist,
                                                               env, wig, gauss);
    UMAB_LAUNCH_CHECK();
}

void launch_geometry_bwd(const float* vec, const float* dist, const float* wig, const float* gauss,
                         const float* g_gauss, const float* g_env, const float* g_wig, int n_edges,
                         float cutoff, float* g_vec, cudaStream_t st) {
    if (n_edges <= 0) return;
    geometry_bwd_kernel<<<(n_edges + 255) / 256, 256, 0, st>>>(vec, dist, wig, gauss, g_gauss, g_env, g_wig,
                                                               n_edges, cutoff, g_vec);
    UMAB_LAUNCH_CHECK();
}

void launch_force_reduce(const float* g_vec, const int* row_ptr, const int* sptr, const int* sedge,
                         int n_nodes, float* forces, cudaStream_t st) {
    if (n_nodes <= 0) return;
    force_reduce_kernel<<<(n_nodes + 255) / 256, 256, 0, st>>>(g_vec, row_ptr, sptr, sedge, n_nodes, forces);
    UMAB_LAUNCH_CHECK();
}

}  // namespace umab
