// Hessian assembly and vibrational pre-processing on the device.
//
//   fd_columns        H[:, k_q] = -(F(x + h e_k) - F(x - h e_k)) / 2h for a batch of displaced force evaluations,
//                     written straight into the [dof, dof] Hessian (reference: the column loop of
//                     uma_pysis._build_fd_hessian_gpu, pdb2reaction/uma_pysis.py:652-675)
//   mw_project        H <- sym( P (M^-1/2 H M^-1/2) P ),  P = I - Q Q^T  (Q = orthonormal translation / rotation
//                     basis, rank r <= 6), in place, fp64 -- the chain of ~10 dense passes of
//                     freq._mw_projected_hessian (pdb2reaction/freq.py:159-205: two mul_, three addmm_, the
//                     symmetrisation) in two passes over H:
//                       pass 1   A = Q^T (S H S)            [r, n]   (column-parallel, deterministic row-slab partials)
//                       pass 1b  B = A Q [r, r],  QB = Q B  [n, r]
//                       pass 2   per pair of 32 x 32 tiles (I <= J): H'[i,j] = s_i s_j H[i,j] - sum_k Q[i,k] A[k,j]
//                                - sum_k A[k,i] Q[j,k] + sum_k QB[i,k] Q[j,k];  out = (H' + H'^T) / 2 for both tiles
//                     (the reference forms H Q as (Q^T H)^T, i.e. with the same A; reproduced exactly).
// HBM-bound, tiny next to the eigendecomposition that follows; kept on the device so a sharded Hessian never
// takes a host round trip.
#include "common.cuh"

namespace umab {

namespace {

constexpr int RMAX = 6;
constexpr int SLAB = 256;          // rows per partial of pass 1

template <class T>
__global__ void __launch_bounds__(256)
fd_columns_kernel(const float* __restrict__ f, const int* __restrict__ ks, int n_cols, int dof, double two_h,
                  T* __restrict__ h, long long ldh) {
    // one thread per (row i, column q): coalesced reads of the two force vectors along i
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_cols * dof) return;
    const int q = (int)(t / dof), i = (int)(t % dof);
    // the reference's arithmetic, in the Hessian dtype T: col = -(Fp - Fm) / (2.0 * eps)   (uma_pysis.py:668-670);
    // torch's CUDA division of a tensor by a host scalar multiplies by the scalar's reciprocal, formed in double
    // and rounded to T (measured: x / 0.002 == x * 500.0f bit for bit) -- so do we
    const T fp = (T)f[(long long)(2 * q) * dof + i], fm = (T)f[(long long)(2 * q + 1) * dof + i];
    const T inv = (T)(1.0 / two_h);
    h[(long long)i * ldh + ks[q]] = -(fp - fm) * inv;
}

// partial[slab][k][j] = s_j * sum_{i in slab} Q[i,k] s_i H[i,j]
__global__ void __launch_bounds__(128)
mw_pass1_kernel(const double* __restrict__ h, int n, const double* __restrict__ s, const double* __restrict__ q, int r,
                double* __restrict__ partial) {
    const int j = blockIdx.x * 128 + threadIdx.x;
    const int slab = blockIdx.y;
    const int i0 = slab * SLAB, i1 = min(n, i0 + SLAB);
    __shared__ double qs[SLAB][RMAX];          // Q[i,k] * s_i of this slab
    for (int t = threadIdx.x; t < (i1 - i0) * RMAX; t += 128) {
        const int ii = t / RMAX, k = t % RMAX;
        qs[ii][k] = k < r ? q[(long long)(i0 + ii) * r + k] * s[i0 + ii] : 0.0;
    }
    __syncthreads();
    if (j >= n) return;
    double acc[RMAX] = {0, 0, 0, 0, 0, 0};
    for (int i = i0; i < i1; ++i) {
        const double v = h[(long long)i * n + j];
#pragma unroll
        for (int k = 0; k < RMAX; ++k) acc[k] += qs[i - i0][k] * v;
    }
    const double sj = s[j];
    for (int k = 0; k < r; ++k) partial[((long long)slab * r + k) * n + j] = acc[k] * sj;
}

// A[k,j] = sum over slabs (fixed order)
__global__ void mw_reduce_kernel(const double* __restrict__ partial, int n, int r, int n_slabs, double* __restrict__ a) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)r * n) return;
    double acc = 0.0;
    for (int sl = 0; sl < n_slabs; ++sl) acc += partial[(long long)sl * r * n + t];
    a[t] = acc;
}

// B = A Q [r,r] (one block, fixed-order tree), then QB = Q B [n,r]
__global__ void __launch_bounds__(256)
mw_small_kernel(const double* __restrict__ a, const double* __restrict__ q, int n, int r, double* __restrict__ b_out,
                double* __restrict__ qb) {
    __shared__ double red[256];
    __shared__ double b[RMAX * RMAX];
    for (int kl = 0; kl < r * r; ++kl) {
        const int k = kl / r, l = kl % r;
        double acc = 0.0;
        for (int j = threadIdx.x; j < n; j += 256) acc += a[(long long)k * n + j] * q[(long long)j * r + l];
        red[threadIdx.x] = acc;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) {
            if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) { b[kl] = red[0]; b_out[kl] = red[0]; }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += 256)
        for (int l = 0; l < r; ++l) {
            double acc = 0.0;
            for (int k = 0; k < r; ++k) acc += q[(long long)i * r + k] * b[k * r + l];
            qb[(long long)i * r + l] = acc;
        }
}

__global__ void __launch_bounds__(256)
mw_pass2_kernel(double* __restrict__ h, int n, const double* __restrict__ s, const double* __restrict__ q,
                const double* __restrict__ a, const double* __restrict__ qb, int r, int n_tiles) {
    // blockIdx.x enumerates tile pairs (I <= J)
    int I = 0, rem = blockIdx.x;
    while (rem >= n_tiles - I) { rem -= n_tiles - I; ++I; }
    const int J = I + rem;
    __shared__ double t1[32][33], t2[32][33];          // H'(I,J) and H'(J,I)
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
    auto proj = [&](int gi, int gj) -> double {
        double v = s[gi] * s[gj] * h[(long long)gi * n + gj];
        for (int k = 0; k < r; ++k)
            v += -q[(long long)gi * r + k] * a[(long long)k * n + gj] - a[(long long)k * n + gi] * q[(long long)gj * r + k] +
                 qb[(long long)gi * r + k] * q[(long long)gj * r + k];
        return v;
    };
    for (int yy = ty; yy < 32; yy += 8) {
        const int gi = I * 32 + yy, gj = J * 32 + tx;
        t1[yy][tx] = (gi < n && gj < n) ? proj(gi, gj) : 0.0;
        const int gi2 = J * 32 + yy, gj2 = I * 32 + tx;
        t2[yy][tx] = (gi2 < n && gj2 < n) ? proj(gi2, gj2) : 0.0;
    }
    __syncthreads();
    for (int yy = ty; yy < 32; yy += 8) {
        const int gi = I * 32 + yy, gj = J * 32 + tx;
        if (gi < n && gj < n) h[(long long)gi * n + gj] = 0.5 * (t1[yy][tx] + t2[tx][yy]);
        const int gi2 = J * 32 + yy, gj2 = I * 32 + tx;
        if (I != J && gi2 < n && gj2 < n) h[(long long)gi2 * n + gj2] = 0.5 * (t2[yy][tx] + t1[tx][yy]);
    }
}

}  // namespace

void launch_fd_columns(const float* f, const int* ks, int n_cols, int dof, double h_step, void* hmat, long long ldh,
                       bool f64, cudaStream_t st) {
    if (n_cols <= 0 || dof <= 0) return;
    const long long total = (long long)n_cols * dof;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (f64) fd_columns_kernel<double><<<blocks, 256, 0, st>>>(f, ks, n_cols, dof, 2.0 * h_step, (double*)hmat, ldh);
    else fd_columns_kernel<float><<<blocks, 256, 0, st>>>(f, ks, n_cols, dof, 2.0 * h_step, (float*)hmat, ldh);
    UMAB_LAUNCH_CHECK();
}

size_t mw_project_workspace_doubles(int n, int r) {
    const int n_slabs = (n + SLAB - 1) / SLAB;
    return (size_t)n_slabs * r * n + (size_t)r * n + (size_t)r * r + (size_t)n * r;
}

void launch_mw_project(double* h, int n, const double* inv_sqrt_m, const double* q, int r, double* ws, cudaStream_t st) {
    if (n <= 0) return;
    if (r < 0 || r > RMAX) throw CudaError("mw_project: the projector rank must be 0..6");
    const int n_slabs = (n + SLAB - 1) / SLAB;
    double* partial = ws;
    double* a = partial + (size_t)n_slabs * r * n;
    double* b = a + (size_t)r * n;
    double* qb = b + (size_t)r * r;
    if (r > 0) {
        mw_pass1_kernel<<<dim3((unsigned)((n + 127) / 128), (unsigned)n_slabs), 128, 0, st>>>(h, n, inv_sqrt_m, q, r, partial);
        UMAB_LAUNCH_CHECK();
        mw_reduce_kernel<<<(unsigned)(((long long)r * n + 255) / 256), 256, 0, st>>>(partial, n, r, n_slabs, a);
        UMAB_LAUNCH_CHECK();
        mw_small_kernel<<<1, 256, 0, st>>>(a, q, n, r, b, qb);
        UMAB_LAUNCH_CHECK();
    }
    const int nt = (n + 31) / 32;
    mw_pass2_kernel<<<(unsigned)((long long)nt * (nt + 1) / 2), 256, 0, st>>>(h, n, inv_sqrt_m, q, a, qb, r, nt);
    UMAB_LAUNCH_CHECK();
}

}  // namespace umab
