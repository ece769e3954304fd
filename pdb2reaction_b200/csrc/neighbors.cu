// K1: per-image radius graph, bit-exact against the oracle (oracle/graph.py).
//
// Replaces the on-the-fly graph fairchem rebuilds on every call of the reference
// (pdb2reaction/uma_pysis.py:313-322, `data_list_collater([data], otf_graph=True)`).
// Rule (float32, this exact operation order, no FMA contraction):
//     d2 = (dx*dx + dy*dy) + dz*dz ;  keep iff d2 <= rc^2 and d2 > 1e-4 ;
//     non-strict max_neighbors: if a target has more than `cap` candidates, keep those with
//     d2 <= (cap+1)-th smallest d2 + 0.01.
// One warp per target atom; source positions are staged through shared memory in tiles, so
// every target of a CTA reuses the tile.  Sources are visited in increasing index, so each CSR
// row comes out sorted by source: the output is the canonical (target, source) order without a
// sort.  Two passes (count, fill) around a prefix sum.
#include "kernels.cuh"

namespace umab {

namespace {

constexpr int WARPS = 8;
constexpr int TILE = 256;

__device__ __forceinline__ float dist2_exact(float xi, float yi, float zi, float xj, float yj, float zj) {
    float dx = __fsub_rn(xj, xi), dy = __fsub_rn(yj, yi), dz = __fsub_rn(zj, zi);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// mode 0: count (writes deg[t], thr[t]);  mode 1: fill (reads row_ptr, thr; writes src/tgt)
template <int MODE>
__global__ void __launch_bounds__(WARPS * 32)
neighbor_kernel(const float* __restrict__ pos, int n_atoms, float rc2, int cap,
                int* __restrict__ deg, float* __restrict__ thr,
                const int* __restrict__ row_ptr, int* __restrict__ src, int* __restrict__ tgt) {
    __shared__ float sx[TILE], sy[TILE], sz[TILE];
    const int img = blockIdx.y;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int t_local = blockIdx.x * WARPS + warp;
    const bool active = t_local < n_atoms;
    const long long base = (long long)img * n_atoms;
    const float* p = pos + base * 3;
    float xi = 0.f, yi = 0.f, zi = 0.f;
    if (active) { xi = p[t_local * 3 + 0]; yi = p[t_local * 3 + 1]; zi = p[t_local * 3 + 2]; }
    float limit = rc2;
    int out = 0;
    if (MODE == 1 && active) { limit = thr[base + t_local]; out = row_ptr[base + t_local]; }
    int count = 0;

    for (int s0 = 0; s0 < n_atoms; s0 += TILE) {
        __syncthreads();
        for (int q = threadIdx.x; q < TILE; q += WARPS * 32) {
            int j = s0 + q;
            if (j < n_atoms) { sx[q] = p[j * 3 + 0]; sy[q] = p[j * 3 + 1]; sz[q] = p[j * 3 + 2]; }
        }
        __syncthreads();
        if (!active) continue;
        const int lim = min(TILE, n_atoms - s0);
        for (int q0 = 0; q0 < lim; q0 += 32) {
            int q = q0 + lane;
            bool keep = false;
            if (q < lim) {
                float d2 = dist2_exact(xi, yi, zi, sx[q], sy[q], sz[q]);
                keep = (d2 <= limit) && (d2 > 1e-4f);
            }
            unsigned m = __ballot_sync(0xffffffffu, keep);
            if (MODE == 1 && keep) {
                int o = out + count + __popc(m & ((1u << lane) - 1u));
                src[o] = (int)base + s0 + q;
                tgt[o] = (int)base + t_local;
            }
            count += __popc(m);
        }
    }
    if (MODE == 0 && active) {
        float th = rc2;
        if (count > cap) {
            // (cap+1)-th smallest candidate d2 by bisection on the (monotonic) float bit pattern
            unsigned lo = 0u, hi = __float_as_uint(rc2);
            while (lo < hi) {
                unsigned mid = lo + (hi - lo) / 2u;
                float tv = __uint_as_float(mid);
                int c = 0;
                for (int j0 = 0; j0 < n_atoms; j0 += 32) {
                    int j = j0 + lane;
                    bool k = false;
                    if (j < n_atoms) {
                        float d2 = dist2_exact(xi, yi, zi, p[j * 3 + 0], p[j * 3 + 1], p[j * 3 + 2]);
                        k = (d2 <= tv) && (d2 > 1e-4f);
                    }
                    c += __popc(__ballot_sync(0xffffffffu, k));
                }
                if (c >= cap + 1) hi = mid; else lo = mid + 1u;
            }
            th = fminf(rc2, __fadd_rn(__uint_as_float(lo), 0.01f));
            count = 0;
            for (int j0 = 0; j0 < n_atoms; j0 += 32) {
                int j = j0 + lane;
                bool k = false;
                if (j < n_atoms) {
                    float d2 = dist2_exact(xi, yi, zi, p[j * 3 + 0], p[j * 3 + 1], p[j * 3 + 2]);
                    k = (d2 <= th) && (d2 > 1e-4f);
                }
                count += __popc(__ballot_sync(0xffffffffu, k));
            }
        }
        if (lane == 0) { deg[base + t_local] = count; thr[base + t_local] = th; }
    }
}

// exclusive scan of n ints into out[0..n] (out[n] = total); single CTA, deterministic
__global__ void __launch_bounds__(1024) scan_kernel(const int* __restrict__ in, int* __restrict__ out, int n) {
    __shared__ int part[1024];
    const int t = threadIdx.x;
    const int per = (n + 1023) / 1024;
    const int b = t * per, e = min(n, b + per);
    int s = 0;
    for (int i = b; i < e; ++i) s += in[i];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        int v = (t >= o) ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int run = part[t] - s;
    for (int i = b; i < e; ++i) { out[i] = run; run += in[i]; }
    if (t == 1023) out[n] = part[1023];
}

__global__ void out_degree_kernel(const int* __restrict__ src, int n_edges, int* __restrict__ odeg) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_edges) atomicAdd(&odeg[src[e]], 1);
}

__global__ void out_fill_kernel(const int* __restrict__ src, int n_edges, const int* __restrict__ sptr,
                                int* __restrict__ cursor, int* __restrict__ sedge_tmp) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_edges) {
        int j = src[e];
        int o = atomicAdd(&cursor[j], 1);
        sedge_tmp[sptr[j] + o] = e;
    }
}

// rank-sort each source's edge-id list (ascending) so the by-source reduction order is fixed
__global__ void out_sort_kernel(const int* __restrict__ sptr, const int* __restrict__ tmp,
                                int* __restrict__ sedge, int n_nodes) {
    int node = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    int lane = threadIdx.x % 32;
    if (node >= n_nodes) return;
    int b = sptr[node], e = sptr[node + 1];
    for (int i = b + lane; i < e; i += 32) {
        int v = tmp[i];
        int r = 0;
        for (int k = b; k < e; ++k) r += (tmp[k] < v);
        sedge[b + r] = v;
    }
}

}  // namespace

void launch_neighbor_count(const float* pos, int n_img, int n_atoms, float cutoff, int cap,
                           int* deg, float* thr, cudaStream_t st) {
    dim3 grid((n_atoms + WARPS - 1) / WARPS, n_img);
    neighbor_kernel<0><<<grid, WARPS * 32, 0, st>>>(pos, n_atoms, cutoff * cutoff, cap, deg, thr,
                                                    nullptr, nullptr, nullptr);
    UMAB_LAUNCH_CHECK();
}

void launch_neighbor_fill(const float* pos, int n_img, int n_atoms, float cutoff, int cap,
                          const float* thr, const int* row_ptr, int* src, int* tgt, cudaStream_t st) {
    dim3 grid((n_atoms + WARPS - 1) / WARPS, n_img);
    neighbor_kernel<1><<<grid, WARPS * 32, 0, st>>>(pos, n_atoms, cutoff * cutoff, cap, nullptr,
                                                    const_cast<float*>(thr), row_ptr, src, tgt);
    UMAB_LAUNCH_CHECK();
}

void launch_scan(const int* in, int* out, int n, cudaStream_t st) {
    scan_kernel<<<1, 1024, 0, st>>>(in, out, n);
    UMAB_LAUNCH_CHECK();
}

// by-source CSR: sptr [n_nodes+1], sedge [n_edges] (edge ids ascending within each source)
void launch_source_csr(const int* src, int n_edges, int n_nodes, int* odeg, int* sptr, int* cursor,
                       int* tmp, int* sedge, cudaStream_t st) {
    UMAB_CUDA(cudaMemsetAsync(odeg, 0, sizeof(int) * n_nodes, st));
    UMAB_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * n_nodes, st));
    if (n_edges > 0) {
        out_degree_kernel<<<(n_edges + 255) / 256, 256, 0, st>>>(src, n_edges, odeg);
        UMAB_LAUNCH_CHECK();
    }
    launch_scan(odeg, sptr, n_nodes, st);
    if (n_edges > 0) {
        out_fill_kernel<<<(n_edges + 255) / 256, 256, 0, st>>>(src, n_edges, sptr, cursor, tmp);
        UMAB_LAUNCH_CHECK();
        out_sort_kernel<<<(n_nodes + 7) / 8, 256, 0, st>>>(sptr, tmp, sedge, n_nodes);
        UMAB_LAUNCH_CHECK();
    }
}

}  // namespace umab
