// K1: per-image radius graph, bit-exact against the oracle (oracle/graph.py).
//
// Replaces the on-the-fly graph fairchem rebuilds on every call of the reference
// (pdb2reaction/uma_pysis.py:313-322, `data_list_collater([data], otf_graph=True)`).
// Rule (float32, this exact operation order, no FMA contraction):
//     d2 = (dx*dx + dy*dy) + dz*dz ;  keep iff d2 <= rc^2 and d2 > 1e-4 ;
//     non-strict max_neighbors: if a target has more than `cap` candidates, keep those with
//     d2 <= (cap+1)-th smallest d2 + 0.01.
// Two searches produce the identical edge list:
//  * brute force (small images): one warp per target atom; source positions are staged through shared
//    memory in tiles, so every target of a CTA reuses the tile.  Sources are visited in increasing
//    index, so each CSR row comes out sorted by source without a sort.
//  * shared-memory cell list (default from 128 atoms per image): atoms are binned into cells of edge
//    >= 1.001 r_c (per image: bounding box, histogram, scan, fill); one CTA per (cell, image) stages
//    the atoms of the 27 surrounding cells through shared memory and one warp per target atom of the
//    cell tests them.  The kept sources are marked in a per-warp shared-memory bitmap over the image's
//    atoms and emitted by scanning the bitmap, so the row is sorted by source whatever the order inside
//    the cells (which comes from integer atomics) -- same canonical (target, source) order, no sort.
// Both: two passes (count, fill) around a prefix sum.
#include <mutex>

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
                const int* __restrict__ row_ptr, int* __restrict__ src, int* __restrict__ tgt, int e_cap) {
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
                if (o < e_cap) {           // capacity-sized edge arrays (sync-free path): an overflow is flagged, never written
                    src[o] = (int)base + s0 + q;
                    tgt[o] = (int)base + t_local;
                }
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

// n_edges = launch bound (capacity of the edge arrays); n_dev (nullable) = the actual edge count on the device
__global__ void out_degree_kernel(const int* __restrict__ src, int n_edges, const int* __restrict__ n_dev,
                                  int* __restrict__ odeg) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dev) n_edges = min(n_edges, *n_dev);
    if (e < n_edges) atomicAdd(&odeg[src[e]], 1);
}

// {actual edge count, overflow flag} of a capacity-sized graph build
__global__ void edge_status_kernel(const int* __restrict__ total, int e_cap, int* __restrict__ status) {
    status[0] = *total;
    status[1] = *total > e_cap ? 1 : 0;
}
// after an overflow every CSR walk must still stay inside the capacity-sized edge arrays (the results of such a call
// are discarded and the call repeated): clamp the offsets
__global__ void clamp_row_ptr_kernel(int* __restrict__ row_ptr, int n, int e_cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && row_ptr[i] > e_cap) row_ptr[i] = e_cap;
}

__global__ void out_fill_kernel(const int* __restrict__ src, int n_edges, const int* __restrict__ n_dev,
                                const int* __restrict__ sptr, int* __restrict__ cursor, int* __restrict__ sedge_tmp) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dev) n_edges = min(n_edges, *n_dev);
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

// ------------------------------------------------------------------ cell list
constexpr int CELL_MAX_DIM = 64;

struct CellGrid {          // per image
    float ox, oy, oz;      // origin (bounding-box minimum)
    float ix, iy, iz;      // 1 / cell edge per axis
    int nx, ny, nz;
};

// one CTA per image: bounding box -> grid (cell edge >= 1.001 rc so that two atoms within rc are always in
// adjacent cells despite the rounding of the index arithmetic; edges grow until the grid fits `cap` cells)
__global__ void __launch_bounds__(256)
cell_grid_kernel(const float* __restrict__ pos, int n_atoms, float rc, int cap, CellGrid* __restrict__ grid,
                 int* __restrict__ cell_count) {
    __shared__ float red[6][8];
    const int img = blockIdx.x;
    const float* p = pos + (long long)img * n_atoms * 3;
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int a = threadIdx.x; a < n_atoms; a += blockDim.x)
#pragma unroll
        for (int d = 0; d < 3; ++d) { float v = p[a * 3 + d]; lo[d] = fminf(lo[d], v); hi[d] = fmaxf(hi[d], v); }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
    if (threadIdx.x % 32 == 0)
#pragma unroll
        for (int d = 0; d < 3; ++d) { red[d][threadIdx.x / 32] = lo[d]; red[3 + d][threadIdx.x / 32] = hi[d]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int d = 0; d < 3; ++d)
            for (int w = 0; w < 8; ++w) { lo[d] = fminf(lo[d], red[d][w]); hi[d] = fmaxf(hi[d], red[3 + d][w]); }
        float edge = rc * 1.001f;
        int n[3];
        for (;;) {
            long long tot = 1;
            for (int d = 0; d < 3; ++d) {
                float ext = fmaxf(hi[d] - lo[d], 0.f);
                float c = floorf(ext / edge) + 1.f;
                n[d] = c > (float)CELL_MAX_DIM ? CELL_MAX_DIM + 1 : (int)c;
                tot *= n[d];
            }
            if (tot <= cap && n[0] <= CELL_MAX_DIM && n[1] <= CELL_MAX_DIM && n[2] <= CELL_MAX_DIM) break;
            edge *= 1.26f;
        }
        CellGrid g;
        g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2];
        g.ix = g.iy = g.iz = 1.0f / edge;
        g.nx = n[0]; g.ny = n[1]; g.nz = n[2];
        grid[img] = g;
    }
    for (int c = threadIdx.x; c < cap; c += blockDim.x) cell_count[(long long)img * cap + c] = 0;
}

__device__ __forceinline__ int cell_coord(float x, float o, float inv, int n) {
    int c = (int)floorf((x - o) * inv);
    return min(max(c, 0), n - 1);
}

// histogram of the atoms over the cells of their image; remembers each atom's cell
__global__ void cell_histogram_kernel(const float* __restrict__ pos, int n_atoms, int n_img, int cap,
                                      const CellGrid* __restrict__ grid, int* __restrict__ cell_count,
                                      int* __restrict__ atom_cell) {
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= (long long)n_img * n_atoms) return;
    const int img = (int)(a / n_atoms);
    const CellGrid g = grid[img];
    const int cx = cell_coord(pos[a * 3 + 0], g.ox, g.ix, g.nx);
    const int cy = cell_coord(pos[a * 3 + 1], g.oy, g.iy, g.ny);
    const int cz = cell_coord(pos[a * 3 + 2], g.oz, g.iz, g.nz);
    const int c = (cz * g.ny + cy) * g.nx + cx;
    atom_cell[a] = c;
    atomicAdd(&cell_count[(long long)img * cap + c], 1);
}

// per image: exclusive scan of the cell counts -> cell_start[img][0..cap]; counts are zeroed (reused as cursors)
__global__ void __launch_bounds__(1024)
cell_scan_kernel(int* __restrict__ cell_count, int cap, int* __restrict__ cell_start) {
    __shared__ int part[1024];
    const int img = blockIdx.x;
    int* in = cell_count + (long long)img * cap;
    int* out = cell_start + (long long)img * (cap + 1);
    const int t = threadIdx.x;
    const int per = (cap + 1023) / 1024;
    const int b = min(cap, t * per), e = min(cap, b + per);
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
    for (int i = b; i < e; ++i) { out[i] = run; run += in[i]; in[i] = 0; }
    if (t == 1023) out[cap] = part[1023];
}

// atoms of every cell, contiguous per cell (order inside a cell is arbitrary; the search does not depend on it)
__global__ void cell_fill_kernel(int n_atoms, int n_img, int cap, const int* __restrict__ atom_cell,
                                 const int* __restrict__ cell_start, int* __restrict__ cursor,
                                 int* __restrict__ cell_atoms) {
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= (long long)n_img * n_atoms) return;
    const int img = (int)(a / n_atoms);
    const int c = atom_cell[a];
    const int o = atomicAdd(&cursor[(long long)img * cap + c], 1);
    cell_atoms[(long long)img * n_atoms + cell_start[(long long)img * (cap + 1) + c] + o] = (int)(a - (long long)img * n_atoms);
}

// one CTA per (cell, image).  mode 0: count (deg, thr);  mode 1: fill (src, tgt) through the per-warp bitmap.
template <int MODE>
__global__ void __launch_bounds__(WARPS * 32)
neighbor_cell_kernel(const float* __restrict__ pos, int n_atoms, float rc2, int cap_nb, int cap_cells,
                     const CellGrid* __restrict__ grid, const int* __restrict__ cell_start,
                     const int* __restrict__ cell_atoms, int* __restrict__ deg, float* __restrict__ thr,
                     const int* __restrict__ row_ptr, int* __restrict__ src, int* __restrict__ tgt, int e_cap) {
    extern __shared__ unsigned bitmap[];                       // MODE 1: WARPS x ceil(n_atoms / 32) words
    __shared__ float sx[TILE], sy[TILE], sz[TILE];
    __shared__ int sid[TILE];
    __shared__ int rng_start[27], rng_off[28];                 // candidate ranges of the 27 cells + prefix sums
    const int img = blockIdx.y;
    const CellGrid g = grid[img];
    const int c = blockIdx.x;
    if (c >= g.nx * g.ny * g.nz) return;
    const int* cs = cell_start + (long long)img * (cap_cells + 1);
    const int t_begin = cs[c], t_end = cs[c + 1];
    if (t_begin == t_end) return;
    const int* ca = cell_atoms + (long long)img * n_atoms;
    const long long base = (long long)img * n_atoms;
    const float* p = pos + base * 3;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (threadIdx.x < 27) {
        const int cx = c % g.nx, cy = (c / g.nx) % g.ny, cz = c / (g.nx * g.ny);
        const int dx = threadIdx.x % 3 - 1, dy = (threadIdx.x / 3) % 3 - 1, dz = threadIdx.x / 9 - 1;
        const int x = cx + dx, y = cy + dy, z = cz + dz;
        int b = 0, n = 0;
        if (x >= 0 && x < g.nx && y >= 0 && y < g.ny && z >= 0 && z < g.nz) {
            const int cc = (z * g.ny + y) * g.nx + x;
            b = cs[cc];
            n = cs[cc + 1] - b;
        }
        rng_start[threadIdx.x] = b;
        rng_off[threadIdx.x + 1] = n;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        rng_off[0] = 0;
        for (int i = 0; i < 27; ++i) rng_off[i + 1] += rng_off[i];
    }
    __syncthreads();
    const int n_cand = rng_off[27];
    // flat candidate position -> atom index (local to the image)
    auto cand_atom = [&](int q) {
        int r = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1)
            if (r + step < 27 && rng_off[r + step] <= q) r += step;
        return ca[rng_start[r] + (q - rng_off[r])];
    };
    const int n_words = (n_atoms + 31) / 32;
    unsigned* bm = bitmap + (size_t)warp * n_words;

    for (int t0 = t_begin; t0 < t_end; t0 += WARPS) {          // groups of WARPS targets of this cell
        const int ti = t0 + warp;
        const bool active = ti < t_end;
        const int t_local = active ? ca[ti] : 0;
        float xi = 0.f, yi = 0.f, zi = 0.f;
        if (active) { xi = p[t_local * 3 + 0]; yi = p[t_local * 3 + 1]; zi = p[t_local * 3 + 2]; }
        float limit = rc2;
        if (MODE == 1 && active) {
            limit = thr[base + t_local];
            for (int w = lane; w < n_words; w += 32) bm[w] = 0u;
            __syncwarp();
        }
        int count = 0;
        for (int s0 = 0; s0 < n_cand; s0 += TILE) {
            __syncthreads();
            for (int q = threadIdx.x; q < TILE && s0 + q < n_cand; q += WARPS * 32) {
                const int j = cand_atom(s0 + q);
                sid[q] = j; sx[q] = p[j * 3 + 0]; sy[q] = p[j * 3 + 1]; sz[q] = p[j * 3 + 2];
            }
            __syncthreads();
            if (!active) continue;
            const int lim = min(TILE, n_cand - s0);
            for (int q0 = 0; q0 < lim; q0 += 32) {
                const int q = q0 + lane;
                bool keep = false;
                if (q < lim) {
                    const float d2 = dist2_exact(xi, yi, zi, sx[q], sy[q], sz[q]);
                    keep = (d2 <= limit) && (d2 > 1e-4f);
                }
                if (MODE == 1 && keep) atomicOr(&bm[sid[q] >> 5], 1u << (sid[q] & 31));
                count += __popc(__ballot_sync(0xffffffffu, keep));
            }
        }
        if (MODE == 0 && active) {
            float th = rc2;
            if (count > cap_nb) {
                // (cap+1)-th smallest candidate d2 by bisection on the (monotonic) float bit pattern
                auto count_le = [&](float tv) {
                    int cnt = 0;
                    for (int q0 = 0; q0 < n_cand; q0 += 32) {
                        const int q = q0 + lane;
                        bool k = false;
                        if (q < n_cand) {
                            const int j = cand_atom(q);
                            const float d2 = dist2_exact(xi, yi, zi, p[j * 3 + 0], p[j * 3 + 1], p[j * 3 + 2]);
                            k = (d2 <= tv) && (d2 > 1e-4f);
                        }
                        cnt += __popc(__ballot_sync(0xffffffffu, k));
                    }
                    return cnt;
                };
                unsigned lo = 0u, hi = __float_as_uint(rc2);
                while (lo < hi) {
                    const unsigned mid = lo + (hi - lo) / 2u;
                    if (count_le(__uint_as_float(mid)) >= cap_nb + 1) hi = mid; else lo = mid + 1u;
                }
                th = fminf(rc2, __fadd_rn(__uint_as_float(lo), 0.01f));
                count = count_le(th);
            }
            if (lane == 0) { deg[base + t_local] = count; thr[base + t_local] = th; }
        }
        if (MODE == 1 && active) {
            __syncwarp();
            int out = row_ptr[base + t_local];
            for (int w0 = 0; w0 < n_words; w0 += 32) {
                const int w = w0 + lane;
                const unsigned bits = w < n_words ? bm[w] : 0u;
                int pc = __popc(bits), incl = pc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                int o = out + incl - pc;
                unsigned b2 = bits;
                while (b2) {
                    const int bit = __ffs(b2) - 1;
                    b2 &= b2 - 1u;
                    if (o < e_cap) {
                        src[o] = (int)base + w * 32 + bit;
                        tgt[o] = (int)base + t_local;
                    }
                    ++o;
                }
                out += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
    }
}

}  // namespace

// cell list of every image: grid [n_img], cell_start [n_img, cap+1], cell_atoms [n_img, n_atoms];
// scratch: cell_count [n_img, cap], atom_cell [n_img * n_atoms]
void launch_cell_list(const float* pos, int n_img, int n_atoms, float cutoff, int cap_cells, void* grid,
                      int* cell_count, int* atom_cell, int* cell_start, int* cell_atoms, cudaStream_t st) {
    const long long n = (long long)n_img * n_atoms;
    CellGrid* g = reinterpret_cast<CellGrid*>(grid);
    cell_grid_kernel<<<n_img, 256, 0, st>>>(pos, n_atoms, cutoff, cap_cells, g, cell_count);
    UMAB_LAUNCH_CHECK();
    cell_histogram_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pos, n_atoms, n_img, cap_cells, g, cell_count, atom_cell);
    UMAB_LAUNCH_CHECK();
    cell_scan_kernel<<<n_img, 1024, 0, st>>>(cell_count, cap_cells, cell_start);
    UMAB_LAUNCH_CHECK();
    cell_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n_atoms, n_img, cap_cells, atom_cell, cell_start, cell_count, cell_atoms);
    UMAB_LAUNCH_CHECK();
}
size_t cell_grid_bytes() { return sizeof(CellGrid); }

void launch_neighbor_cell_count(const float* pos, int n_img, int n_atoms, float cutoff, int cap, int cap_cells,
                                const void* grid, const int* cell_start, const int* cell_atoms, int* deg, float* thr,
                                cudaStream_t st) {
    dim3 g(cap_cells, n_img);
    neighbor_cell_kernel<0><<<g, WARPS * 32, 0, st>>>(pos, n_atoms, cutoff * cutoff, cap, cap_cells,
                                                      reinterpret_cast<const CellGrid*>(grid), cell_start, cell_atoms, deg,
                                                      thr, nullptr, nullptr, nullptr, 0);
    UMAB_LAUNCH_CHECK();
}

void launch_neighbor_cell_fill(const float* pos, int n_img, int n_atoms, float cutoff, int cap, int cap_cells,
                               const void* grid, const int* cell_start, const int* cell_atoms, const float* thr,
                               const int* row_ptr, int* src, int* tgt, int e_cap, cudaStream_t st) {
    dim3 g(cap_cells, n_img);
    const size_t smem = (size_t)WARPS * ((n_atoms + 31) / 32) * sizeof(unsigned);
    // the opt-in above 48 KB is a per-device attribute; engines on several GPUs run on their own host threads
    static std::mutex mu;
    static size_t configured[64] = {0};
    if (smem > 227 * 1024) throw CudaError("neighbour search: image too large for the shared-memory bitmap (> ~226k atoms)");
    if (smem > 40 * 1024) {
        int dev = 0;
        UMAB_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(mu);
        if (smem > configured[dev & 63]) {
            UMAB_CUDA(cudaFuncSetAttribute(neighbor_cell_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured[dev & 63] = smem;
        }
    }
    neighbor_cell_kernel<1><<<g, WARPS * 32, smem, st>>>(pos, n_atoms, cutoff * cutoff, cap, cap_cells,
                                                         reinterpret_cast<const CellGrid*>(grid), cell_start, cell_atoms,
                                                         nullptr, const_cast<float*>(thr), row_ptr, src, tgt, e_cap);
    UMAB_LAUNCH_CHECK();
}

void launch_neighbor_count(const float* pos, int n_img, int n_atoms, float cutoff, int cap,
                           int* deg, float* thr, cudaStream_t st) {
    dim3 grid((n_atoms + WARPS - 1) / WARPS, n_img);
    neighbor_kernel<0><<<grid, WARPS * 32, 0, st>>>(pos, n_atoms, cutoff * cutoff, cap, deg, thr,
                                                    nullptr, nullptr, nullptr, 0);
    UMAB_LAUNCH_CHECK();
}

void launch_neighbor_fill(const float* pos, int n_img, int n_atoms, float cutoff, int cap,
                          const float* thr, const int* row_ptr, int* src, int* tgt, int e_cap, cudaStream_t st) {
    dim3 grid((n_atoms + WARPS - 1) / WARPS, n_img);
    neighbor_kernel<1><<<grid, WARPS * 32, 0, st>>>(pos, n_atoms, cutoff * cutoff, cap, nullptr,
                                                    const_cast<float*>(thr), row_ptr, src, tgt, e_cap);
    UMAB_LAUNCH_CHECK();
}

void launch_scan(const int* in, int* out, int n, cudaStream_t st) {
    scan_kernel<<<1, 1024, 0, st>>>(in, out, n);
    UMAB_LAUNCH_CHECK();
}

// by-source CSR: sptr [n_nodes+1], sedge [n_edges] (edge ids ascending within each source)
__global__ void fill_int_kernel(int* __restrict__ p, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
void launch_fill_int(int* p, int n, int v, cudaStream_t st) {
    if (n <= 0) return;
    fill_int_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, n, v);
    UMAB_LAUNCH_CHECK();
}

void launch_edge_status(int* row_ptr, int n_nodes, int e_cap, int* status_dev, cudaStream_t st) {
    edge_status_kernel<<<1, 1, 0, st>>>(row_ptr + n_nodes, e_cap, status_dev);
    UMAB_LAUNCH_CHECK();
    clamp_row_ptr_kernel<<<(n_nodes + 1 + 255) / 256, 256, 0, st>>>(row_ptr, n_nodes + 1, e_cap);
    UMAB_LAUNCH_CHECK();
}

// n_edges: launch bound (edge-array capacity); n_edges_dev (nullable): actual count on the device
void launch_source_csr(const int* src, int n_edges, const int* n_edges_dev, int n_nodes, int* odeg, int* sptr, int* cursor,
                       int* tmp, int* sedge, cudaStream_t st) {
    UMAB_CUDA(cudaMemsetAsync(odeg, 0, sizeof(int) * n_nodes, st));
    UMAB_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * n_nodes, st));
    if (n_edges > 0) {
        out_degree_kernel<<<(n_edges + 255) / 256, 256, 0, st>>>(src, n_edges, n_edges_dev, odeg);
        UMAB_LAUNCH_CHECK();
    }
    launch_scan(odeg, sptr, n_nodes, st);
    if (n_edges > 0) {
        out_fill_kernel<<<(n_edges + 255) / 256, 256, 0, st>>>(src, n_edges, n_edges_dev, sptr, cursor, tmp);
        UMAB_LAUNCH_CHECK();
        out_sort_kernel<<<(n_nodes + 7) / 8, 256, 0, st>>>(sptr, tmp, sedge, n_nodes);
        UMAB_LAUNCH_CHECK();
    }
}

}  // namespace umab
