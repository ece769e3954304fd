// Engine: owns weights, graph, node state and the chunked edge workspace; runs the batched
// forward (energy) and hand-written backward (forces) over all images of a call and exposes
// the C ABI declared in include/umab.h.
//
// Replaces, for a whole batch of images at once, what the reference does per image in
// UMAcore.compute (pdb2reaction/uma_pysis.py:330-419): graph build, predict, force readout.
// Memory model (180 GB HBM3e): per-layer node features are kept for the backward
// ([N,9,128] fp32 each); per-edge activations are never kept across layers -- the backward
// recomputes them chunk by chunk (chunks = contiguous target-node ranges of the CSR), so the
// edge workspace is bounded by `workspace_bytes` regardless of batch size.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/umab.h"
#include "kernels.cuh"

namespace umab {

std::atomic<long long> g_launch_count{0};
// bumped whenever any device / pinned buffer of the library is (re)allocated or freed: a captured CUDA graph bakes the
// buffer addresses in, so it is replayed only while the generation it was captured under is still current
std::atomic<long long> g_alloc_gen{0};
struct TcPlaneCache;                                  // gemm_tc.cu
TcPlaneCache* tc_cache_create();
void tc_cache_destroy(TcPlaneCache* c);
void tc_cache_clear(TcPlaneCache* c);
void gemm_tc(const GemmArgs& a, cudaStream_t st, TcPlaneCache* cache);
bool gemm_tc_supported(const GemmArgs& a);
struct Tc2Cache;                                     // gemm_tc2.cu
Tc2Cache* tc2_cache_create();
void tc2_cache_destroy(Tc2Cache* c);
void tc2_cache_clear(Tc2Cache* c);
void gemm_tc2(const GemmArgs& a, cudaStream_t st, Tc2Cache* cache, int bk, int pair = -1);
bool gemm_tc2_supported(const GemmArgs& a, int bk);
bool gemm_tc2_gate_epilogue_available();
void tc2_split(const float* x, long long n, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st);
void launch_fd_columns(const float* f, const int* ks, int n_cols, int dof, double h_step, void* hmat, long long ldh,
                       bool f64, cudaStream_t st);                                         // hessian_ops.cu
size_t mw_project_workspace_doubles(int n, int r);
void launch_mw_project(double* h, int n, const double* inv_sqrt_m, const double* q, int r, double* ws, cudaStream_t st);

namespace {

thread_local std::string g_last_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    static std::atomic<long long> total;       // engines on several GPUs allocate from their own host threads
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if (p) { cudaFree(p); total -= (long long)cap; }
        p = nullptr; cap = 0;
        ++g_alloc_gen;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            p = nullptr;
            char buf[256];
            snprintf(buf, sizeof buf, "CUDA out of memory: cudaMalloc of %zu bytes failed (%s)", bytes, cudaGetErrorString(e));
            throw CudaError(buf);
        }
        cap = want;
        total += (long long)want;
    }
    void release() { if (p) { cudaFree(p); total -= (long long)cap; ++g_alloc_gen; } p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
    float* f() const { return reinterpret_cast<float*>(p); }
    int* i() const { return reinterpret_cast<int*>(p); }
};
std::atomic<long long> DevBuf::total{0};

// a tensor of the pipeline: value plane + (Hessian path only) tangent plane
struct TBuf {
    DevBuf v, d;
    template <class S> void ensure(size_t bytes) {
        v.ensure(bytes);
        if (std::is_same<S, D1>::value) d.ensure(bytes);
    }
    void release() { v.release(); d.release(); }
};
template <class S> GP<S> gp(const TBuf& b, long long off = 0);
template <> GP<float> gp<float>(const TBuf& b, long long off) { return GP<float>{b.v.f() + off}; }
template <> GP<D1> gp<D1>(const TBuf& b, long long off) { return GP<D1>{b.v.f() + off, b.d.f() + off}; }
template <class S> GP<S> gnull();
template <> GP<float> gnull<float>() { return GP<float>{nullptr}; }
template <> GP<D1> gnull<D1>() { return GP<D1>{nullptr, nullptr}; }
inline void value_rows(AP<float>&, long long) {}
inline void value_rows(AP<D1>& a, long long elems) { a.vlim = elems; }
template <class S> AP<S> anull();
template <> AP<float> anull<float>() { return AP<float>{nullptr, nullptr, nullptr}; }
template <> AP<D1> anull<D1>() { return AP<D1>{{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}}; }
template <class S> constexpr int planes() { return std::is_same<S, D1>::value ? 2 : 1; }
inline float* plane(const GP<float>& g, int) { return g.p; }
inline float* plane(const GP<D1>& g, int k) { return k == 0 ? g.v : g.d; }
inline AP<float> plane(const AP<float>& a, int) { return a; }
inline AP<float> plane(const AP<D1>& a, int k) { return k == 0 ? a.v : a.d; }
// A-operand view of `cap_elems` elements at float offset `off` of a workspace buffer: fp32, or bf16 hi | lo planes
inline AP<float> apf(float* base, long long off, long long cap_elems, bool split) {
    if (!split) return AP<float>{base + off, nullptr, nullptr};
    __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(base + off);
    return AP<float>{nullptr, hi, hi + cap_elems};
}
template <class S> AP<S> ap(const TBuf& b, long long off, long long cap_elems, bool split);
template <> AP<float> ap<float>(const TBuf& b, long long off, long long cap_elems, bool split) {
    return apf(b.v.f(), off, cap_elems, split);
}
template <> AP<D1> ap<D1>(const TBuf& b, long long off, long long cap_elems, bool split) {
    return AP<D1>{apf(b.v.f(), off, cap_elems, split), apf(b.d.f(), off, cap_elems, split)};
}
struct Weight { DevBuf buf; size_t numel = 0; };

struct Chunk { int node0, n_nodes; long long e0; int n_e; };

struct RadialW {
    const float *w1g, *w1g_t, *t_src, *t_tgt, *b1, *ln1w, *ln1b, *w2, *w2_t, *b2, *ln2w, *ln2b, *w3, *w3_t, *b3;
    int n_out;
};

struct LayerW {
    const float *n1w, *n1b, *n2w, *n2b;
    RadialW rad;
    const float *c1m0, *c1m0_t, *c1m0_b, *c1m1, *c1m1_t, *c1m2, *c1m2_t;
    const float *c2m0, *c2m0_t, *c2m0_b, *c2m1, *c2m1_t, *c2m2, *c2m2_t;
    const float *smlp, *smlp_t, *smlp_b, *so3_1, *so3_1_t, *so3_1_b, *so3_2, *so3_2_t, *so3_2_b;
};

// ---- per-kernel-family timing with CUDA events on the launching stream (bench.py roofline)
enum ProfCat { P_GEMM = 0, P_GATHER, P_GATHER_BWD, P_COMBINE, P_COMBINE_BWD, P_ROTBACK, P_ROTBACK_BWD,
               P_SRC_REDUCE, P_LN_SILU, P_NODE, P_GRAPH, P_GEOMETRY, P_COUNT };
const char* const kProfNames[P_COUNT] = {"gemm", "gather_rotate_scale", "gather_rotate_bwd", "combine_gate_fwd",
                                         "combine_gate_bwd", "rotate_back_reduce", "rotate_back_bwd",
                                         "source_reduce", "ln_silu", "node_ops", "graph", "geometry"};
struct Prof {
    bool on = false;
    struct Rec { int cat; cudaEvent_t a, b; double work, bytes; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    double ms[P_COUNT] = {0}; double work[P_COUNT] = {0}; double bytes[P_COUNT] = {0}; long long n[P_COUNT] = {0};
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; UMAB_CUDA(cudaEventCreate(&e)); return e;
    }
    void collect() {
        for (auto& r : recs) {
            UMAB_CUDA(cudaEventSynchronize(r.b));
            float t = 0.f;
            UMAB_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
            ms[r.cat] += t; work[r.cat] += r.work; bytes[r.cat] += r.bytes; n[r.cat] += 1;
            pool.push_back(r.a); pool.push_back(r.b);
        }
        recs.clear();
    }
    void reset() { collect(); for (int i = 0; i < P_COUNT; ++i) { ms[i] = 0; work[i] = 0; bytes[i] = 0; n[i] = 0; } }
};

// switch to the engine's device for the duration of a C-ABI call and restore the caller's
// current device afterwards (the host may be driving several engines / torch on other GPUs)
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        UMAB_CUDA(cudaSetDevice(dev));
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

constexpr int YW = 640 + 512 + 256;      // conv-1 outputs per edge: m=0 | m=1 (o_r|o_i) | m=2 (p_r|p_i)
constexpr int ZW = 384 + 512 + 256;      // conv-2 outputs per edge
constexpr size_t EDGE_WS_FLOATS = 2304 + 1536 /* wY: max(YW, g_rad) */ + 1152 + ZW + 1536 + 4 * 128;   // 8192 per edge
constexpr size_t EDGE_WS_EXTRA_RECOMPUTE = YW + ZW;
constexpr int STORE_FLOATS = YW + ZW + 1536 + 2 * 128;   // per edge and layer in store mode: Y, Z, rad, u1, u2   // adjoint operands g_Y / g_Z when Y / Z live in the workspace

}  // namespace
}  // namespace umab

using namespace umab;

struct umab_engine {
    umab_config cfg;
    std::unordered_map<std::string, Weight> weights;
    bool finalized = false;
    // resolved weights
    const float *sphere_emb = nullptr, *csd = nullptr, *normw = nullptr, *normb = nullptr;
    const float *h0 = nullptr, *h0_t = nullptr, *h0_b = nullptr, *h2 = nullptr, *h2_t = nullptr, *h2_b = nullptr,
                *h4 = nullptr, *h4_b = nullptr;
    RadialW ed_rad{};
    std::vector<LayerW> layers;

    // system
    int n_atoms = 0;
    DevBuf z1;
    // batch
    int n_img = 0, n_nodes = 0;
    long long n_edges = 0;
    int zt_img = -1;
    DevBuf pos_own, zt, deg, thr, row_ptr, src, tgt, odeg, sptr, cursor, stmp, sedge;
    DevBuf cgrid, ccount, cstart, catoms, acell;       // cell list of the neighbour search
    int neighbor_mode = 0;                             // 0 auto (cell list from 128 atoms per image), 1 brute force, 2 cell list
    int* h_pinned = nullptr; size_t h_pinned_cap = 0;
    // ---- sync-free graph build (single-chunk calls): the edge arrays are sized from a CAPACITY -- the complete graph
    // for images below 128 atoms (cannot overflow), else the largest edges-per-image seen so far + 3 % -- instead of
    // the row_ptr read-back; every per-edge kernel runs over the capacity (padding rows hold valid indices and are
    // never reduced: all reductions walk the CSR), the by-source CSR takes the actual count from the device, and an
    // overflow is flagged on the device, checked when the results are collected, and the call repeated.
    long long hist_edges_per_image = 0;
    bool fast_graph = false;                           // this evaluation ran on capacity-sized edge arrays
    long long n_edges_actual = 0;                      // true edge count of the last evaluation (after resolve_status)
    DevBuf status_dev;
    int* h_status = nullptr;                           // pinned {actual edges, overflow}
    cudaEvent_t status_ev = nullptr;
    bool status_pending = false;
    int status_nimg = 0;
    long long total_mem = 0;                           // device memory size (store budget), queried once
    // last public call (possibly several sub-batches)
    long long call_images = 0, call_edges = 0, call_subcalls = 0;
    // ---- CUDA-graph replay of the host-buffer entry point for launch-bound calls
    struct GraphEntry { cudaGraphExec_t exec = nullptr; int nimg = 0; long long fcap = 0; bool want_f = false;
                        long long gen = -1; int warm = 0; int n_nodes = 0; long long n_launch = 0; };
    std::vector<GraphEntry> graphs;
    // per-engine switches (umab_set_option "nosync" / "cuda_graphs"); defaults from UMAB_NOSYNC / UMAB_CUDA_GRAPHS
    bool opt_nosync = [] { const char* e = getenv("UMAB_NOSYNC"); return !(e && atoi(e) == 0); }();
    bool opt_graphs = [] { const char* e = getenv("UMAB_CUDA_GRAPHS"); return !(e && atoi(e) == 0); }();
    bool nosync_enabled() const { return opt_nosync; }
    bool graphs_enabled() const { return opt_graphs; }
    long long graph_replays = 0, graph_captures = 0, overflow_retries = 0, fast_calls = 0;
    // Hessian columns of ONE base geometry (umab_set_option "jvp_shared_base"; calculator.hessian_columns): every image
    // of the dual-number batch has the same positions, so the VALUE plane of every tensor is the same for all images.
    // The value-plane GEMMs then run on the rows of ONE image and the result block is copied to the others (a streaming
    // write instead of a GEMM); the tangent planes and all elementwise kernels are untouched, so the result bits are the
    // same as without the option (GEMM rows are independent of M).  Checked on the device per call (same_images).
    bool opt_shared_base = false;
    bool dedupe = false;                               // this call: positions of all images verified identical
    long long rep_rows = 0;                            // rows per image of the GEMMs being issued (0: do not de-duplicate)
    long long e_img = 0;                               // edges per image (dedupe)
    long long dedupe_gemms = 0;                        // value-plane GEMMs run on one image (umab_get_option "dedupe_gemms")
    DevBuf same_flag;
    // Second step (UMAB_SHARE_VALUES, default on): the edge kernels that consume those GEMM outputs read the value rows of
    // the FIRST image of the chunk (ImgShare, dual.cuh) in an order that lets the images of a batch share them in L2, so
    // the block is not even copied: per edge the value planes cross HBM once per batch instead of once per column.
    bool opt_share = [] { const char* e = getenv("UMAB_SHARE_VALUES"); return !(e && atoi(e) == 0); }();
    bool cur_share = false;                            // inside a chunk whose edge kernels run with ImgShare
    // value-plane GEMM on one image (+ replication unless its consumers share the block), or the plain GEMM
    void gemm_plane(GemmArgs& g, int k, long long M, long long block_ld, cudaStream_t st, bool consumers_share = false) {
        if (k == 0 && rep_rows > 0 && M > rep_rows && M % rep_rows == 0 && !g.gate) {
            g.M = (int)rep_rows;
            gemm(g, st);
            if (!(consumers_share && cur_share)) launch_replicate_block(g.Cmat, rep_rows * block_ld, (int)(M / rep_rows), st);
            ++dedupe_gemms;
        } else {
            g.M = (int)M;
            gemm(g, st);
        }
    }
    struct RepScope {                                  // rows per image inside a chunk loop, restored on exit
        umab_engine* e; long long saved; bool saved_share;
        RepScope(umab_engine* e_, long long rows, bool share) : e(e_), saved(e_->rep_rows), saved_share(e_->cur_share) {
            e->rep_rows = rows; e->cur_share = share;
        }
        ~RepScope() { e->rep_rows = saved; e->cur_share = saved_share; }
    };
    long long chunk_rep_rows() const { return dedupe && chunks_closed ? e_img : 0; }
    // images of a chunk whose edge kernels share the value planes (0: the plain launch)
    int chunk_images(const Chunk& c) const {
        if (!dedupe || !opt_share || !chunks_closed || n_atoms <= 0 || c.n_nodes % n_atoms) return 0;
        const int k = c.n_nodes / n_atoms;
        return (k > 1 && (long long)c.n_e == (long long)k * e_img) ? k : 0;
    }
    ImgShare share_e(const Chunk& c) const { return ImgShare{chunk_images(c), (int)e_img}; }    // one warp per edge
    ImgShare share_n(const Chunk& c) const { return ImgShare{chunk_images(c), n_atoms}; }       // one warp per node
    std::vector<Chunk> chunks;
    bool chunks_closed = false;                        // chunks hold whole images (see plan_chunks)
    static bool store_radial_enabled() {               // UMAB_STORE_RADIAL=0: recompute the radial MLP in the backward (A/B)
        static const bool on = [] { const char* e = getenv("UMAB_STORE_RADIAL"); return !(e && atoi(e) == 0); }();
        return on;
    }
    static bool closed_chunks_enabled() {
        static const bool on = [] { const char* e = getenv("UMAB_CLOSED_CHUNKS"); return !(e && atoi(e) == 0); }();
        return on;
    }
    // geometry
    TBuf vec, dist, env, wig, gauss, g_gauss, g_env, g_wig, g_vec;      // g_wig: per-edge torque (t_x, t_z, 0, 0) of the edge frame
    // nodes
    std::vector<TBuf> xs, x1s, y1s, gps;
    TBuf nbuf, abuf, gx, gx1, gn, ggp, p1, s1, p2, gp2, gs1, Gbuf;
    DevBuf node_e;
    // edge workspace
    TBuf wA, wY, wB, wZ, wRAD, wU1, wH1, wU2, wH2, wGY, wGZ;
    DevBuf wSG;                                        // sigmoid(gates) [chunk, 256] of the fused-gate path
    // OFF by default: measured (round 2, two boxes) combine_gate_fwd 19.9 -> 9.9 ms but GEMMs +10.6 .. +15.3 ms per C4 step
    // (net +4 .. +5 ms): the epilogue moves 9.5 KB per edge and layer where the separate kernel moved 10.2 KB, and under
    // the 1000 W cap a byte costs the same inside the GEMM epilogue.  Bit-identical either way.
    bool opt_fuse_gate = [] { const char* e = getenv("UMAB_FUSE_GATE"); return e && atoi(e) != 0; }();
    long long chunk_cap = 0;
    // store mode: conv-1 / conv-2 outputs of every layer are kept for the backward instead of being
    // recomputed (16.4 KB per edge and layer of HBM) when they fit `store_bytes`
    std::vector<TBuf> ystore, zstore;
    // ... and so are the radial weights rad [E,1536] and the two radial pre-activations u1 / u2 [E,128]: the backward
    // then recomputes nothing at all (7.2 KB more per edge and layer)
    std::vector<TBuf> rstore, u1store, u2store;
    bool store_mode = false;
    bool want_adjoint = false;                        // this evaluation runs the backward (forces requested)
    // host staging for umab_energy_forces_host
    DevBuf e_dev, f_dev, t_dev, df_dev;
    // debug
    std::map<std::string, std::pair<DevBuf, size_t>> dbg;
    TcPlaneCache* tc_cache = tc_cache_create();       // bf16 planes of this engine's weights
    Tc2Cache* tc2_cache = tc2_cache_create();         // same for the TMA-fed GEMM (+ tensor maps of the workspace)
    // profiling + pinned host staging
    Prof prof;
    void* hp_pos = nullptr; void* hp_f = nullptr; void* hp_e = nullptr; size_t hp_cap = 0, hp_ecap = 0;

    long long device_memory() {
        if (!total_mem) {
            size_t fr = 0, tot = 0;
            UMAB_CUDA(cudaMemGetInfo(&fr, &tot));
            total_mem = (long long)tot;
        }
        return total_mem;
    }
    // images one evaluation should take: bounded by the node state (~100 KB per atom) and, when forces are wanted, by
    // the per-layer stores (so that the backward recomputes nothing); larger batches run as consecutive sub-batches
    int images_per_call(bool forces, int nplanes) {
        const long long max_atoms = 49152 / nplanes;
        long long cap = std::max<long long>(1, max_atoms / std::max(n_atoms, 1));
        if (forces && cfg.store_bytes >= 0) {
            const double budget = cfg.store_bytes > 0 ? (double)cfg.store_bytes : 0.45 * (double)device_memory();
            const double epi = hist_edges_per_image > 0 ? 1.03 * (double)hist_edges_per_image : 85.0 * n_atoms;
            const double per_image = epi * (store_radial_enabled() ? STORE_FLOATS : YW + ZW) * 4.0 * cfg.num_layers * nplanes;
            const long long fit = (long long)(budget / per_image);
            if (fit >= 1) cap = std::min(cap, fit);
        }
        return (int)cap;
    }
    template <class F> void timed(int cat, double work, cudaStream_t st, F&& f, double bytes = -1.0) {
        if (!prof.on) { f(); return; }
        Prof::Rec r{cat, prof.get(), prof.get(), work, bytes < 0 ? work : bytes};
        UMAB_CUDA(cudaEventRecord(r.a, st));
        f();
        UMAB_CUDA(cudaEventRecord(r.b, st));
        prof.recs.push_back(r);
    }

    const float* W(const std::string& name, size_t numel) {
        auto it = weights.find(name);
        if (it == weights.end()) throw CudaError("missing weight: " + name);
        if (it->second.numel != numel) {
            char buf[256];
            snprintf(buf, sizeof buf, "weight %s has %zu elements, expected %zu", name.c_str(), it->second.numel, numel);
            throw CudaError(buf);
        }
        return it->second.buf.f();
    }

    RadialW radial(const std::string& p, int n_out) {
        RadialW r;
        r.w1g = W(p + ".w1g", 128 * NB); r.w1g_t = W(p + ".w1g_t", NB * 128);
        r.t_src = W(p + ".t_src", 100 * 128); r.t_tgt = W(p + ".t_tgt", 100 * 128);
        r.b1 = W(p + ".lin1.bias", 128); r.ln1w = W(p + ".ln1.weight", 128); r.ln1b = W(p + ".ln1.bias", 128);
        r.w2 = W(p + ".lin2.weight", 128 * 128); r.w2_t = W(p + ".lin2.weight_t", 128 * 128);
        r.b2 = W(p + ".lin2.bias", 128); r.ln2w = W(p + ".ln2.weight", 128); r.ln2b = W(p + ".ln2.bias", 128);
        r.w3 = W(p + ".lin3.weight", (size_t)n_out * 128); r.w3_t = W(p + ".lin3.weight_t", (size_t)n_out * 128);
        r.b3 = W(p + ".lin3.bias", n_out);
        r.n_out = n_out;
        return r;
    }

    void finalize() {
        sphere_emb = W("sphere_embedding.weight", 100 * C);
        csd = W("csd", C);
        normw = W("norm.affine_weight", 3 * C); normb = W("norm.affine_bias", C);
        h0 = W("head.0.weight", H * C); h0_t = W("head.0.weight_t", C * H); h0_b = W("head.0.bias", H);
        h2 = W("head.2.weight", H * H); h2_t = W("head.2.weight_t", H * H); h2_b = W("head.2.bias", H);
        h4 = W("head.4.weight", H); h4_b = W("head.4.bias", 1);
        ed_rad = radial("edge_degree.rad", 3 * C);
        layers.clear();
        for (int l = 0; l < cfg.num_layers; ++l) {
            std::string p = "blocks." + std::to_string(l);
            LayerW w;
            w.n1w = W(p + ".norm_1.affine_weight", 3 * C); w.n1b = W(p + ".norm_1.affine_bias", C);
            w.n2w = W(p + ".norm_2.affine_weight", 3 * C); w.n2b = W(p + ".norm_2.affine_bias", C);
            w.rad = radial(p + ".edge.conv1.rad", RAD1);
            w.c1m0 = W(p + ".edge.conv1.fc_m0.weight", 640 * 768); w.c1m0_t = W(p + ".edge.conv1.fc_m0.weight_t", 640 * 768);
            w.c1m0_b = W(p + ".edge.conv1.fc_m0.bias", 640);
            // m > 0: complex block weights [[W_r, -W_i], [W_i, W_r]] (engine.prepare_engine_weights)
            w.c1m1 = W(p + ".edge.conv1.fc_m1.weight", 512 * 1024); w.c1m1_t = W(p + ".edge.conv1.fc_m1.weight_t", 512 * 1024);
            w.c1m2 = W(p + ".edge.conv1.fc_m2.weight", 256 * 512); w.c1m2_t = W(p + ".edge.conv1.fc_m2.weight_t", 256 * 512);
            w.c2m0 = W(p + ".edge.conv2.fc_m0.weight", 384 * 384); w.c2m0_t = W(p + ".edge.conv2.fc_m0.weight_t", 384 * 384);
            w.c2m0_b = W(p + ".edge.conv2.fc_m0.bias", 384);
            w.c2m1 = W(p + ".edge.conv2.fc_m1.weight", 512 * 512); w.c2m1_t = W(p + ".edge.conv2.fc_m1.weight_t", 512 * 512);
            w.c2m2 = W(p + ".edge.conv2.fc_m2.weight", 256 * 256); w.c2m2_t = W(p + ".edge.conv2.fc_m2.weight_t", 256 * 256);
            w.smlp = W(p + ".ffn.scalar_mlp.weight", 2 * H * C); w.smlp_t = W(p + ".ffn.scalar_mlp.weight_t", 2 * H * C);
            w.smlp_b = W(p + ".ffn.scalar_mlp.bias", 2 * H);
            w.so3_1 = W(p + ".ffn.so3_1.weight", 3 * H * C); w.so3_1_t = W(p + ".ffn.so3_1.weight_t", 3 * H * C);
            w.so3_1_b = W(p + ".ffn.so3_1.bias", H);
            w.so3_2 = W(p + ".ffn.so3_2.weight", 3 * C * H); w.so3_2_t = W(p + ".ffn.so3_2.weight_t", 3 * C * H);
            w.so3_2_b = W(p + ".ffn.so3_2.bias", C);
            layers.push_back(w);
        }
        finalized = true;
    }

    // ------------------------------------------------------------------ helpers
    // gemm_mode 0: fp32 SIMT; 1: tcgen05 bf16x3; 2 (auto): tensor cores for images of >= 100 atoms.  The
    // choice depends on the image size only, never on the batch, so an image evaluated alone and
    // inside a batch goes through the same arithmetic.
    static bool splitk_enabled() {
        static const bool on = [] { const char* e = getenv("UMAB_SIMT_SPLITK"); return !(e && atoi(e) == 0); }();
        return on;
    }
    bool use_tc() const { return cfg.gemm_mode == 1 || (cfg.gemm_mode == 2 && n_atoms >= 100); }
    // precision study (set by umab_set_option "simt_round_fwd" / "simt_round_bwd"; SIMT path only)
    int simt_round_fwd = 0, simt_round_bwd = 0;
    bool in_backward = false;
    void gemm(const GemmArgs& a_in, cudaStream_t st) {
        GemmArgs a = a_in;
        a.round_mode = in_backward ? simt_round_bwd : simt_round_fwd;
        timed(P_GEMM, 2.0 * a.M * (double)a.N * a.K * a.batch, st, [&] {
            if (a.A_hi) gemm_tc2(a, st, tc2_cache, 0);
            else if (use_tc() && gemm_tc_supported(a)) gemm_tc(a, st, tc_cache);
            else if (n_atoms < 100 && splitk_enabled()) gemm_simt_splitk(a, st);     // small molecules: launch-latency bound
            else gemm_simt(a, st);
        }, /* algorithmic bytes: A and C once (+C again when accumulating), W once */
           4.0 * a.batch * ((double)a.M * a.K + (double)a.M * a.N * (a.accumulate ? 2 : 1)) + 4.0 * a.N * (double)a.K);
    }
    // C = A W^T (+bias): value plane with the bias, tangent plane (Hessian path) without
    template <class S>
    void mm(GP<S> A, long long lda, const float* Wt, int N, int K, GP<S> Cm, long long ldc, long long M,
            const float* bias, int accumulate, cudaStream_t st) {
        for (int k = 0; k < planes<S>(); ++k) {
            GemmArgs g;
            g.A = plane(A, k); g.lda = lda; g.W = Wt; g.ldw = K; g.Cmat = plane(Cm, k); g.ldc = ldc;
            g.bias = k == 0 ? bias : nullptr;
            g.N = N; g.K = K; g.accumulate = accumulate;
            gemm_plane(g, k, M, ldc, st);
        }
    }
    // same with an A operand written by an elementwise kernel (fp32, or bf16 hi/lo planes for the TMA-fed GEMM)
    template <class S>
    void mm_ap(AP<S> A, const float* Wt, int N, int K, GP<S> Cm, long long ldc, long long M, const float* bias,
               cudaStream_t st) {
        if (M <= 0) return;
        for (int k = 0; k < planes<S>(); ++k) {
            const AP<float> a = plane(A, k);
            GemmArgs g;
            g.A = a.p; g.A_hi = a.hi; g.A_lo = a.lo; g.lda = K; g.W = Wt; g.ldw = K; g.Cmat = plane(Cm, k); g.ldc = ldc;
            g.bias = k == 0 ? bias : nullptr;
            g.N = N; g.K = K; g.accumulate = 0;
            gemm_plane(g, k, M, ldc, st, /* every consumer of an mm_ap output is an ImgShare kernel */ true);
        }
    }
    // conv-1 m != 0 GEMM with the gate applied in the epilogue (float, CTA-pair kernel): C = A W^T (fp32, kept) and
    // C * sigmoid(gate) as the bf16 planes `Bout` of the conv-2 GEMM
    void mm_gated(AP<float> A, const float* Wt, int N, int K, GP<float> Cm, long long M, int gate_mode, AP<float> Bout,
                  cudaStream_t st) {
        GemmArgs g;
        g.A = A.p; g.A_hi = A.hi; g.A_lo = A.lo; g.lda = K; g.W = Wt; g.ldw = K; g.Cmat = Cm.p; g.ldc = N;
        g.M = (int)M; g.N = N; g.K = K;
        g.gate = wSG.f(); g.gate_ld = 256; g.gate_mode = gate_mode; g.out_hi = Bout.hi; g.out_lo = Bout.lo;
        gemm(g, st);
    }
    // per-coefficient SO(3) linear: rows (n, i) use weight l(i);  A [N,9,K] -> C [N,9,Nout]
    template <class S>
    void so3_mm(GP<S> A, int K, const float* Wt, int Nout, GP<S> Cm, const float* bias, int accumulate,
                cudaStream_t st) {
        for (int k = 0; k < planes<S>(); ++k) {
            GemmArgs g;
            g.A = plane(A, k); g.lda = 9LL * K; g.strideA = K;
            g.W = Wt; g.ldw = K; g.strideW = (long long)Nout * K;
            g.Cmat = plane(Cm, k); g.ldc = 9LL * Nout; g.strideC = Nout;
            g.bias = k == 0 ? bias : nullptr; g.bias_first_batch_only = 1;
            g.N = Nout; g.K = K; g.batch = 9; g.accumulate = accumulate;
            const int lsel[9] = {0, 1, 1, 1, 2, 2, 2, 2, 2};
            for (int i = 0; i < 9; ++i) g.wsel[i] = lsel[i];
            gemm_plane(g, k, n_nodes, 9LL * Nout, st);
        }
    }
    template <class S> void zero(const TBuf& b, size_t bytes, cudaStream_t st) {
        UMAB_CUDA(cudaMemsetAsync(b.v.p, 0, bytes, st));
        if (std::is_same<S, D1>::value) UMAB_CUDA(cudaMemsetAsync(b.d.p, 0, bytes, st));
    }
    template <class S> void copy(const TBuf& dst, const TBuf& srcb, size_t bytes, cudaStream_t st) {
        UMAB_CUDA(cudaMemcpyAsync(dst.v.p, srcb.v.p, bytes, cudaMemcpyDeviceToDevice, st));
        if (std::is_same<S, D1>::value) UMAB_CUDA(cudaMemcpyAsync(dst.d.p, srcb.d.p, bytes, cudaMemcpyDeviceToDevice, st));
    }
    void save_dbg(const std::string& name, const void* p, size_t numel, cudaStream_t st) {
        if (!cfg.debug) return;
        auto& slot = dbg[name];
        slot.first.ensure(numel * 4);
        slot.second = numel;
        UMAB_CUDA(cudaMemcpyAsync(slot.first.p, p, numel * 4, cudaMemcpyDeviceToDevice, st));
    }

    // debug: a per-chunk tensor copied into its rows [row0, row0 + rows) of a whole-batch debug tensor
    void save_dbg_rows(const std::string& name, const void* p, size_t rows, size_t width, size_t row0, size_t total_rows,
                       cudaStream_t st) {
        if (!cfg.debug) return;
        auto& slot = dbg[name];
        slot.first.ensure(total_rows * width * 4);
        slot.second = total_rows * width;
        UMAB_CUDA(cudaMemcpyAsync(slot.first.f() + row0 * width, p, rows * width * 4, cudaMemcpyDeviceToDevice, st));
    }

    // ------------------------------------------------------------------ graph
    // fcap > 0: sync-free build on edge arrays of that capacity (see the members above); 0: read row_ptr back
    void build_graph(const float* pos, int nimg, cudaStream_t st, long long fcap = 0) {
        if (n_atoms <= 0) throw CudaError("umab_set_system has not been called");
        if (nimg <= 0) throw CudaError("n_images must be positive");
        n_img = nimg;
        n_nodes = nimg * n_atoms;
        if ((long long)nimg * n_atoms > 0x7fffffffLL / 16) throw CudaError("batch too large: split it on the host");
        if (zt_img != nimg) {
            zt.ensure(sizeof(int) * n_nodes);
            launch_tile_int(z1.i(), n_atoms, nimg, zt.i(), st);
            zt_img = nimg;
        }
        deg.ensure(sizeof(int) * n_nodes); thr.ensure(sizeof(float) * n_nodes);
        row_ptr.ensure(sizeof(int) * (n_nodes + 1));
        const bool cells = neighbor_mode == 2 || (neighbor_mode == 0 && n_atoms >= 128);
        int cap_cells = 64;
        while (cap_cells < n_atoms / 2 && cap_cells < 65536) cap_cells *= 2;
        if (cells) {
            cgrid.ensure(cell_grid_bytes() * nimg); ccount.ensure(sizeof(int) * (size_t)nimg * cap_cells);
            cstart.ensure(sizeof(int) * (size_t)nimg * (cap_cells + 1));
            catoms.ensure(sizeof(int) * n_nodes); acell.ensure(sizeof(int) * n_nodes);
            launch_cell_list(pos, nimg, n_atoms, cfg.cutoff, cap_cells, cgrid.p, ccount.i(), acell.i(), cstart.i(), catoms.i(), st);
            launch_neighbor_cell_count(pos, nimg, n_atoms, cfg.cutoff, cfg.max_neighbors, cap_cells, cgrid.p, cstart.i(),
                                       catoms.i(), deg.i(), thr.f(), st);
        } else {
            launch_neighbor_count(pos, nimg, n_atoms, cfg.cutoff, cfg.max_neighbors, deg.i(), thr.f(), st);
        }
        launch_scan(deg.i(), row_ptr.i(), n_nodes, st);
        fast_graph = fcap > 0;
        const int* n_edges_dev = nullptr;
        if (fast_graph) {
            ++fast_calls;
            n_edges = fcap;
            n_edges_dev = row_ptr.i() + n_nodes;
            status_dev.ensure(2 * sizeof(int));
            launch_edge_status(row_ptr.i(), n_nodes, (int)fcap, status_dev.i(), st);
        } else {
            size_t need = sizeof(int) * (n_nodes + 1);
            if (need > h_pinned_cap) {
                if (h_pinned) cudaFreeHost(h_pinned);
                UMAB_CUDA(cudaMallocHost(&h_pinned, need + need / 4));
                h_pinned_cap = need + need / 4;
                ++g_alloc_gen;
            }
            UMAB_CUDA(cudaMemcpyAsync(h_pinned, row_ptr.p, need, cudaMemcpyDeviceToHost, st));
            UMAB_CUDA(cudaStreamSynchronize(st));
            n_edges = h_pinned[n_nodes];
            n_edges_actual = n_edges;
            hist_edges_per_image = std::max<long long>(hist_edges_per_image, (n_edges + nimg - 1) / nimg);
        }
        if (n_edges > 0x7fffffffLL / 40) throw CudaError("too many edges in one batch: split it on the host");
        size_t ne = (size_t)std::max<long long>(n_edges, 1);
        src.ensure(sizeof(int) * ne); tgt.ensure(sizeof(int) * ne);
        if (fast_graph) {
            // padding rows [actual, capacity) must hold valid node indices; source 1 -> target 0 is a regular pair of
            // distinct atoms (a self pair would have zero length: NaN Wigner blocks in the padding rows)
            launch_fill_int(src.i(), (int)ne, 1, st);
            UMAB_CUDA(cudaMemsetAsync(tgt.p, 0, sizeof(int) * ne, st));
        }
        if (cells)
            launch_neighbor_cell_fill(pos, nimg, n_atoms, cfg.cutoff, cfg.max_neighbors, cap_cells, cgrid.p, cstart.i(),
                                      catoms.i(), thr.f(), row_ptr.i(), src.i(), tgt.i(), (int)ne, st);
        else
            launch_neighbor_fill(pos, nimg, n_atoms, cfg.cutoff, cfg.max_neighbors, thr.f(), row_ptr.i(), src.i(), tgt.i(),
                                 (int)ne, st);
        odeg.ensure(sizeof(int) * n_nodes); sptr.ensure(sizeof(int) * (n_nodes + 1)); cursor.ensure(sizeof(int) * n_nodes);
        stmp.ensure(sizeof(int) * ne); sedge.ensure(sizeof(int) * ne);
        launch_source_csr(src.i(), (int)n_edges, n_edges_dev, n_nodes, odeg.i(), sptr.i(), cursor.i(), stmp.i(), sedge.i(), st);
    }

    // edges one chunk of the workspace can hold (worst case: the backward recomputes)
    template <class S> long long chunk_capacity(bool extra) const {
        const long long budget = cfg.workspace_bytes > 0 ? cfg.workspace_bytes : (28LL << 30);   // one 10k-atom image (~0.8 M edges) stays a closed chunk
        const size_t per_edge = EDGE_WS_FLOATS + (extra ? EDGE_WS_EXTRA_RECOMPUTE : 0);
        return std::max<long long>(budget / (long long)(per_edge * 4 * planes<S>()), 1024);
    }
    // capacity of the sync-free build for `nimg` images, or 0 when this call must read the edge count back: debug /
    // profiling runs, no history yet for images of >= 128 atoms, or more edges than ONE chunk of the workspace holds
    bool allow_fast = true;                            // false inside a call that runs as several sub-batches
    template <class S> long long fast_capacity(int nimg) const {
        static const bool debug_fast = getenv("UMAB_DEBUG_FAST") != nullptr;      // keep the debug tensors on the sync-free path
        if (!nosync_enabled() || !allow_fast || (cfg.debug && !debug_fast) || prof.on || n_atoms < 2) return 0;
        const long long complete = (long long)n_atoms * (n_atoms - 1);
        long long per;
        if (n_atoms < 128) per = complete;
        else if (hist_edges_per_image > 0) {
            // + 3 % head-room, rounded up to a 1/16-octave grid so that the capacity (= the key of the captured CUDA
            // graphs and the shape of every launch) stays put while the geometry drifts
            per = hist_edges_per_image + hist_edges_per_image / 32 + 64;
            long long q = 1;
            while ((q << 5) <= per) q <<= 1;              // q = 2^(floor(log2 per) - 4)
            per = std::min(complete, (per + q - 1) / q * q);
        } else return 0;
        const long long tot = per * nimg;
        return tot <= chunk_capacity<S>(true) ? tot : 0;
    }
    // collect {actual edge count, overflow} of a sync-free evaluation; true = the capacity was exceeded (results invalid:
    // repeat the call -- the history now knows the larger count)
    bool resolve_status() {
        if (!status_pending) return false;
        status_pending = false;
        UMAB_CUDA(cudaEventSynchronize(status_ev));
        n_edges_actual = h_status[0];
        hist_edges_per_image = std::max<long long>(hist_edges_per_image, (n_edges_actual + status_nimg - 1) / std::max(status_nimg, 1));
        return h_status[1] != 0;
    }
    void enqueue_status(cudaStream_t st, bool record_event) {
        if (!h_status) { UMAB_CUDA(cudaMallocHost(&h_status, 2 * sizeof(int))); ++g_alloc_gen; }
        if (!status_ev) UMAB_CUDA(cudaEventCreateWithFlags(&status_ev, cudaEventDisableTiming));
        UMAB_CUDA(cudaMemcpyAsync(h_status, status_dev.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        if (record_event) UMAB_CUDA(cudaEventRecord(status_ev, st));
        status_nimg = n_img;
    }

    template <class S> void plan_chunks() {
        const bool extra = want_adjoint && !store_mode;
        long long cap = chunk_capacity<S>(extra);
        chunks.clear();
        int node0 = 0;
        long long biggest = 0;
        if (fast_graph) {          // one closed chunk over the whole capacity (fast_capacity checked that it fits)
            chunks.push_back(Chunk{0, n_nodes, 0, (int)n_edges});
            chunks_closed = closed_chunks_enabled();
            if (!chunks_closed) throw CudaError("UMAB_CLOSED_CHUNKS=0 needs UMAB_NOSYNC=0");
            biggest = n_edges;
            node0 = n_nodes;
        } else {
        // closed chunks = whole images: every out-edge of a node of the chunk is an edge of the chunk, so the adjoint
        // can reduce the source halves by source node inside the chunk (no per-edge G buffer, no source_reduce pass)
        long long max_img = 0;
        for (int b = 0; b < n_img; ++b)
            max_img = std::max<long long>(max_img, (long long)h_pinned[(b + 1) * n_atoms] - h_pinned[b * n_atoms]);
        chunks_closed = closed_chunks_enabled() && max_img <= cap;
        if (chunks_closed) {
            int b0 = 0;
            while (b0 < n_img) {
                const long long e0 = h_pinned[b0 * n_atoms];
                int b1 = b0 + 1;
                while (b1 < n_img && (long long)h_pinned[(b1 + 1) * n_atoms] - e0 <= cap) ++b1;
                Chunk c{b0 * n_atoms, (b1 - b0) * n_atoms, e0, (int)(h_pinned[b1 * n_atoms] - e0)};
                chunks.push_back(c);
                biggest = std::max<long long>(biggest, c.n_e);
                b0 = b1;
            }
            node0 = n_nodes;
        }
        }
        while (node0 < n_nodes) {
            long long e0 = h_pinned[node0];
            int n1 = node0;
            while (n1 < n_nodes && (long long)h_pinned[n1 + 1] - e0 <= cap) ++n1;
            if (n1 == node0) throw CudaError("edge workspace too small for a single atom's neighbour list");
            Chunk c{node0, n1 - node0, e0, (int)(h_pinned[n1] - e0)};
            chunks.push_back(c);
            biggest = std::max<long long>(biggest, c.n_e);
            node0 = n1;
        }
        chunk_cap = std::max<long long>(biggest, 1);
        size_t f = sizeof(float) * (size_t)chunk_cap;
        // wY also holds g_rad [n_e, 1536] in store mode (bufs_for: b.grad), which is wider than Y [n_e, 1408]
        wA.ensure<S>(f * 2304); wY.ensure<S>(f * std::max(YW, 1536)); wB.ensure<S>(f * 1152); wZ.ensure<S>(f * ZW);
        wRAD.ensure<S>(f * 1536);
        wU1.ensure<S>(f * 128); wH1.ensure<S>(f * 128); wU2.ensure<S>(f * 128); wH2.ensure<S>(f * 128);
        if (extra) { wGY.ensure<S>(f * std::max(YW, 1536)); wGZ.ensure<S>(f * ZW); }
        if (std::is_same<S, float>::value) wSG.ensure(f * 256);
    }

    // ------------------------------------------------------------------ edge stages
    // Buffers of one chunk.  A-operand activations (AP) are written by the elementwise kernels in the format of
    // the GEMM that consumes them; the adjoint operands g_Y / g_Z / g_rad get buffers of their own (the format
    // change rules out the in-place updates of an all-fp32 pipeline): the Y / Z workspace when Y / Z live in the
    // per-layer stores, else wGY / wGZ.
    template <class S> struct EB {
        AP<S> a0, a1, a2, b0, b1, b2;            // conv-1 / conv-2 inputs
        GP<S> ga0, ga1, ga2, gb0, gb1, gb2;      // their gradients (adjoint GEMM outputs, fp32, same memory)
        GP<S> y0, y1, y2, z0, z1, z2;            // conv outputs
        AP<S> gy0, gy1, gy2, gz0, gz1, gz2;      // gradients of the conv outputs (adjoint GEMM inputs)
        AP<S> grad, ged;                         // g_rad [n_e,1536] / edge-degree g_rad [n_e,384]
        GP<S> rad, u1, u2, h1f, h2f;             // radial MLP: fp32 tensors (h1f / h2f = fp32 views of h1 / h2)
        AP<S> h1, h2, gu;                        // radial MLP: GEMM operands (gu = dL/du_2 in the h1 buffer)
    };
    template <class S> EB<S> bufs_for(int layer, const Chunk& c) {
        EB<S> b;
        const long long cc = chunk_cap;
        const bool sp = use_tc();
        b.a0 = ap<S>(wA, 0, cc * 768, sp); b.a1 = ap<S>(wA, cc * 768, cc * 1024, sp);
        b.a2 = ap<S>(wA, cc * (768 + 1024), cc * 512, sp);
        b.ga0 = gp<S>(wA); b.ga1 = gp<S>(wA, cc * 768); b.ga2 = gp<S>(wA, cc * (768 + 1024));
        b.b0 = ap<S>(wB, 0, cc * 384, sp); b.b1 = ap<S>(wB, cc * 384, cc * 512, sp);
        b.b2 = ap<S>(wB, cc * (384 + 512), cc * 256, sp);
        b.gb0 = gp<S>(wB); b.gb1 = gp<S>(wB, cc * 384); b.gb2 = gp<S>(wB, cc * (384 + 512));
        const bool stored = store_mode && layer >= 0;
        if (stored) {
            const long long E = n_edges;
            b.y0 = gp<S>(ystore[layer], c.e0 * 640); b.y1 = gp<S>(ystore[layer], E * 640 + c.e0 * 512);
            b.y2 = gp<S>(ystore[layer], E * 1152 + c.e0 * 256);
            b.z0 = gp<S>(zstore[layer], c.e0 * 384); b.z1 = gp<S>(zstore[layer], E * 384 + c.e0 * 512);
            b.z2 = gp<S>(zstore[layer], E * 896 + c.e0 * 256);
        } else {
            b.y0 = gp<S>(wY); b.y1 = gp<S>(wY, cc * 640); b.y2 = gp<S>(wY, cc * (640 + 512));
            b.z0 = gp<S>(wZ); b.z1 = gp<S>(wZ, cc * 384); b.z2 = gp<S>(wZ, cc * (384 + 512));
        }
        const TBuf& gyb = store_mode ? wY : wGY;
        const TBuf& gzb = store_mode ? wZ : wGZ;
        b.gy0 = ap<S>(gyb, 0, cc * 640, sp); b.gy1 = ap<S>(gyb, cc * 640, cc * 512, sp);
        b.gy2 = ap<S>(gyb, cc * (640 + 512), cc * 256, sp);
        b.gz0 = ap<S>(gzb, 0, cc * 384, sp); b.gz1 = ap<S>(gzb, cc * 384, cc * 512, sp);
        b.gz2 = ap<S>(gzb, cc * (384 + 512), cc * 256, sp);
        b.grad = ap<S>(gyb, 0, cc * 1536, sp);       // written after the conv-1 adjoint GEMMs have consumed g_Y
        b.ged = ap<S>(gyb, 0, cc * 384, sp);
        if (stored && store_radial_enabled()) {
            b.rad = gp<S>(rstore[layer], c.e0 * 1536); b.u1 = gp<S>(u1store[layer], c.e0 * 128);
            b.u2 = gp<S>(u2store[layer], c.e0 * 128);
        } else {
            b.rad = gp<S>(wRAD); b.u1 = gp<S>(wU1); b.u2 = gp<S>(wU2);
        }
        b.h1f = gp<S>(wH1); b.h2f = gp<S>(wH2);
        b.h1 = ap<S>(wH1, 0, cc * 128, sp); b.h2 = ap<S>(wH2, 0, cc * 128, sp); b.gu = ap<S>(wH1, 0, cc * 128, sp);
        // Hessian columns of one base geometry: only the first image of the chunk feeds the value-plane GEMMs, so the
        // value planes of the bf16-plane A operands (consumed by nothing but those GEMMs) are not stored for the others
        static const bool skip_dead = [] { const char* e = getenv("UMAB_SKIP_DEAD_VALUES"); return !(e && atoi(e) == 0); }();   // A/B switch
        if (skip_dead && sp && chunk_rep_rows() > 0 && c.n_e > e_img && c.n_e % e_img == 0) {
            const long long r = e_img;
            value_rows(b.a0, r * 768); value_rows(b.a1, r * 1024); value_rows(b.a2, r * 512);
            value_rows(b.b0, r * 384); value_rows(b.b1, r * 512); value_rows(b.b2, r * 256);
            value_rows(b.gy0, r * 640); value_rows(b.gy1, r * 512); value_rows(b.gy2, r * 256);
            value_rows(b.gz0, r * 384); value_rows(b.gz1, r * 512); value_rows(b.gz2, r * 256);
            value_rows(b.grad, r * 1536); value_rows(b.ged, r * 384);
            value_rows(b.h1, r * 128); value_rows(b.h2, r * 128); value_rows(b.gu, r * 128);
        }
        return b;
    }

    template <class S> void radial_fwd(const RadialW& r, const Chunk& c, const EB<S>& b, cudaStream_t st) {
        const long long e0 = c.e0;
        mm<S>(gp<S>(gauss, e0 * NB), NB, r.w1g, 128, NB, b.u1, 128, c.n_e, nullptr, 0, st);
        const double lnb = planes<S>() * c.n_e * 128.0 * 4.0;     // one [n_e,128] fp32 tensor
        timed(P_LN_SILU, 3 * lnb, st, [&] {
            launch_ln_silu_fwd_t<S>(b.u1, b.h1, r.ln1w, r.ln1b, r.b1, r.t_src, r.t_tgt, zt.i(), src.i() + e0, tgt.i() + e0, c.n_e, st); });
        mm_ap<S>(b.h1, r.w2, 128, 128, b.u2, 128, c.n_e, r.b2, st);
        timed(P_LN_SILU, 2 * lnb, st, [&] {
            launch_ln_silu_fwd_t<S>(b.u2, b.h2, r.ln2w, r.ln2b, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, c.n_e, st,
                                    share_e(c)); });
        mm_ap<S>(b.h2, r.w3, r.n_out, 128, b.rad, r.n_out, c.n_e, r.b3, st);
    }
    // g_rad: [n_e, n_out] A operand; accumulates into g_gauss.  u1 / u2 are the forward pre-activations.
    template <class S> void radial_bwd(const RadialW& r, const Chunk& c, const EB<S>& b, AP<S> g_rad, cudaStream_t st) {
        mm_ap<S>(g_rad, r.w3_t, 128, r.n_out, b.h2f, 128, c.n_e, nullptr, st);                  // dL/dh_2
        const double lnb = planes<S>() * c.n_e * 128.0 * 4.0;
        timed(P_LN_SILU, 3 * lnb, st, [&] {
            launch_ln_silu_bwd_t<S>(b.u2, b.h2f, b.gu, r.ln2w, r.ln2b, c.n_e, st, share_e(c), true); });           // dL/du_2
        mm_ap<S>(b.gu, r.w2_t, 128, 128, b.h2f, 128, c.n_e, nullptr, st);                        // dL/dh_1
        // dL/du_1 stays fp32: the last GEMM (N = 64, accumulating) runs on the in-kernel-split path
        timed(P_LN_SILU, 3 * lnb, st, [&] {
            launch_ln_silu_bwd_t<S>(b.u1, b.h2f, ap<S>(wH1, 0, 0, false), r.ln1w, r.ln1b, c.n_e, st, share_e(c),
                                    /* u_1 (per-element tables added in place) is per image */ false); });
        mm<S>(b.h1f, 128, r.w1g_t, NB, 128, gp<S>(g_gauss, c.e0 * NB), NB, c.n_e, nullptr, 1, st);
    }
    // conv-1 radial, gather/rotate, conv-1, gate, conv-2 for one chunk (everything up to Z)
    template <class S>
    void edge_fwd_chunk(const LayerW& w, GP<S> n1, const Chunk& c, int layer, bool dbg_on, cudaStream_t st) {
        const EB<S> b = bufs_for<S>(layer, c);
        const double P = planes<S>();
        radial_fwd<S>(w.rad, c, b, st);
        timed(P_GATHER, P * (c.n_e * 15512.0 + c.n_nodes * 4608.0), st, [&] {
            launch_gather_rotate_scale_t<S>(n1, src.i(), tgt.i(), gp<S>(wig), b.rad, c.e0, c.n_e, b.a0, b.a1, b.a2, st,
                                            share_e(c)); });
        mm_ap<S>(b.a0, w.c1m0, 640, 768, b.y0, 640, c.n_e, w.c1m0_b, st);
        if constexpr (std::is_same<S, float>::value) {
            if (opt_fuse_gate && use_tc() && gemm_tc2_gate_epilogue_available() && c.n_e > 0) {
                // gate fused into the conv-1 GEMM epilogues: the m = 0 part (needs its own gate columns, other N tiles
                // of the same GEMM) stays a small kernel that also leaves sigmoid(gates); the m = +-1 / +-2 GEMMs
                // write Y (kept for the backward) AND the gated bf16 planes B1 / B2 -- Y1 / Y2 are never re-read and
                // B1 / B2 never pass through a separate kernel (same bits as the un-fused path)
                timed(P_COMBINE, c.n_e * (640 + 384 + 256) * 4.0, st, [&] {
                    launch_gate_b0(b.y0, c.n_e, b.b0, wSG.f(), st); });
                mm_gated(b.a1, w.c1m1, 512, 1024, b.y1, c.n_e, 1, b.b1, st);
                mm_gated(b.a2, w.c1m2, 256, 512, b.y2, c.n_e, 2, b.b2, st);
                goto conv2;
            }
        }
        mm_ap<S>(b.a1, w.c1m1, 512, 1024, b.y1, 512, c.n_e, nullptr, st);
        mm_ap<S>(b.a2, w.c1m2, 256, 512, b.y2, 256, c.n_e, nullptr, st);
        timed(P_COMBINE, P * c.n_e * (YW + 1152) * 4.0, st, [&] {
            launch_combine_gate_fwd_t<S>(b.y0, b.y1, b.y2, c.n_e, b.b0, b.b1, b.b2, st, share_e(c)); });
    conv2:
        mm_ap<S>(b.b0, w.c2m0, 384, 384, b.z0, 384, c.n_e, w.c2m0_b, st);
        mm_ap<S>(b.b1, w.c2m1, 512, 512, b.z1, 512, c.n_e, nullptr, st);
        mm_ap<S>(b.b2, w.c2m2, 256, 256, b.z2, 256, c.n_e, nullptr, st);
        if (dbg_on && chunks.size() == 1) {
            std::string p = "l" + std::to_string(layer) + ".";
            save_dbg(p + "rad", plane(b.rad, 0), (size_t)c.n_e * 1536, st);
            save_dbg(p + "y0", plane(b.y0, 0), (size_t)c.n_e * 640, st);
            save_dbg(p + "y1", plane(b.y1, 0), (size_t)c.n_e * 512, st);
            save_dbg(p + "y2", plane(b.y2, 0), (size_t)c.n_e * 256, st);
            save_dbg(p + "z0", plane(b.z0, 0), (size_t)c.n_e * 384, st);
        }
    }
    template <class S>
    void edge_bwd_chunk(const LayerW& w, GP<S> n1, const Chunk& c, int layer, GP<S> g_out, GP<S> g_n1, cudaStream_t st) {
        const EB<S> b = bufs_for<S>(layer, c);
        const double P = planes<S>();
        if (!store_mode) edge_fwd_chunk<S>(w, n1, c, layer, false, st);      // recompute everything up to Z
        else if (!store_radial_enabled()) radial_fwd<S>(w.rad, c, b, st);
        timed(P_ROTBACK_BWD, P * (c.n_e * (2 * ZW * 4.0 + 2 * 148.0 + 8.0) + c.n_nodes * 4608.0), st, [&] {
            launch_rotate_back_bwd_t<S>(0, b.z0, b.z1, b.z2, tgt.i(), gp<S>(wig), gp<S>(env), 1.0f, c.e0, c.n_e, g_out,
                                        b.gz0, b.gz1, b.gz2, gp<S>(g_env), gp<S>(g_wig), st, share_e(c)); });
        if (cfg.debug && !use_tc()) {
            const std::string p = "bwd.l" + std::to_string(layer) + ".";
            save_dbg_rows(p + "gz0", plane(b.gb0, 0) /* same memory as gz0 in fp32 mode */, c.n_e, 384, c.e0, n_edges, st);
        }
        mm_ap<S>(b.gz0, w.c2m0_t, 384, 384, b.gb0, 384, c.n_e, nullptr, st);
        mm_ap<S>(b.gz1, w.c2m1_t, 512, 512, b.gb1, 512, c.n_e, nullptr, st);
        mm_ap<S>(b.gz2, w.c2m2_t, 256, 256, b.gb2, 256, c.n_e, nullptr, st);
        timed(P_COMBINE_BWD, P * c.n_e * (YW * 4.0 * 2 + 4608.0), st, [&] {
            launch_combine_gate_bwd_t<S>(b.y0, b.y1, b.y2, c.n_e, b.gb0, b.gb1, b.gb2, b.gy0, b.gy1, b.gy2, st, share_e(c)); });
        if (cfg.debug && !use_tc()) {
            const std::string p = "bwd.l" + std::to_string(layer) + ".";
            save_dbg_rows(p + "gb0", plane(b.gb0, 0), c.n_e, 384, c.e0, n_edges, st);
            save_dbg_rows(p + "gy0", plane(b.gy0, 0).p, c.n_e, 640, c.e0, n_edges, st);
        }
        mm_ap<S>(b.gy0, w.c1m0_t, 768, 640, b.ga0, 768, c.n_e, nullptr, st);
        mm_ap<S>(b.gy1, w.c1m1_t, 1024, 512, b.ga1, 1024, c.n_e, nullptr, st);
        mm_ap<S>(b.gy2, w.c1m2_t, 512, 256, b.ga2, 512, c.n_e, nullptr, st);
        if (cfg.debug && !use_tc()) {
            const std::string p = "bwd.l" + std::to_string(layer) + ".";
            save_dbg_rows(p + "ga0", plane(b.ga0, 0), c.n_e, 768, c.e0, n_edges, st);
            save_dbg_rows(p + "rad", plane(b.rad, 0), c.n_e, 1536, c.e0, n_edges, st);
        }
        if (chunks_closed)
            timed(P_GATHER_BWD, P * (c.n_e * (9216.0 + 2 * 6144.0 + 6 * 144.0 + 8.0) + c.n_nodes * 5 * 4608.0), st, [&] {
                launch_gather_rotate_bwd_closed_t<S>(n1, row_ptr.i(), sptr.i(), sedge.i(), gp<S>(wig), b.rad, c.e0, c.node0,
                                                     c.n_nodes, b.ga0, b.ga1, b.ga2, b.grad, g_n1, gp<S>(g_wig), st,
                                                     share_n(c), (int)e_img); });
        else
        timed(P_GATHER_BWD, P * (c.n_e * (9216.0 + 2 * 6144.0 + 4608.0 + 3 * 144.0 + 4.0) + c.n_nodes * 2 * 4608.0), st, [&] {
            launch_gather_rotate_bwd_t<S>(n1, row_ptr.i(), src.i(), gp<S>(wig), b.rad, c.e0, c.node0, c.n_nodes, b.ga0, b.ga1,
                                          b.ga2, b.grad, gp<S>(Gbuf), g_n1, gp<S>(g_wig), st); });
        radial_bwd<S>(w.rad, c, b, b.grad, st);
    }

    // ------------------------------------------------------------------ full evaluation
    // S = float: energies + forces.  S = D1: positions carry a tangent (displacement direction);
    // forces.d then is d(forces)/d(eps) = -H.t, one analytic Hessian column per image.
    bool capturing = false;                            // inside cudaStreamBeginCapture / EndCapture
    cudaStream_t own_stream = nullptr;                 // host-buffer calls issued on the legacy default stream run here
    cudaEvent_t order_ev = nullptr;
    template <class S>
    void evaluate(GP<S> pos, int nimg, double* energy_dev, GP<S> forces, cudaStream_t st) {
        evaluate_impl<S>(pos, nimg, energy_dev, forces, st);
        if (fast_graph) {
            enqueue_status(st, !capturing);
            status_pending = !capturing;
        }
        call_images += nimg; call_subcalls += 1;
    }
    // a whole public call: sub-batches sized by images_per_call
    template <class S>
    void evaluate_batched(GP<S> pos, int nimg, double* energy_dev, GP<S> forces, cudaStream_t st) {
        call_images = 0; call_edges = 0; call_subcalls = 0;
        const long long stride = (long long)n_atoms * 3;
        int s0 = 0;
        allow_fast = images_per_call((bool)forces, planes<S>()) >= nimg;
        struct Restore { bool& f; ~Restore() { f = true; } } restore{allow_fast};
        while (s0 < nimg) {
            const int step = std::max(1, images_per_call((bool)forces, planes<S>()));
            const int nb = std::min(step, nimg - s0);
            GP<S> p = pos, f = forces;
            offset_gp(p, s0 * stride);
            if (forces) offset_gp(f, s0 * stride);
            evaluate<S>(p, nb, energy_dev ? energy_dev + s0 : nullptr, f, st);
            if (!fast_graph) call_edges += n_edges;
            s0 += nb;
        }
    }
    static void offset_gp(GP<float>& g, long long off) { g.p += off; }
    static void offset_gp(GP<D1>& g, long long off) { g.v += off; g.d += off; }
    template <class S>
    void evaluate_impl(GP<S> pos, int nimg, double* energy_dev, GP<S> forces, cudaStream_t st) {
        if (!finalized) throw CudaError("umab_finalize_weights has not been called");
        if (resolve_status()) { /* an unchecked overflow of an earlier device-pointer call: the history has been updated */ }
        const bool want_shared = std::is_same<S, D1>::value && opt_shared_base && nimg > 1;
        const long long fcap = want_shared ? 0 : fast_capacity<S>(nimg);      // the de-duplication needs the edge counts on the host
        timed(P_GRAPH, (double)nimg * n_atoms * 12.0, st, [&] { build_graph(plane(pos, 0), nimg, st, fcap); });
        dedupe = false; rep_rows = 0; e_img = 0;
        if (want_shared) {
            same_flag.ensure(sizeof(int));
            launch_same_images(plane(pos, 0), (long long)n_atoms * 3, nimg, same_flag.i(), st);
            int differ = 1;
            UMAB_CUDA(cudaMemcpyAsync(&differ, same_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            UMAB_CUDA(cudaStreamSynchronize(st));
            e_img = h_pinned[n_atoms];
            bool same = differ == 0 && e_img > 0;
            for (int b = 1; b < nimg && same; ++b)
                same = (long long)h_pinned[(b + 1) * n_atoms] - h_pinned[b * n_atoms] == e_img;
            dedupe = same;
            rep_rows = dedupe ? n_atoms : 0;          // node-level GEMMs; the chunk loops set the edge-level value
        }
        struct DedupeOff { umab_engine* e; ~DedupeOff() { e->dedupe = false; e->rep_rows = 0; } } dedupe_off{this};
        const int L = cfg.num_layers;
        const size_t ne = (size_t)std::max<long long>(n_edges, 1);
        const size_t nf = (size_t)n_nodes * 9 * C * sizeof(float);
        const bool want_f = (bool)forces;
        want_adjoint = want_f;
        {
            const double need = (double)n_edges * (store_radial_enabled() ? STORE_FLOATS : YW + ZW) * 4.0 * L * planes<S>();
            double budget = (double)cfg.store_bytes;
            if (cfg.store_bytes == 0) budget = 0.45 * (double)device_memory();
            store_mode = want_f && cfg.store_bytes >= 0 && need <= budget;
            if (store_mode) {
                ystore.resize(L); zstore.resize(L); rstore.resize(L); u1store.resize(L); u2store.resize(L);
                for (auto& b : ystore) b.ensure<S>(ne * YW * 4);
                for (auto& b : zstore) b.ensure<S>(ne * ZW * 4);
                if (store_radial_enabled()) {
                    for (auto& b : rstore) b.ensure<S>(ne * 1536 * 4);
                    for (auto& b : u1store) b.ensure<S>(ne * 128 * 4);
                    for (auto& b : u2store) b.ensure<S>(ne * 128 * 4);
                }
            }
        }
        plan_chunks<S>();
        vec.ensure<S>(ne * 12); dist.ensure<S>(ne * 4); env.ensure<S>(ne * 4); wig.ensure<S>(ne * WIG * 4);
        gauss.ensure<S>(ne * NB * 4);
        timed(P_GEOMETRY, planes<S>() * (double)n_edges * (24.0 + 4.0 * (5 + WIG + NB)), st, [&] {
            launch_geometry_fwd_t<S>(pos, src.i(), tgt.i(), (int)n_edges, cfg.cutoff, gp<S>(vec), gp<S>(dist), gp<S>(env),
                                     gp<S>(wig), gp<S>(gauss), st); });
        xs.resize(L + 1); x1s.resize(L); y1s.resize(L); gps.resize(L);
        for (auto& b : xs) b.ensure<S>(nf);
        for (auto& b : x1s) b.ensure<S>(nf);
        for (auto& b : y1s) b.ensure<S>(nf);
        for (auto& b : gps) b.ensure<S>((size_t)n_nodes * 2 * H * 4);
        nbuf.ensure<S>(nf); abuf.ensure<S>(nf);
        p1.ensure<S>((size_t)n_nodes * H * 4); s1.ensure<S>((size_t)n_nodes * H * 4); p2.ensure<S>((size_t)n_nodes * H * 4);
        node_e.ensure((size_t)n_nodes * 4);

        // ---- embedding (position independent: zero tangent) + edge-degree embedding
        launch_embed(sphere_emb, csd, zt.i(), n_nodes, xs[0].v.f(), st);
        if (std::is_same<S, D1>::value) UMAB_CUDA(cudaMemsetAsync(xs[0].d.p, 0, nf, st));
        for (const Chunk& c : chunks) {
            RepScope rs(this, chunk_rep_rows(), chunk_images(c) > 1);
            const EB<S> b = bufs_for<S>(-1, c);
            radial_fwd<S>(ed_rad, c, b, st);
            launch_rotate_back_reduce_t<S>(1, b.rad, gnull<S>(), gnull<S>(), row_ptr.i(), gp<S>(wig), gp<S>(env),
                                           1.0f / cfg.edge_degree_rescale, c.e0, c.node0, c.n_nodes, gp<S>(xs[0]),
                                           gp<S>(xs[0]), st, share_n(c), (int)e_img);
        }
        save_dbg("x0", xs[0].v.p, (size_t)n_nodes * 9 * C, st);
        save_dbg("gauss", gauss.v.p, (size_t)n_edges * NB, st);
        save_dbg("wig", wig.v.p, (size_t)n_edges * WIG, st);
        save_dbg("env", env.v.p, (size_t)n_edges, st);

        // ---- layers
        for (int l = 0; l < L; ++l) {
            const LayerW& w = layers[l];
            launch_rms_fwd_t<S>(gp<S>(xs[l]), w.n1w, w.n1b, csd, n_nodes, gp<S>(nbuf), st);
            save_dbg("l" + std::to_string(l) + ".n1", nbuf.v.p, (size_t)n_nodes * 9 * C, st);
            for (const Chunk& c : chunks) {
                RepScope rs(this, chunk_rep_rows(), chunk_images(c) > 1);
                edge_fwd_chunk<S>(w, gp<S>(nbuf), c, l, true, st);
                const EB<S> b = bufs_for<S>(l, c);
                timed(P_ROTBACK, planes<S>() * (c.n_e * (ZW * 4.0 + 148.0) + c.n_nodes * 9216.0), st, [&] {
                    launch_rotate_back_reduce_t<S>(0, b.z0, b.z1, b.z2, row_ptr.i(), gp<S>(wig), gp<S>(env), 1.0f, c.e0,
                                                   c.node0, c.n_nodes, gp<S>(xs[l]), gp<S>(x1s[l]), st, share_n(c), (int)e_img); });
            }
            save_dbg("l" + std::to_string(l) + ".x1", x1s[l].v.p, (size_t)n_nodes * 9 * C, st);
            launch_rms_fwd_t<S>(gp<S>(x1s[l]), w.n2w, w.n2b, nullptr, n_nodes, gp<S>(nbuf), st);
            mm<S>(gp<S>(nbuf), 9LL * C, w.smlp, 2 * H, C, gp<S>(gps[l]), 2 * H, n_nodes, w.smlp_b, 0, st);
            so3_mm<S>(gp<S>(nbuf), C, w.so3_1, H, gp<S>(y1s[l]), w.so3_1_b, 0, st);
            launch_ffn_gate_fwd_t<S>(gp<S>(y1s[l]), gp<S>(gps[l]), n_nodes, gp<S>(abuf), st);
            copy<S>(xs[l + 1], x1s[l], nf, st);
            so3_mm<S>(gp<S>(abuf), H, w.so3_2, C, gp<S>(xs[l + 1]), w.so3_2_b, 1, st);
            save_dbg("l" + std::to_string(l) + ".x", xs[l + 1].v.p, (size_t)n_nodes * 9 * C, st);
        }
        // ---- head
        launch_rms_fwd_t<S>(gp<S>(xs[L]), normw, normb, nullptr, n_nodes, gp<S>(nbuf), st);
        mm<S>(gp<S>(nbuf), 9LL * C, h0, H, C, gp<S>(p1), H, n_nodes, h0_b, 0, st);
        launch_eltwise_t<S>(0, gp<S>(p1), gnull<S>(), (long long)n_nodes * H, gp<S>(s1), st);
        mm<S>(gp<S>(s1), H, h2, H, H, gp<S>(p2), H, n_nodes, h2_b, 0, st);
        if (want_f) gp2.ensure<S>((size_t)n_nodes * H * 4);
        launch_head_final_t<S>(gp<S>(p2), h4, h4_b, n_nodes, node_e.f(), want_f ? gp<S>(gp2) : gnull<S>(), st);
        if (energy_dev) launch_energy_reduce(node_e.f(), n_img, n_atoms, energy_dev, st);
        save_dbg("node_e", node_e.p, (size_t)n_nodes, st);
        if (!want_f) return;

        // ================= backward: dE_total/dpos
        in_backward = true;
        struct Leave { bool& f; ~Leave() { f = false; } } leave{in_backward};
        gx.ensure<S>(nf); gx1.ensure<S>(nf); gn.ensure<S>(nf); ggp.ensure<S>((size_t)n_nodes * 2 * H * 4);
        gs1.ensure<S>((size_t)n_nodes * H * 4);
        if (!chunks_closed) Gbuf.ensure<S>(ne * 9 * C * 4);
        g_gauss.ensure<S>(ne * NB * 4); g_env.ensure<S>(ne * 4); g_wig.ensure<S>(ne * 4 * 4); g_vec.ensure<S>(ne * 12);   // g_wig: torque record [E, 4]
        zero<S>(g_gauss, ne * NB * 4, st);
        zero<S>(g_env, ne * 4, st);
        zero<S>(g_wig, ne * 4 * 4, st);

        mm<S>(gp<S>(gp2), H, h2_t, H, H, gp<S>(gs1), H, n_nodes, nullptr, 0, st);
        launch_eltwise_t<S>(1, gp<S>(gs1), gp<S>(p1), (long long)n_nodes * H, gp<S>(gs1), st);      // g_p1
        zero<S>(gn, nf, st);
        mm<S>(gp<S>(gs1), H, h0_t, C, H, gp<S>(gn), 9LL * C, n_nodes, nullptr, 0, st);             // g_xf (row 0 only)
        launch_rms_bwd_t<S>(gp<S>(xs[L]), normw, gp<S>(gn), gnull<S>(), n_nodes, gp<S>(gx), st);
        for (int l = L - 1; l >= 0; --l) {
            const LayerW& w = layers[l];
            // FFN adjoint (dL/dy2 = gx)
            so3_mm<S>(gp<S>(gx), C, w.so3_2_t, H, gp<S>(abuf), nullptr, 0, st);                     // g_a
            launch_ffn_gate_bwd_t<S>(gp<S>(y1s[l]), gp<S>(gps[l]), gp<S>(abuf), n_nodes, gp<S>(abuf), gp<S>(ggp), st);
            so3_mm<S>(gp<S>(abuf), H, w.so3_1_t, C, gp<S>(gn), nullptr, 0, st);                     // g_n2
            mm<S>(gp<S>(ggp), 2 * H, w.smlp_t, C, 2 * H, gp<S>(gn), 9LL * C, n_nodes, nullptr, 1, st);
            // the FFN adjoint needs n2 = norm_2(x1) only through its stored outputs (y1, gp): no recompute
            launch_rms_bwd_t<S>(gp<S>(x1s[l]), w.n2w, gp<S>(gn), gp<S>(gx), n_nodes, gp<S>(gx1), st);   // g_x1
            // Edgewise adjoint
            launch_rms_fwd_t<S>(gp<S>(xs[l]), w.n1w, w.n1b, csd, n_nodes, gp<S>(nbuf), st);         // recompute n1
            for (const Chunk& c : chunks) {
                RepScope rs(this, chunk_rep_rows(), chunk_images(c) > 1);
                edge_bwd_chunk<S>(w, gp<S>(nbuf), c, l, gp<S>(gx1), gp<S>(gn), st);
            }
            if (!chunks_closed) timed(P_SRC_REDUCE, planes<S>() * (n_edges * 4612.0 + n_nodes * 9216.0), st, [&] {
                for (int k = 0; k < planes<S>(); ++k)
                    launch_source_reduce(plane(gp<S>(Gbuf), k), sptr.i(), sedge.i(), n_nodes, plane(gp<S>(gn), k), st); });
            save_dbg("l" + std::to_string(l) + ".g_n1", gn.v.p, (size_t)n_nodes * 9 * C, st);
            launch_rms_bwd_t<S>(gp<S>(xs[l]), w.n1w, gp<S>(gn), gp<S>(gx1), n_nodes, gp<S>(gx), st);    // g_x_l
            save_dbg("l" + std::to_string(l) + ".g_x", gx.v.p, (size_t)n_nodes * 9 * C, st);
        }
        // edge-degree embedding adjoint
        for (const Chunk& c : chunks) {
            RepScope rs(this, chunk_rep_rows(), chunk_images(c) > 1);
            const EB<S> b = bufs_for<S>(-1, c);
            radial_fwd<S>(ed_rad, c, b, st);
            launch_rotate_back_bwd_t<S>(1, b.rad, gnull<S>(), gnull<S>(), tgt.i(), gp<S>(wig), gp<S>(env),
                                        1.0f / cfg.edge_degree_rescale, c.e0, c.n_e, gp<S>(gx), b.ged, anull<S>(),
                                        anull<S>(), gp<S>(g_env), gp<S>(g_wig), st, share_e(c));
            radial_bwd<S>(ed_rad, c, b, b.ged, st);
        }
        save_dbg("g_gauss", g_gauss.v.p, (size_t)n_edges * NB, st);
        save_dbg("g_env", g_env.v.p, (size_t)n_edges, st);
        save_dbg("g_tau", g_wig.v.p, (size_t)n_edges * 4, st);
        launch_geometry_bwd_t<S>(gp<S>(vec), gp<S>(dist), gp<S>(wig), gp<S>(gauss), gp<S>(g_gauss), gp<S>(g_env),
                                 gp<S>(g_wig), (int)n_edges, cfg.cutoff, gp<S>(g_vec), st);
        save_dbg("g_vec", g_vec.v.p, (size_t)n_edges * 3, st);
        for (int k = 0; k < planes<S>(); ++k)
            launch_force_reduce(plane(gp<S>(g_vec), k), row_ptr.i(), sptr.i(), sedge.i(), n_nodes, plane(forces, k), st);
    }

    // give the per-call device memory back (edge workspace, per-layer stores, per-edge and per-node state): everything
    // re-grows on the next call.  Weights, their bf16 planes and the learned edge capacity stay.
    void release_workspace() {
        drop_graphs();
        TBuf* tall[] = {&vec, &dist, &env, &wig, &gauss, &g_gauss, &g_env, &g_wig, &g_vec, &nbuf, &abuf, &gx, &gx1,
                        &gn, &ggp, &p1, &s1, &p2, &gp2, &gs1, &Gbuf, &wA, &wY, &wB, &wZ, &wRAD, &wU1, &wH1, &wU2, &wH2, &wGY, &wGZ};
        for (TBuf* b : tall) b->release();
        for (auto* v : {&xs, &x1s, &y1s, &gps, &ystore, &zstore, &rstore, &u1store, &u2store}) for (auto& b : *v) b.release();
        DevBuf* all[] = {&src, &tgt, &stmp, &sedge, &node_e, &f_dev, &t_dev, &df_dev, &pos_own, &wSG};
        for (DevBuf* b : all) b->release();
        for (auto& kv : dbg) kv.second.first.release();
        dbg.clear();
    }
    void drop_graphs() {
        for (auto& g : graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
        graphs.clear();
    }
    ~umab_engine() {
        drop_graphs();
        if (own_stream) cudaStreamDestroy(own_stream);
        if (order_ev) cudaEventDestroy(order_ev);
        if (status_ev) cudaEventDestroy(status_ev);
        if (h_status) cudaFreeHost(h_status);
        status_dev.release();
        tc_cache_destroy(tc_cache);
        tc2_cache_destroy(tc2_cache);
        for (auto& kv : weights) kv.second.buf.release();
        DevBuf* all[] = {&z1, &pos_own, &zt, &deg, &thr, &row_ptr, &src, &tgt, &odeg, &sptr, &cursor, &stmp, &sedge,
                         &cgrid, &ccount, &cstart, &catoms, &acell, &node_e, &e_dev, &f_dev, &t_dev, &df_dev, &wSG, &status_dev};
        for (DevBuf* b : all) b->release();
        TBuf* tall[] = {&vec, &dist, &env, &wig, &gauss, &g_gauss, &g_env, &g_wig, &g_vec, &nbuf, &abuf, &gx, &gx1,
                        &gn, &ggp, &p1, &s1, &p2, &gp2, &gs1, &Gbuf, &wA, &wY, &wB, &wZ, &wRAD, &wU1, &wH1, &wU2, &wH2, &wGY, &wGZ};
        for (TBuf* b : tall) b->release();
        for (auto* v : {&xs, &x1s, &y1s, &gps, &ystore, &zstore, &rstore, &u1store, &u2store}) for (auto& b : *v) b.release();
        for (auto& kv : dbg) kv.second.first.release();
        if (h_pinned) cudaFreeHost(h_pinned);
        if (hp_pos) cudaFreeHost(hp_pos);
        if (hp_f) cudaFreeHost(hp_f);
        if (hp_e) cudaFreeHost(hp_e);
        prof.collect();
        for (auto ev : prof.pool) cudaEventDestroy(ev);
    }
};

// ====================================================================== C ABI
#define UMAB_TRY try {
#define UMAB_CATCH                                                        \
    }                                                                     \
    catch (const std::exception& ex) { g_last_error = ex.what(); return 1; } \
    catch (...) { g_last_error = "unknown error"; return 1; }             \
    return 0;

extern "C" {

int32_t umab_abi_version(void) { return UMAB_ABI_VERSION; }
const char* umab_last_error(void) { return g_last_error.c_str(); }

int32_t umab_create(const umab_config* cfg, umab_engine** out) {
    UMAB_TRY
    if (!cfg || !out) throw CudaError("null argument");
    if (cfg->sphere_channels != C || cfg->hidden_channels != H || cfg->num_distance_basis != NB)
        throw CudaError("this build is compiled for sphere_channels = hidden_channels = 128, 64 distance basis functions");
    if (cfg->num_layers < 1 || cfg->num_layers > 16) throw CudaError("num_layers out of range");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        throw CudaError("no CUDA device available: umab has no CPU fallback");
    }
    DeviceGuard guard(cfg->device);
    cudaDeviceProp prop;
    UMAB_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) throw CudaError("umab kernels are built for sm_100a (Blackwell) only");
    auto* eng = new umab_engine();
    eng->cfg = *cfg;
    *out = eng;
    UMAB_CATCH
}

void umab_destroy(umab_engine* e) {
    if (!e) return;
    try {
        DeviceGuard guard(e->cfg.device);
        delete e;
    } catch (...) {
        delete e;
    }
}

int32_t umab_set_weight(umab_engine* e, const char* name, const float* host, size_t numel) {
    UMAB_TRY
    if (!e || !name || !host) throw CudaError("null argument");
    DeviceGuard guard(e->cfg.device);
    Weight& w = e->weights[name];
    w.buf.ensure(std::max<size_t>(numel, 4) * sizeof(float));
    w.numel = numel;
    UMAB_CUDA(cudaMemcpy(w.buf.p, host, numel * sizeof(float), cudaMemcpyHostToDevice));
    e->finalized = false;
    tc_cache_clear(e->tc_cache);                 // bf16 planes are derived from the fp32 weights
    tc2_cache_clear(e->tc2_cache);
    UMAB_CATCH
}

int32_t umab_finalize_weights(umab_engine* e) {
    UMAB_TRY
    if (!e) throw CudaError("null argument");
    e->finalize();
    UMAB_CATCH
}

int32_t umab_set_system(umab_engine* e, const int32_t* z_host, int32_t n_atoms) {
    UMAB_TRY
    if (!e || !z_host || n_atoms <= 0) throw CudaError("bad argument");
    for (int i = 0; i < n_atoms; ++i)
        if (z_host[i] < 0 || z_host[i] >= 100) throw CudaError("atomic number outside the embedding table");
    DeviceGuard guard(e->cfg.device);
    e->z1.ensure(sizeof(int) * n_atoms);
    UMAB_CUDA(cudaMemcpy(e->z1.p, z_host, sizeof(int) * n_atoms, cudaMemcpyHostToDevice));
    e->n_atoms = n_atoms;
    e->zt_img = -1;
    UMAB_CATCH
}

int32_t umab_set_option(umab_engine* e, const char* name, int64_t value) {
    UMAB_TRY
    if (!e || !name) throw CudaError("null argument");
    const std::string n(name);
    if (n == "neighbor_mode") {
        if (value < 0 || value > 2) throw CudaError("neighbor_mode: 0 auto, 1 brute force, 2 cell list");
        e->neighbor_mode = (int)value;
    } else if (n == "nosync") {
        e->opt_nosync = value != 0;
    } else if (n == "fuse_gate") {
        e->opt_fuse_gate = value != 0;
        e->drop_graphs();
    } else if (n == "jvp_shared_base") {
        // umab_forces_jvp batches whose images all sit at ONE base geometry (Hessian columns): value-plane GEMMs once
        e->opt_shared_base = value != 0;
    } else if (n == "cuda_graphs") {
        e->opt_graphs = value != 0;
        if (!e->opt_graphs) e->drop_graphs();
    } else if (n == "simt_round_fwd" || n == "simt_round_bwd") {
        // precision study: operand rounding emulated by the fp32 SIMT GEMM (GemmArgs::round_mode), forward / adjoint GEMMs
        if (value < 0 || value > 4) throw CudaError("simt_round_*: 0 exact, 1 tf32, 2 bf16x2 (W bf16), 3 bf16x2 (A bf16), 4 bf16");
        (n == "simt_round_fwd" ? e->simt_round_fwd : e->simt_round_bwd) = (int)value;
    } else {
        throw CudaError("unknown option: " + n);
    }
    UMAB_CATCH
}

int32_t umab_release_workspace(umab_engine* e) {
    UMAB_TRY
    if (!e) throw CudaError("null argument");
    DeviceGuard guard(e->cfg.device);
    e->resolve_status();
    UMAB_CUDA(cudaDeviceSynchronize());
    e->release_workspace();
    UMAB_CATCH
}

int32_t umab_get_option(umab_engine* e, const char* name, int64_t* value) {
    UMAB_TRY
    if (!e || !name || !value) throw CudaError("null argument");
    const std::string n(name);
    if (n == "nosync") *value = e->opt_nosync;
    else if (n == "cuda_graphs") *value = e->opt_graphs;
    else if (n == "jvp_shared_base") *value = e->opt_shared_base;
    else if (n == "dedupe_gemms") *value = e->dedupe_gemms;
    else if (n == "fuse_gate") *value = e->opt_fuse_gate;
    else if (n == "graph_replays") *value = e->graph_replays;
    else if (n == "graph_captures") *value = e->graph_captures;
    else if (n == "overflow_retries") *value = e->overflow_retries;
    else if (n == "fast_calls") *value = e->fast_calls;
    else if (n == "edges_per_image_seen") *value = e->hist_edges_per_image;
    else if (n == "neighbor_mode") *value = e->neighbor_mode;
    else throw CudaError("unknown option: " + n);
    UMAB_CATCH
}

int32_t umab_build_graph(umab_engine* e, const float* pos_dev, int32_t n_images, void* stream) {
    UMAB_TRY
    if (!e || !pos_dev) throw CudaError("null argument");
    DeviceGuard guard(e->cfg.device);
    e->build_graph(pos_dev, n_images, (cudaStream_t)stream);
    e->plan_chunks<float>();
    UMAB_CATCH
}

int32_t umab_graph_counts(umab_engine* e, int64_t* n_nodes, int64_t* n_edges) {
    UMAB_TRY
    if (!e) throw CudaError("null argument");
    if (n_nodes) *n_nodes = e->n_nodes;
    // sync-free evaluations: the edge arrays are capacity-sized; report the actual count once it has been collected
    if (n_edges) *n_edges = (e->fast_graph && !e->status_pending) ? e->n_edges_actual : e->n_edges;
    UMAB_CATCH
}

int32_t umab_graph_copy(umab_engine* e, int32_t* src_dev, int32_t* tgt_dev, int32_t* row_ptr_dev, void* stream) {
    UMAB_TRY
    if (!e) throw CudaError("null argument");
    DeviceGuard guard(e->cfg.device);
    cudaStream_t st = (cudaStream_t)stream;
    if (src_dev && e->n_edges) UMAB_CUDA(cudaMemcpyAsync(src_dev, e->src.p, sizeof(int) * e->n_edges, cudaMemcpyDeviceToDevice, st));
    if (tgt_dev && e->n_edges) UMAB_CUDA(cudaMemcpyAsync(tgt_dev, e->tgt.p, sizeof(int) * e->n_edges, cudaMemcpyDeviceToDevice, st));
    if (row_ptr_dev) UMAB_CUDA(cudaMemcpyAsync(row_ptr_dev, e->row_ptr.p, sizeof(int) * (e->n_nodes + 1), cudaMemcpyDeviceToDevice, st));
    UMAB_CATCH
}

int32_t umab_energy_forces(umab_engine* e, const float* pos_dev, int32_t n_images, double* energy_dev,
                           float* forces_dev, void* stream) {
    UMAB_TRY
    if (!e || !pos_dev || !energy_dev) throw CudaError("null argument");
    DeviceGuard guard(e->cfg.device);
    e->evaluate_batched<float>(gpf(pos_dev), n_images, energy_dev, GP<float>{forces_dev}, (cudaStream_t)stream);
    UMAB_CATCH
}

int32_t umab_last_call(umab_engine* e, int64_t* n_images, int64_t* n_edges, int64_t* n_subcalls) {
    if (!e) { g_last_error = "null argument"; return 1; }
    try {
        DeviceGuard guard(e->cfg.device);
        const bool was_pending = e->status_pending;
        const bool overflow = e->resolve_status();
        if (was_pending) e->call_edges += e->n_edges_actual;
        if (n_images) *n_images = e->call_images;
        if (n_edges) *n_edges = e->call_edges;
        if (n_subcalls) *n_subcalls = e->call_subcalls;
        if (overflow) {
            g_last_error = "edge capacity exceeded in the last sync-free call: its results are invalid, repeat the call";
            return UMAB_RETRY;
        }
    } catch (const std::exception& ex) { g_last_error = ex.what(); return 1; }
    return 0;
}

int32_t umab_energy_forces_host(umab_engine* e, const float* pos_host, int32_t n_images, double* energy_host,
                                float* forces_host, void* stream) {
    UMAB_TRY
    if (!e || !pos_host || !energy_host) throw CudaError("null argument");
    if (e->n_atoms <= 0) throw CudaError("umab_set_system has not been called");
    if (n_images <= 0) throw CudaError("n_images must be positive");
    DeviceGuard guard(e->cfg.device);
    cudaStream_t st = (cudaStream_t)stream;
    if (st == nullptr || st == cudaStreamLegacy) {
        // the legacy default stream cannot be captured into a CUDA graph: host-buffer calls are self-contained, so they
        // run on a stream of the engine, ordered after whatever the caller has queued on its stream so far
        if (!e->own_stream) {
            UMAB_CUDA(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
            UMAB_CUDA(cudaEventCreateWithFlags(&e->order_ev, cudaEventDisableTiming));
        }
        UMAB_CUDA(cudaEventRecord(e->order_ev, st));
        UMAB_CUDA(cudaStreamWaitEvent(e->own_stream, e->order_ev, 0));
        st = e->own_stream;
    }
    size_t nb = (size_t)n_images * e->n_atoms * 3 * sizeof(float);
    e->pos_own.ensure(nb);
    e->e_dev.ensure(sizeof(double) * n_images);
    if (forces_host) e->f_dev.ensure(nb);
    // pinned staging on both sides so the copies are true async DMA transfers
    if (nb > e->hp_cap) {
        if (e->hp_pos) cudaFreeHost(e->hp_pos);
        if (e->hp_f) cudaFreeHost(e->hp_f);
        UMAB_CUDA(cudaMallocHost(&e->hp_pos, nb + nb / 4));
        UMAB_CUDA(cudaMallocHost(&e->hp_f, nb + nb / 4));
        e->hp_cap = nb + nb / 4;
        ++g_alloc_gen;
    }
    if (sizeof(double) * n_images > e->hp_ecap) {
        if (e->hp_e) cudaFreeHost(e->hp_e);
        UMAB_CUDA(cudaMallocHost(&e->hp_e, sizeof(double) * n_images * 2));
        e->hp_ecap = sizeof(double) * n_images * 2;
        ++g_alloc_gen;
    }
    memcpy(e->hp_pos, pos_host, nb);
    const bool want_f = forces_host != nullptr;
    auto enqueue_all = [&] {
        UMAB_CUDA(cudaMemcpyAsync(e->pos_own.p, e->hp_pos, nb, cudaMemcpyHostToDevice, st));
        e->evaluate_batched<float>(GP<float>{e->pos_own.f()}, n_images, e->e_dev.as<double>(),
                                   GP<float>{want_f ? e->f_dev.f() : nullptr}, st);
        UMAB_CUDA(cudaMemcpyAsync(e->hp_e, e->e_dev.p, sizeof(double) * n_images, cudaMemcpyDeviceToHost, st));
        if (want_f) UMAB_CUDA(cudaMemcpyAsync(e->hp_f, e->f_dev.p, nb, cudaMemcpyDeviceToHost, st));
    };
    for (int attempt = 0;; ++attempt) {
        // launch-bound calls (one sub-batch, one chunk, sync-free graph build): the whole enqueue sequence -- H2D, ~200-700
        // kernel launches, D2H -- is captured once per (images, capacity) and replayed as ONE cudaGraphLaunch
        e->resolve_status();
        const bool one = e->images_per_call(want_f, 1) >= n_images;
        const long long fcap = one ? e->fast_capacity<float>(n_images) : 0;
        umab_engine::GraphEntry* ge = nullptr;
        if (fcap > 0 && e->graphs_enabled()) {
            for (auto& g : e->graphs) if (g.nimg == n_images && g.fcap == fcap && g.want_f == want_f) ge = &g;
            if (!ge) {
                if (e->graphs.size() >= 16) e->drop_graphs();
                e->graphs.push_back(umab_engine::GraphEntry{});
                ge = &e->graphs.back();
                ge->nimg = n_images; ge->fcap = fcap; ge->want_f = want_f;
            }
            if (ge->gen != g_alloc_gen.load()) {          // a buffer moved since: the captured addresses are stale
                if (ge->exec) { cudaGraphExecDestroy(ge->exec); ge->exec = nullptr; }
                ge->warm = 0;
            }
        }
        bool launched = false;
        if (ge && ge->exec) {
            UMAB_CUDA(cudaGraphLaunch(ge->exec, st));
            ++e->graph_replays;
            g_launch_count += ge->n_launch;            // the kernels of the replayed graph
            e->n_img = n_images; e->n_nodes = ge->n_nodes; e->n_edges = fcap; e->fast_graph = true;
            e->call_images = n_images; e->call_edges = 0; e->call_subcalls = 1;
            launched = true;
        } else if (ge && ge->warm >= 1) {
            // second call with stable buffers: capture.  Anything unexpected (an allocation, a sync) fails the capture;
            // the call then simply runs un-captured.
            cudaGraph_t graph = nullptr;
            const long long launches0 = g_launch_count.load();
            static const bool dbg_graph = getenv("UMAB_DEBUG_GRAPH") != nullptr;
            cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
            bool ok = ce == cudaSuccess;
            if (!ok && dbg_graph) fprintf(stderr, "umab: BeginCapture failed: %s\n", cudaGetErrorString(ce));
            if (ok) {
                e->capturing = true;
                try { enqueue_all(); } catch (const std::exception& ex) {
                    ok = false;
                    if (dbg_graph) fprintf(stderr, "umab: capture aborted: %s\n", ex.what());
                }
                e->capturing = false;
                ce = cudaStreamEndCapture(st, &graph);
                if (ce != cudaSuccess || !graph) {
                    ok = false;
                    if (dbg_graph) fprintf(stderr, "umab: EndCapture failed: %s\n", cudaGetErrorString(ce));
                }
            }
            if (ok && e->fast_graph && ge->gen == g_alloc_gen.load() &&
                cudaGraphInstantiate(&ge->exec, graph, 0) == cudaSuccess) {
                ge->n_nodes = e->n_nodes;
                ge->n_launch = g_launch_count.load() - launches0;
                UMAB_CUDA(cudaGraphLaunch(ge->exec, st));
                ++e->graph_captures;
                launched = true;
            } else {
                cudaGetLastError();
                ge->exec = nullptr; ge->warm = -1000000;       // never try again for this key
            }
            if (graph) cudaGraphDestroy(graph);
        }
        if (!launched) {
            enqueue_all();
            if (ge) { ge->warm += 1; ge->gen = g_alloc_gen.load(); }
        }
        UMAB_CUDA(cudaStreamSynchronize(st));
        bool overflow = false;
        if (e->fast_graph) {                 // the status words were copied by the same stream: valid after the sync
            e->status_pending = false;
            e->n_edges_actual = e->h_status[0];
            e->hist_edges_per_image = std::max<long long>(e->hist_edges_per_image, (e->n_edges_actual + n_images - 1) / n_images);
            e->call_edges += e->n_edges_actual;
            overflow = e->h_status[1] != 0;
        }
        if (!overflow) break;
        ++e->overflow_retries;
        if (attempt >= 3) throw CudaError("edge capacity exceeded repeatedly");
    }
    memcpy(energy_host, e->hp_e, sizeof(double) * n_images);
    if (forces_host) memcpy(forces_host, e->hp_f, nb);
    UMAB_CATCH
}

int32_t umab_forces_jvp(umab_engine* e, const float* pos_dev, const float* tangent_dev, int32_t n_images,
                        double* energy_dev, float* forces_dev, float* dforces_dev, void* stream) {
    UMAB_TRY
    if (!e || !pos_dev || !tangent_dev || !dforces_dev) throw CudaError("null argument");
    DeviceGuard guard(e->cfg.device);
    cudaStream_t st = (cudaStream_t)stream;
    if (!forces_dev) {
        e->f_dev.ensure((size_t)n_images * e->n_atoms * 3 * sizeof(float));
        forces_dev = e->f_dev.f();
    }
    e->evaluate_batched<D1>(GP<D1>{const_cast<float*>(pos_dev), const_cast<float*>(tangent_dev)}, n_images, energy_dev,
                            GP<D1>{forces_dev, dforces_dev}, st);
    UMAB_CATCH
}

int32_t umab_hessian_fd_columns(const float* forces_dev, const int32_t* dof_idx_dev, int32_t n_cols, int32_t dof,
                                double h_step, void* hessian_dev, int64_t ld, int32_t is_f64, void* stream) {
    UMAB_TRY
    if (!forces_dev || !dof_idx_dev || !hessian_dev || n_cols < 0 || dof <= 0 || ld < dof || !(h_step > 0.0))
        throw CudaError("bad argument");
    launch_fd_columns(forces_dev, dof_idx_dev, n_cols, dof, h_step, hessian_dev, ld, is_f64 != 0, (cudaStream_t)stream);
    UMAB_CATCH
}

int64_t umab_hessian_mw_workspace(int32_t n, int32_t r) {
    if (n <= 0 || r < 0 || r > 6) return -1;
    return (int64_t)mw_project_workspace_doubles(n, r);
}

int32_t umab_hessian_mw_project(double* hessian_dev, int32_t n, const double* inv_sqrt_m_dev, const double* q_dev, int32_t r,
                                double* workspace_dev, int64_t workspace_doubles, void* stream) {
    UMAB_TRY
    if (!hessian_dev || !inv_sqrt_m_dev || n <= 0 || r < 0 || r > 6 || (r > 0 && (!q_dev || !workspace_dev)))
        throw CudaError("bad argument");
    if (r > 0 && workspace_doubles < (int64_t)mw_project_workspace_doubles(n, r)) throw CudaError("workspace too small");
    launch_mw_project(hessian_dev, n, inv_sqrt_m_dev, q_dev, r, workspace_dev, (cudaStream_t)stream);
    UMAB_CATCH
}

int32_t umab_gemm(int32_t mode, const float* a_dev, const float* w_dev, const float* bias_dev, float* c_dev,
                  int64_t m, int32_t n, int32_t k, void* stream) {
    UMAB_TRY
    GemmArgs g;
    g.A = a_dev; g.lda = k; g.W = w_dev; g.ldw = k; g.Cmat = c_dev; g.ldc = n; g.bias = bias_dev;
    g.M = (int)m; g.N = n; g.K = k;
    if (mode == 3 || mode == 5) {
        // TMA-fed kernel (gemm_tc2.cu): single-CTA (mode 3) / CTA pair (mode 5, the engine's default); the fp32
        // activation is split into planes for the call
        if (!gemm_tc2_supported(g, 64)) throw CudaError("shape not supported by the TMA-fed tensor-core GEMM");
        gemm_tc2(g, (cudaStream_t)stream, nullptr, 64, mode == 5 ? 1 : 0);
    } else if (mode == 1 || mode == 2) {
        if (!gemm_tc_supported(g)) throw CudaError("shape not supported by the tensor-core GEMM");
        // mode 1: weight planes rebuilt on every call (never cached by pointer);
        // mode 2: planes cached by pointer for timing loops (the caller keeps W alive and unchanged)
        static TcPlaneCache* bench_cache = tc_cache_create();
        gemm_tc(g, (cudaStream_t)stream, mode == 2 ? bench_cache : nullptr);
    } else if (mode == 7) {
        gemm_simt_splitk(g, (cudaStream_t)stream);
    } else {
        gemm_simt(g, (cudaStream_t)stream);
    }
    UMAB_CATCH
}

int32_t umab_gemm_bench(int32_t mode, const float* a_dev, const float* w_dev, float* c_dev, int64_t m, int32_t n,
                        int32_t k, int32_t iters, double* ms_per_iter, void* stream) {
    UMAB_TRY
    if (!a_dev || !w_dev || !c_dev || !ms_per_iter || iters <= 0) throw CudaError("bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    GemmArgs g;
    g.A = a_dev; g.lda = k; g.W = w_dev; g.ldw = k; g.Cmat = c_dev; g.ldc = n; g.M = (int)m; g.N = n; g.K = k;
    static TcPlaneCache* c1 = tc_cache_create();
    static Tc2Cache* c2 = tc2_cache_create();
    __nv_bfloat16 *hi = nullptr, *lo = nullptr;
    const int bk = 64;
    const bool tc2 = mode == 3 || mode == 5;
    if (tc2) {
        if (!gemm_tc2_supported(g, bk)) throw CudaError("shape not supported by the TMA-fed tensor-core GEMM");
        UMAB_CUDA(cudaMalloc(&hi, (size_t)m * k * 2));
        UMAB_CUDA(cudaMalloc(&lo, (size_t)m * k * 2));
        tc2_split(a_dev, (long long)m * k, hi, lo, st);
        g.A_hi = hi; g.A_lo = lo;
    } else if (mode == 1 || mode == 2) {
        if (!gemm_tc_supported(g)) throw CudaError("shape not supported by the tensor-core GEMM");
    }
    auto run = [&] {
        if (tc2) gemm_tc2(g, st, c2, bk, mode >= 5 ? 1 : 0);
        else if (mode == 1 || mode == 2) gemm_tc(g, st, c1);
        else gemm_simt(g, st);
    };
    run();                                             // warm-up: weight planes, tensor maps
    cudaEvent_t e0, e1;
    UMAB_CUDA(cudaEventCreate(&e0));
    UMAB_CUDA(cudaEventCreate(&e1));
    UMAB_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) run();
    UMAB_CUDA(cudaEventRecord(e1, st));
    UMAB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    UMAB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_iter = (double)ms / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    tc_cache_clear(c1); tc2_cache_clear(c2);          // the caller's W / A pointers are not stable across calls
    if (hi) { cudaFree(hi); cudaFree(lo); }
    UMAB_CATCH
}

int32_t umab_debug_tensor(umab_engine* e, const char* name, const float** ptr_dev, size_t* numel) {
    UMAB_TRY
    if (!e || !name) throw CudaError("null argument");
    auto it = e->dbg.find(name);
    if (it == e->dbg.end()) throw CudaError(std::string("no debug tensor named ") + name);
    if (ptr_dev) *ptr_dev = it->second.first.f();
    if (numel) *numel = it->second.second;
    UMAB_CATCH
}

int32_t umab_profile(umab_engine* e, int32_t enable) {
    UMAB_TRY
    if (!e) throw CudaError("null argument");
    DeviceGuard guard(e->cfg.device);
    e->prof.reset();
    e->prof.on = enable != 0;
    UMAB_CATCH
}

int32_t umab_profile_read(umab_engine* e, int32_t cat, double* ms, int64_t* launches, double* work, double* bytes) {
    UMAB_TRY
    if (!e || cat < 0 || cat >= P_COUNT) throw CudaError("bad argument");
    DeviceGuard guard(e->cfg.device);
    e->prof.collect();
    if (ms) *ms = e->prof.ms[cat];
    if (launches) *launches = e->prof.n[cat];
    if (work) *work = e->prof.work[cat];
    if (bytes) *bytes = e->prof.bytes[cat];
    UMAB_CATCH
}

const char* umab_profile_name(int32_t cat) { return (cat >= 0 && cat < P_COUNT) ? kProfNames[cat] : nullptr; }

int32_t umab_stats(umab_engine* e, int64_t* kernel_launches, int64_t* device_bytes) {
    UMAB_TRY
    (void)e;
    if (kernel_launches) *kernel_launches = g_launch_count.load();
    if (device_bytes) *device_bytes = DevBuf::total.load();
    UMAB_CATCH
}

}  // extern "C"
