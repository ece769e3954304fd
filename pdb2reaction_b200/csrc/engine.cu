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
#include "common.cuh"

namespace umab {

std::atomic<long long> g_launch_count{0};
struct TcPlaneCache;                                  // gemm_tc.cu
TcPlaneCache* tc_cache_create();
void tc_cache_destroy(TcPlaneCache* c);
void tc_cache_clear(TcPlaneCache* c);
void gemm_tc(const GemmArgs& a, cudaStream_t st, TcPlaneCache* cache);
bool gemm_tc_supported(const GemmArgs& a);

namespace {

thread_local std::string g_last_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    static long long total;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if (p) { cudaFree(p); total -= (long long)cap; }
        p = nullptr; cap = 0;
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
    void release() { if (p) { cudaFree(p); total -= (long long)cap; } p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
    float* f() const { return reinterpret_cast<float*>(p); }
    int* i() const { return reinterpret_cast<int*>(p); }
};
long long DevBuf::total = 0;

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

constexpr size_t EDGE_WS_FLOATS = 2304 + 2176 + 1152 + 1920 + 1536 + 4 * 128;   // 9600 per edge

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
    std::vector<int> h_row_ptr;
    int* h_pinned = nullptr; size_t h_pinned_cap = 0;
    std::vector<Chunk> chunks;
    // geometry
    DevBuf vec, dist, env, wig, gauss, g_gauss, g_env, g_wig, g_vec;
    // nodes
    std::vector<DevBuf> xs, x1s, y1s, gps;
    DevBuf nbuf, abuf, gx, gx1, gn, ggp, p1, s1, p2, gp2, gs1, node_e, Gbuf;
    // edge workspace
    DevBuf wA, wY, wB, wZ, wRAD, wU1, wH1, wU2, wH2;
    long long chunk_cap = 0;
    // store mode: conv-1 / conv-2 outputs of every layer are kept for the backward instead of being
    // recomputed (16.4 KB per edge and layer of HBM) when they fit `store_bytes`
    std::vector<DevBuf> ystore, zstore;
    bool store_mode = false;
    // host staging for umab_energy_forces_host
    DevBuf e_dev, f_dev;
    // debug
    std::map<std::string, std::pair<DevBuf, size_t>> dbg;
    TcPlaneCache* tc_cache = tc_cache_create();       // bf16 planes of this engine's weights
    // profiling + pinned host staging
    Prof prof;
    void* hp_pos = nullptr; void* hp_f = nullptr; void* hp_e = nullptr; size_t hp_cap = 0, hp_ecap = 0;

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
            w.c1m1 = W(p + ".edge.conv1.fc_m1.weight", 512 * 512); w.c1m1_t = W(p + ".edge.conv1.fc_m1.weight_t", 512 * 512);
            w.c1m2 = W(p + ".edge.conv1.fc_m2.weight", 256 * 256); w.c1m2_t = W(p + ".edge.conv1.fc_m2.weight_t", 256 * 256);
            w.c2m0 = W(p + ".edge.conv2.fc_m0.weight", 384 * 384); w.c2m0_t = W(p + ".edge.conv2.fc_m0.weight_t", 384 * 384);
            w.c2m0_b = W(p + ".edge.conv2.fc_m0.bias", 384);
            w.c2m1 = W(p + ".edge.conv2.fc_m1.weight", 512 * 256); w.c2m1_t = W(p + ".edge.conv2.fc_m1.weight_t", 512 * 256);
            w.c2m2 = W(p + ".edge.conv2.fc_m2.weight", 256 * 128); w.c2m2_t = W(p + ".edge.conv2.fc_m2.weight_t", 256 * 128);
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
    bool use_tc() const { return cfg.gemm_mode == 1 || (cfg.gemm_mode == 2 && n_atoms >= 100); }
    void gemm(const GemmArgs& a, cudaStream_t st) {
        timed(P_GEMM, 2.0 * a.M * (double)a.N * a.K * a.batch, st, [&] {
            if (use_tc() && gemm_tc_supported(a)) gemm_tc(a, st, tc_cache);
            else gemm_simt(a, st);
        }, /* algorithmic bytes: A and C once (+C again when accumulating), W once */
           4.0 * a.batch * ((double)a.M * a.K + (double)a.M * a.N * (a.accumulate ? 2 : 1)) + 4.0 * a.N * (double)a.K);
    }
    void mm(const float* A, long long lda, const float* Wt, int N, int K, float* Cm, long long ldc, long long M,
            const float* bias, int accumulate, cudaStream_t st) {
        GemmArgs g;
        g.A = A; g.lda = lda; g.W = Wt; g.ldw = K; g.Cmat = Cm; g.ldc = ldc; g.bias = bias;
        g.M = (int)M; g.N = N; g.K = K; g.accumulate = accumulate;
        gemm(g, st);
    }
    // per-coefficient SO(3) linear: rows (n, i) use weight l(i);  A [N,9,K] -> C [N,9,Nout]
    void so3_mm(const float* A, int K, const float* Wt, int Nout, float* Cm, const float* bias, int accumulate,
                cudaStream_t st) {
        GemmArgs g;
        g.A = A; g.lda = 9LL * K; g.strideA = K;
        g.W = Wt; g.ldw = K; g.strideW = (long long)Nout * K;
        g.Cmat = Cm; g.ldc = 9LL * Nout; g.strideC = Nout;
        g.bias = bias; g.bias_first_batch_only = 1;
        g.M = n_nodes; g.N = Nout; g.K = K; g.batch = 9; g.accumulate = accumulate;
        const int lsel[9] = {0, 1, 1, 1, 2, 2, 2, 2, 2};
        for (int i = 0; i < 9; ++i) g.wsel[i] = lsel[i];
        gemm(g, st);
    }
    void save_dbg(const std::string& name, const void* p, size_t numel, cudaStream_t st) {
        if (!cfg.debug) return;
        auto& slot = dbg[name];
        slot.first.ensure(numel * 4);
        slot.second = numel;
        UMAB_CUDA(cudaMemcpyAsync(slot.first.p, p, numel * 4, cudaMemcpyDeviceToDevice, st));
    }

    // ------------------------------------------------------------------ graph
    void build_graph(const float* pos, int nimg, cudaStream_t st) {
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
        launch_neighbor_count(pos, nimg, n_atoms, cfg.cutoff, cfg.max_neighbors, deg.i(), thr.f(), st);
        launch_scan(deg.i(), row_ptr.i(), n_nodes, st);
        size_t need = sizeof(int) * (n_nodes + 1);
        if (need > h_pinned_cap) {
            if (h_pinned) cudaFreeHost(h_pinned);
            UMAB_CUDA(cudaMallocHost(&h_pinned, need + need / 4));
            h_pinned_cap = need + need / 4;
        }
        UMAB_CUDA(cudaMemcpyAsync(h_pinned, row_ptr.p, need, cudaMemcpyDeviceToHost, st));
        UMAB_CUDA(cudaStreamSynchronize(st));
        n_edges = h_pinned[n_nodes];
        if (n_edges > 0x7fffffffLL / 40) throw CudaError("too many edges in one batch: split it on the host");
        size_t ne = (size_t)std::max<long long>(n_edges, 1);
        src.ensure(sizeof(int) * ne); tgt.ensure(sizeof(int) * ne);
        launch_neighbor_fill(pos, nimg, n_atoms, cfg.cutoff, cfg.max_neighbors, thr.f(), row_ptr.i(), src.i(), tgt.i(), st);
        odeg.ensure(sizeof(int) * n_nodes); sptr.ensure(sizeof(int) * (n_nodes + 1)); cursor.ensure(sizeof(int) * n_nodes);
        stmp.ensure(sizeof(int) * ne); sedge.ensure(sizeof(int) * ne);
        launch_source_csr(src.i(), (int)n_edges, n_nodes, odeg.i(), sptr.i(), cursor.i(), stmp.i(), sedge.i(), st);
        plan_chunks();
    }

    void plan_chunks() {
        long long budget = cfg.workspace_bytes > 0 ? cfg.workspace_bytes : (32LL << 30);
        long long cap = std::max<long long>(budget / (long long)(EDGE_WS_FLOATS * 4), 1024);
        chunks.clear();
        int node0 = 0;
        long long biggest = 0;
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
        wA.ensure(f * 2304); wY.ensure(f * 2176); wB.ensure(f * 1152); wZ.ensure(f * 1920); wRAD.ensure(f * 1536);
        wU1.ensure(f * 128); wH1.ensure(f * 128); wU2.ensure(f * 128); wH2.ensure(f * 128);
    }

    // ------------------------------------------------------------------ edge stages
    float* A0() { return wA.f(); }
    float* A1() { return wA.f() + chunk_cap * 768; }
    float* A2() { return wA.f() + chunk_cap * (768 + 1024); }
    float* Y0() { return wY.f(); }
    float* Y1() { return wY.f() + chunk_cap * 640; }
    float* Y2() { return wY.f() + chunk_cap * (640 + 1024); }
    float* B0() { return wB.f(); }
    float* B1() { return wB.f() + chunk_cap * 384; }
    float* B2() { return wB.f() + chunk_cap * (384 + 512); }
    float* Z0() { return wZ.f(); }
    float* Z1() { return wZ.f() + chunk_cap * 384; }
    float* Z2() { return wZ.f() + chunk_cap * (384 + 1024); }

    struct YZ { float *y0, *y1, *y2, *z0, *z1, *z2; };
    YZ yz_for(int layer, const Chunk& c) {
        if (store_mode && layer >= 0) {
            float* yb = ystore[layer].f();
            float* zb = zstore[layer].f();
            const long long E = n_edges;
            return {yb + c.e0 * 640, yb + E * 640 + c.e0 * 1024, yb + E * 1664 + c.e0 * 512,
                    zb + c.e0 * 384, zb + E * 384 + c.e0 * 1024, zb + E * 1408 + c.e0 * 512};
        }
        return {Y0(), Y1(), Y2(), Z0(), Z1(), Z2()};
    }

    void radial_fwd(const RadialW& r, const Chunk& c, cudaStream_t st) {
        const long long e0 = c.e0;
        mm(gauss.f() + e0 * NB, NB, r.w1g, 128, NB, wU1.f(), 128, c.n_e, nullptr, 0, st);
        launch_ln_silu_fwd(wU1.f(), wH1.f(), r.ln1w, r.ln1b, r.b1, r.t_src, r.t_tgt, zt.i(), src.i() + e0, tgt.i() + e0, c.n_e, st);
        mm(wH1.f(), 128, r.w2, 128, 128, wU2.f(), 128, c.n_e, r.b2, 0, st);
        launch_ln_silu_fwd(wU2.f(), wH2.f(), r.ln2w, r.ln2b, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, c.n_e, st);
        mm(wH2.f(), 128, r.w3, r.n_out, 128, wRAD.f(), r.n_out, c.n_e, r.b3, 0, st);
    }
    // g_rad lives in wRAD [n_e, n_out]; accumulates into g_gauss
    void radial_bwd(const RadialW& r, const Chunk& c, cudaStream_t st) {
        mm(wRAD.f(), r.n_out, r.w3_t, 128, r.n_out, wH2.f(), 128, c.n_e, nullptr, 0, st);
        launch_ln_silu_bwd(wU2.f(), wH2.f(), r.ln2w, r.ln2b, c.n_e, st);
        mm(wH2.f(), 128, r.w2_t, 128, 128, wH1.f(), 128, c.n_e, nullptr, 0, st);
        launch_ln_silu_bwd(wU1.f(), wH1.f(), r.ln1w, r.ln1b, c.n_e, st);
        mm(wH1.f(), 128, r.w1g_t, NB, 128, g_gauss.f() + c.e0 * NB, NB, c.n_e, nullptr, 1, st);
    }
    // conv-1 radial, gather/rotate, conv-1, gate, conv-2 for one chunk (everything up to Z)
    void edge_fwd_chunk(const LayerW& w, const float* n1, const Chunk& c, int layer, bool dbg_on, cudaStream_t st) {
        const YZ b = yz_for(layer, c);
        radial_fwd(w.rad, c, st);
        timed(P_GATHER, c.n_e * 15512.0 + c.n_nodes * 4608.0, st, [&] {
            launch_gather_rotate_scale(n1, src.i(), tgt.i(), wig.f(), wRAD.f(), c.e0, c.n_e, A0(), A1(), A2(), st); });
        mm(A0(), 768, w.c1m0, 640, 768, b.y0, 640, c.n_e, w.c1m0_b, 0, st);
        mm(A1(), 512, w.c1m1, 512, 512, b.y1, 512, 2LL * c.n_e, nullptr, 0, st);
        mm(A2(), 256, w.c1m2, 256, 256, b.y2, 256, 2LL * c.n_e, nullptr, 0, st);
        timed(P_COMBINE, c.n_e * 13312.0, st, [&] { launch_combine_gate_fwd(b.y0, b.y1, b.y2, c.n_e, B0(), B1(), B2(), st); });
        mm(B0(), 384, w.c2m0, 384, 384, b.z0, 384, c.n_e, w.c2m0_b, 0, st);
        mm(B1(), 256, w.c2m1, 512, 256, b.z1, 512, 2LL * c.n_e, nullptr, 0, st);
        mm(B2(), 128, w.c2m2, 256, 128, b.z2, 256, 2LL * c.n_e, nullptr, 0, st);
        if (dbg_on && chunks.size() == 1) {
            std::string p = "l" + std::to_string(layer) + ".";
            save_dbg(p + "rad", wRAD.f(), (size_t)c.n_e * 1536, st);
            save_dbg(p + "a0", A0(), (size_t)c.n_e * 768, st);
            save_dbg(p + "a1", A1(), (size_t)c.n_e * 1024, st);
            save_dbg(p + "a2", A2(), (size_t)c.n_e * 512, st);
            save_dbg(p + "y0", b.y0, (size_t)c.n_e * 640, st);
            save_dbg(p + "y1", b.y1, (size_t)c.n_e * 1024, st);
            save_dbg(p + "y2", b.y2, (size_t)c.n_e * 512, st);
            save_dbg(p + "z0", b.z0, (size_t)c.n_e * 384, st);
            save_dbg(p + "z1", b.z1, (size_t)c.n_e * 1024, st);
            save_dbg(p + "z2", b.z2, (size_t)c.n_e * 512, st);
        }
    }
    void edge_bwd_chunk(const LayerW& w, const float* n1, const Chunk& c, int layer, const float* g_out, float* g_n1,
                        cudaStream_t st) {
        const YZ b = yz_for(layer, c);
        if (store_mode) radial_fwd(w.rad, c, st);            // only the radial weights are recomputed
        else edge_fwd_chunk(w, n1, c, layer, false, st);     // recompute everything up to Z
        timed(P_ROTBACK_BWD, c.n_e * (2 * 7680.0 + 2 * 148.0 + 8.0) + c.n_nodes * 4608.0, st, [&] {
            launch_rotate_back_bwd(0, b.z0, b.z1, b.z2, tgt.i(), wig.f(), env.f(), 1.0f, c.e0, c.n_e, g_out,
                                   b.z0, b.z1, b.z2, g_env.f(), g_wig.f(), st); });
        mm(b.z0, 384, w.c2m0_t, 384, 384, B0(), 384, c.n_e, nullptr, 0, st);
        mm(b.z1, 512, w.c2m1_t, 256, 512, B1(), 256, 2LL * c.n_e, nullptr, 0, st);
        mm(b.z2, 256, w.c2m2_t, 128, 256, B2(), 128, 2LL * c.n_e, nullptr, 0, st);
        timed(P_COMBINE_BWD, c.n_e * (8704.0 * 2 + 4608.0), st, [&] {
            launch_combine_gate_bwd(b.y0, b.y1, b.y2, c.n_e, B0(), B1(), B2(), b.y0, b.y1, b.y2, st); });
        mm(b.y0, 640, w.c1m0_t, 768, 640, A0(), 768, c.n_e, nullptr, 0, st);
        mm(b.y1, 512, w.c1m1_t, 512, 512, A1(), 512, 2LL * c.n_e, nullptr, 0, st);
        mm(b.y2, 256, w.c1m2_t, 256, 256, A2(), 256, 2LL * c.n_e, nullptr, 0, st);
        timed(P_GATHER_BWD, c.n_e * (9216.0 + 2 * 6144.0 + 4608.0 + 3 * 144.0 + 4.0) + c.n_nodes * 2 * 4608.0, st, [&] {
            launch_gather_rotate_bwd(n1, row_ptr.i(), src.i(), wig.f(), wRAD.f(), c.e0, c.node0, c.n_nodes, A0(), A1(), A2(),
                                     wRAD.f(), Gbuf.f(), g_n1, g_wig.f(), st); });
        radial_bwd(w.rad, c, st);
    }

    // ------------------------------------------------------------------ full evaluation
    void evaluate(const float* pos, int nimg, double* energy_dev, float* forces_dev, cudaStream_t st) {
        if (!finalized) throw CudaError("umab_finalize_weights has not been called");
        timed(P_GRAPH, (double)nimg * n_atoms * 12.0, st, [&] { build_graph(pos, nimg, st); });
        const int L = cfg.num_layers;
        const size_t ne = (size_t)std::max<long long>(n_edges, 1);
        const size_t nf = (size_t)n_nodes * 9 * C * sizeof(float);
        const bool want_f = forces_dev != nullptr;
        {
            const double need = (double)n_edges * 4096.0 * 4.0 * L;
            double budget = (double)cfg.store_bytes;
            if (cfg.store_bytes == 0) {
                size_t fr = 0, tot = 0;
                UMAB_CUDA(cudaMemGetInfo(&fr, &tot));
                budget = 0.55 * (double)tot;
            }
            store_mode = want_f && cfg.store_bytes >= 0 && need <= budget;
            if (store_mode) {
                ystore.resize(L); zstore.resize(L);
                for (auto& b : ystore) b.ensure(ne * 2176 * 4);
                for (auto& b : zstore) b.ensure(ne * 1920 * 4);
            }
        }
        vec.ensure(ne * 12); dist.ensure(ne * 4); env.ensure(ne * 4); wig.ensure(ne * WIG * 4); gauss.ensure(ne * NB * 4);
        timed(P_GEOMETRY, (double)n_edges * (24.0 + 4.0 * (5 + WIG + NB)), st, [&] {
            launch_geometry_fwd(pos, src.i(), tgt.i(), (int)n_edges, cfg.cutoff, vec.f(), dist.f(), env.f(), wig.f(), gauss.f(), st); });
        xs.resize(L + 1); x1s.resize(L); y1s.resize(L); gps.resize(L);
        for (auto& b : xs) b.ensure(nf);
        for (auto& b : x1s) b.ensure(nf);
        for (auto& b : y1s) b.ensure(nf);
        for (auto& b : gps) b.ensure((size_t)n_nodes * 2 * H * 4);
        nbuf.ensure(nf); abuf.ensure(nf);
        p1.ensure((size_t)n_nodes * H * 4); s1.ensure((size_t)n_nodes * H * 4); p2.ensure((size_t)n_nodes * H * 4);
        node_e.ensure((size_t)n_nodes * 4);

        // ---- embedding + edge-degree embedding
        launch_embed(sphere_emb, csd, zt.i(), n_nodes, xs[0].f(), st);
        for (const Chunk& c : chunks) {
            radial_fwd(ed_rad, c, st);
            launch_rotate_back_reduce(1, wRAD.f(), nullptr, nullptr, row_ptr.i(), wig.f(), env.f(),
                                      1.0f / cfg.edge_degree_rescale, c.e0, c.node0, c.n_nodes, xs[0].f(), xs[0].f(), st);
        }
        save_dbg("x0", xs[0].p, (size_t)n_nodes * 9 * C, st);
        save_dbg("gauss", gauss.p, (size_t)n_edges * NB, st);
        save_dbg("wig", wig.p, (size_t)n_edges * WIG, st);
        save_dbg("env", env.p, (size_t)n_edges, st);

        // ---- layers
        for (int l = 0; l < L; ++l) {
            const LayerW& w = layers[l];
            launch_rms_fwd(xs[l].f(), w.n1w, w.n1b, csd, n_nodes, nbuf.f(), st);
            save_dbg("l" + std::to_string(l) + ".n1", nbuf.p, (size_t)n_nodes * 9 * C, st);
            for (const Chunk& c : chunks) {
                edge_fwd_chunk(w, nbuf.f(), c, l, true, st);
                const YZ b = yz_for(l, c);
                timed(P_ROTBACK, c.n_e * 7828.0 + c.n_nodes * 9216.0, st, [&] {
                    launch_rotate_back_reduce(0, b.z0, b.z1, b.z2, row_ptr.i(), wig.f(), env.f(), 1.0f, c.e0, c.node0,
                                              c.n_nodes, xs[l].f(), x1s[l].f(), st); });
            }
            save_dbg("l" + std::to_string(l) + ".x1", x1s[l].p, (size_t)n_nodes * 9 * C, st);
            launch_rms_fwd(x1s[l].f(), w.n2w, w.n2b, nullptr, n_nodes, nbuf.f(), st);
            mm(nbuf.f(), 9LL * C, w.smlp, 2 * H, C, gps[l].f(), 2 * H, n_nodes, w.smlp_b, 0, st);
            so3_mm(nbuf.f(), C, w.so3_1, H, y1s[l].f(), w.so3_1_b, 0, st);
            launch_ffn_gate_fwd(y1s[l].f(), gps[l].f(), n_nodes, abuf.f(), st);
            UMAB_CUDA(cudaMemcpyAsync(xs[l + 1].p, x1s[l].p, nf, cudaMemcpyDeviceToDevice, st));
            so3_mm(abuf.f(), H, w.so3_2, C, xs[l + 1].f(), w.so3_2_b, 1, st);
            save_dbg("l" + std::to_string(l) + ".x", xs[l + 1].p, (size_t)n_nodes * 9 * C, st);
        }
        // ---- head
        launch_rms_fwd(xs[L].f(), normw, normb, nullptr, n_nodes, nbuf.f(), st);
        mm(nbuf.f(), 9LL * C, h0, H, C, p1.f(), H, n_nodes, h0_b, 0, st);
        launch_eltwise(0, p1.f(), nullptr, (long long)n_nodes * H, s1.f(), st);
        mm(s1.f(), H, h2, H, H, p2.f(), H, n_nodes, h2_b, 0, st);
        if (want_f) gp2.ensure((size_t)n_nodes * H * 4);
        launch_head_final(p2.f(), h4, h4_b, n_nodes, node_e.f(), want_f ? gp2.f() : nullptr, st);
        launch_energy_reduce(node_e.f(), n_img, n_atoms, energy_dev, st);
        save_dbg("node_e", node_e.p, (size_t)n_nodes, st);
        if (!want_f) return;

        // ================= backward: dE_total/dpos
        gx.ensure(nf); gx1.ensure(nf); gn.ensure(nf); ggp.ensure((size_t)n_nodes * 2 * H * 4);
        gs1.ensure((size_t)n_nodes * H * 4);
        Gbuf.ensure(ne * 9 * C * 4);
        g_gauss.ensure(ne * NB * 4); g_env.ensure(ne * 4); g_wig.ensure(ne * WIG * 4); g_vec.ensure(ne * 12);
        UMAB_CUDA(cudaMemsetAsync(g_gauss.p, 0, ne * NB * 4, st));
        UMAB_CUDA(cudaMemsetAsync(g_env.p, 0, ne * 4, st));
        UMAB_CUDA(cudaMemsetAsync(g_wig.p, 0, ne * WIG * 4, st));

        mm(gp2.f(), H, h2_t, H, H, gs1.f(), H, n_nodes, nullptr, 0, st);
        launch_eltwise(1, gs1.f(), p1.f(), (long long)n_nodes * H, gs1.f(), st);      // g_p1
        UMAB_CUDA(cudaMemsetAsync(gn.p, 0, nf, st));
        mm(gs1.f(), H, h0_t, C, H, gn.f(), 9LL * C, n_nodes, nullptr, 0, st);         // g_xf (row 0 only)
        launch_rms_bwd(xs[L].f(), normw, gn.f(), nullptr, n_nodes, gx.f(), st);
        for (int l = L - 1; l >= 0; --l) {
            const LayerW& w = layers[l];
            // FFN adjoint (dL/dy2 = gx)
            so3_mm(gx.f(), C, w.so3_2_t, H, abuf.f(), nullptr, 0, st);                 // g_a
            launch_ffn_gate_bwd(y1s[l].f(), gps[l].f(), abuf.f(), n_nodes, abuf.f(), ggp.f(), st);   // g_y1, g_gp
            so3_mm(abuf.f(), H, w.so3_1_t, C, gn.f(), nullptr, 0, st);                 // g_n2
            mm(ggp.f(), 2 * H, w.smlp_t, C, 2 * H, gn.f(), 9LL * C, n_nodes, nullptr, 1, st);
            launch_rms_bwd(x1s[l].f(), w.n2w, gn.f(), gx.f(), n_nodes, gx1.f(), st);   // g_x1 = gx + norm2^T g_n2
            // Edgewise adjoint
            launch_rms_fwd(xs[l].f(), w.n1w, w.n1b, csd, n_nodes, nbuf.f(), st);       // recompute n1
            for (const Chunk& c : chunks) edge_bwd_chunk(w, nbuf.f(), c, l, gx1.f(), gn.f(), st);
            timed(P_SRC_REDUCE, n_edges * 4612.0 + n_nodes * 9216.0, st, [&] {
                launch_source_reduce(Gbuf.f(), sptr.i(), sedge.i(), n_nodes, gn.f(), st); });
            save_dbg("l" + std::to_string(l) + ".g_n1", gn.p, (size_t)n_nodes * 9 * C, st);
            launch_rms_bwd(xs[l].f(), w.n1w, gn.f(), gx1.f(), n_nodes, gx.f(), st);    // g_x_l
            save_dbg("l" + std::to_string(l) + ".g_x", gx.p, (size_t)n_nodes * 9 * C, st);
        }
        // edge-degree embedding adjoint
        for (const Chunk& c : chunks) {
            radial_fwd(ed_rad, c, st);
            launch_rotate_back_bwd(1, wRAD.f(), nullptr, nullptr, tgt.i(), wig.f(), env.f(),
                                   1.0f / cfg.edge_degree_rescale, c.e0, c.n_e, gx.f(), wRAD.f(), nullptr, nullptr,
                                   g_env.f(), g_wig.f(), st);
            radial_bwd(ed_rad, c, st);
        }
        save_dbg("g_gauss", g_gauss.p, (size_t)n_edges * NB, st);
        save_dbg("g_env", g_env.p, (size_t)n_edges, st);
        save_dbg("g_wig", g_wig.p, (size_t)n_edges * WIG, st);
        launch_geometry_bwd(vec.f(), dist.f(), wig.f(), gauss.f(), g_gauss.f(), g_env.f(), g_wig.f(), (int)n_edges,
                            cfg.cutoff, g_vec.f(), st);
        save_dbg("g_vec", g_vec.p, (size_t)n_edges * 3, st);
        launch_force_reduce(g_vec.f(), row_ptr.i(), sptr.i(), sedge.i(), n_nodes, forces_dev, st);
    }

    ~umab_engine() {
        tc_cache_destroy(tc_cache);
        for (auto& kv : weights) kv.second.buf.release();
        DevBuf* all[] = {&z1, &pos_own, &zt, &deg, &thr, &row_ptr, &src, &tgt, &odeg, &sptr, &cursor, &stmp, &sedge,
                         &vec, &dist, &env, &wig, &gauss, &g_gauss, &g_env, &g_wig, &g_vec, &nbuf, &abuf, &gx, &gx1,
                         &gn, &ggp, &p1, &s1, &p2, &gp2, &gs1, &node_e, &Gbuf, &wA, &wY, &wB, &wZ, &wRAD, &wU1, &wH1,
                         &wU2, &wH2, &e_dev, &f_dev};
        for (DevBuf* b : all) b->release();
        for (auto* v : {&xs, &x1s, &y1s, &gps, &ystore, &zstore}) for (auto& b : *v) b.release();
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

int32_t umab_build_graph(umab_engine* e, const float* pos_dev, int32_t n_images, void* stream) {
    UMAB_TRY
    if (!e || !pos_dev) throw CudaError("null argument");
    DeviceGuard guard(e->cfg.device);
    e->build_graph(pos_dev, n_images, (cudaStream_t)stream);
    UMAB_CATCH
}

int32_t umab_graph_counts(umab_engine* e, int64_t* n_nodes, int64_t* n_edges) {
    UMAB_TRY
    if (!e) throw CudaError("null argument");
    if (n_nodes) *n_nodes = e->n_nodes;
    if (n_edges) *n_edges = e->n_edges;
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
    e->evaluate(pos_dev, n_images, energy_dev, forces_dev, (cudaStream_t)stream);
    UMAB_CATCH
}

int32_t umab_energy_forces_host(umab_engine* e, const float* pos_host, int32_t n_images, double* energy_host,
                                float* forces_host, void* stream) {
    UMAB_TRY
    if (!e || !pos_host || !energy_host) throw CudaError("null argument");
    if (e->n_atoms <= 0) throw CudaError("umab_set_system has not been called");
    DeviceGuard guard(e->cfg.device);
    cudaStream_t st = (cudaStream_t)stream;
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
    }
    if (sizeof(double) * n_images > e->hp_ecap) {
        if (e->hp_e) cudaFreeHost(e->hp_e);
        UMAB_CUDA(cudaMallocHost(&e->hp_e, sizeof(double) * n_images * 2));
        e->hp_ecap = sizeof(double) * n_images * 2;
    }
    memcpy(e->hp_pos, pos_host, nb);
    UMAB_CUDA(cudaMemcpyAsync(e->pos_own.p, e->hp_pos, nb, cudaMemcpyHostToDevice, st));
    e->evaluate(e->pos_own.f(), n_images, e->e_dev.as<double>(), forces_host ? e->f_dev.f() : nullptr, st);
    UMAB_CUDA(cudaMemcpyAsync(e->hp_e, e->e_dev.p, sizeof(double) * n_images, cudaMemcpyDeviceToHost, st));
    if (forces_host) UMAB_CUDA(cudaMemcpyAsync(e->hp_f, e->f_dev.p, nb, cudaMemcpyDeviceToHost, st));
    UMAB_CUDA(cudaStreamSynchronize(st));
    memcpy(energy_host, e->hp_e, sizeof(double) * n_images);
    if (forces_host) memcpy(forces_host, e->hp_f, nb);
    UMAB_CATCH
}

int32_t umab_gemm(int32_t mode, const float* a_dev, const float* w_dev, const float* bias_dev, float* c_dev,
                  int64_t m, int32_t n, int32_t k, void* stream) {
    UMAB_TRY
    GemmArgs g;
    g.A = a_dev; g.lda = k; g.W = w_dev; g.ldw = k; g.Cmat = c_dev; g.ldc = n; g.bias = bias_dev;
    g.M = (int)m; g.N = n; g.K = k;
    if (mode == 1 || mode == 2) {
        if (!gemm_tc_supported(g)) throw CudaError("shape not supported by the tensor-core GEMM");
        // mode 1: weight planes rebuilt on every call (never cached by pointer);
        // mode 2: planes cached by pointer for timing loops (the caller keeps W alive and unchanged)
        static TcPlaneCache* bench_cache = tc_cache_create();
        gemm_tc(g, (cudaStream_t)stream, mode == 2 ? bench_cache : nullptr);
    } else {
        gemm_simt(g, (cudaStream_t)stream);
    }
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
    if (device_bytes) *device_bytes = DevBuf::total;
    UMAB_CATCH
}

}  // extern "C"
