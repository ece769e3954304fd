/* umab -- C ABI of the B200-native UMA (eSCN-MD + MoLE) energy / force engine.
 *
 * This is the drop-in boundary for pdb2reaction's one data-parallel hot path.  The reference
 * is pure Python and has no FFI of its own; the interface each entry point replaces is the
 * fairchem call the reference's calculator makes (file:line into /root/reference):
 *
 *   umab_create / umab_set_weight / umab_finalize_weights
 *        <- pretrained_mlip.get_predict_unit(model, device=...)        pdb2reaction/uma_pysis.py:246-250
 *           (+ the MoLE merge fairchem does per system; done on the host once per calculator)
 *   umab_set_system
 *        <- UMAcore.__init__ latching elem / charge / spin / task      pdb2reaction/uma_pysis.py:266-269
 *   umab_build_graph (+ umab_graph_counts / umab_graph_copy)
 *        <- AtomicData.from_ase + data_list_collater(otf_graph=True)    pdb2reaction/uma_pysis.py:313-322
 *   umab_energy_forces / umab_energy_forces_host
 *        <- self.predict.predict(batch) (energy + conservative forces)  pdb2reaction/uma_pysis.py:385-392
 *           called once per image by the reference; here once per BATCH of images
 *           (GSM/DMF string, FD-Hessian displacements: uma_pysis.py:652-675)
 *
 * Conventions: plain pointers and sizes only; device pointers are owned by the caller (e.g.
 * PyTorch tensors), `stream` is a cudaStream_t passed as void*; every function returns 0 on
 * success, non-zero on failure with the message available from umab_last_error().  All
 * images of one engine share one composition (atomic numbers), charge, spin and task, as one
 * reference calculator instance does.  No CPU fallback exists: without a CUDA device every
 * compute entry point fails.
 */
#ifndef UMAB_H
#define UMAB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define UMAB_API __attribute__((visibility("default")))
#else
#define UMAB_API
#endif

typedef struct umab_engine umab_engine;

typedef struct umab_config {
    int32_t sphere_channels;    /* 128 (compiled in) */
    int32_t hidden_channels;    /* 128 (compiled in) */
    int32_t num_distance_basis; /* 64  (compiled in) */
    int32_t num_layers;         /* 4 */
    int32_t max_neighbors;      /* 300 */
    int32_t device;             /* CUDA ordinal */
    int32_t debug;              /* 1: keep named intermediates for umab_debug_tensor */
    int32_t gemm_mode;          /* 0: fp32 SIMT, 1: tcgen05 bf16x3, 2: auto (tensor cores for images >= 100 atoms) */
    float cutoff;               /* 6.0 Angstrom */
    float edge_degree_rescale;  /* 5.0 */
    int64_t workspace_bytes;    /* per-chunk edge workspace budget; 0 = default */
    int64_t store_bytes;        /* HBM budget for keeping the conv outputs of all layers for the backward
                                   (16.4 KB per edge and layer) instead of recomputing them: 0 = auto
                                   (45 % of the device memory), -1 = never (always recompute) */
} umab_config;

/* ABI version of this header; umab_abi_version() must return the same value. */
#define UMAB_ABI_VERSION 7
/* return code of umab_last_call: the sync-free graph build of the last call overflowed its edge capacity */
#define UMAB_RETRY 2

UMAB_API int32_t umab_abi_version(void);
UMAB_API const char* umab_last_error(void);

UMAB_API int32_t umab_create(const umab_config* cfg, umab_engine** out);
UMAB_API void umab_destroy(umab_engine* e);

/* Copy one named fp32 parameter (host memory) into the engine. */
UMAB_API int32_t umab_set_weight(umab_engine* e, const char* name, const float* host, size_t numel);
/* Verify that every parameter the kernels need is present with the expected size. */
UMAB_API int32_t umab_finalize_weights(umab_engine* e);

/* Run-time options.  "neighbor_mode": 0 = auto (shared-memory cell list from 128 atoms per image, brute
 * force below), 1 = brute force, 2 = cell list; both searches return the identical edge list.
 * "nosync" (default 1): sync-free graph build for calls that fit one chunk of the edge workspace;
 * "cuda_graphs" (default 1): capture / replay of launch-bound umab_energy_forces_host calls;
 * "simt_round_fwd" / "simt_round_bwd": operand rounding of the precision study (fp32 SIMT GEMMs only);
 * "jvp_shared_base" (default 0): the images of the following umab_forces_jvp calls all sit at ONE geometry (the columns
 * of an analytic Hessian, reference uma_pysis.py:394-415): verified on the device per call, then the value-plane GEMMs
 * run on one image and the block is copied to the others -- identical result bits, counter "dedupe_gemms". */
UMAB_API int32_t umab_set_option(umab_engine* e, const char* name, int64_t value);
/* Read an option back, or a counter: "graph_replays", "graph_captures", "overflow_retries",
 * "edges_per_image_seen". */
UMAB_API int32_t umab_get_option(umab_engine* e, const char* name, int64_t* value);

/* Atomic numbers of ONE image (host, n_atoms ints); all images share them. */
UMAB_API int32_t umab_set_system(umab_engine* e, const int32_t* z_host, int32_t n_atoms);

/* Neighbour search only: pos_dev [n_images, n_atoms, 3] fp32 on the device. */
UMAB_API int32_t umab_build_graph(umab_engine* e, const float* pos_dev, int32_t n_images, void* stream);
UMAB_API int32_t umab_graph_counts(umab_engine* e, int64_t* n_nodes, int64_t* n_edges);
/* Copy the canonical (target, source)-sorted edge list out; either pointer may be NULL. */
UMAB_API int32_t umab_graph_copy(umab_engine* e, int32_t* src_dev, int32_t* tgt_dev, int32_t* row_ptr_dev, void* stream);

/* Energies [n_images] (double, eV) and forces [n_images, n_atoms, 3] (fp32, eV/A; NULL = energy
 * only, skips the backward pass).  Device pointers.  Rebuilds the graph on every call, as the
 * reference does.  Batches larger than the per-layer stores allow run as consecutive sub-batches inside
 * the call.  Calls that fit ONE chunk of the edge workspace (a GSM string of 300-atom images, a 500-atom
 * Hessian displacement batch, small molecules) do NOT synchronise the stream: the edge arrays are sized
 * from a capacity (complete graph below 128 atoms per image, else the largest edges-per-image seen + 3 %)
 * and an overflow is flagged on the device -- collect it with umab_last_call() after the stream has been
 * synchronised and repeat the call on UMAB_RETRY.  Larger calls read the edge count back once. */
UMAB_API int32_t umab_energy_forces(umab_engine* e, const float* pos_dev, int32_t n_images,
                           double* energy_dev, float* forces_dev, void* stream);

/* Images, edges (summed over the sub-batches) and sub-batches of the last umab_energy_forces* /
 * umab_forces_jvp call.  Waits for the call's status words; returns UMAB_RETRY when its sync-free graph
 * build overflowed (results invalid: repeat the call), 0 otherwise. */
UMAB_API int32_t umab_last_call(umab_engine* e, int64_t* n_images, int64_t* n_edges, int64_t* n_subcalls);

/* Same with HOST buffers (pinned or pageable): H2D, compute, D2H and a final sync inside (overflows of the
 * sync-free path are handled inside).  Launch-bound calls are captured into a CUDA graph on their second
 * occurrence and replayed as one graph launch afterwards (UMAB_CUDA_GRAPHS=0 disables). */
UMAB_API int32_t umab_energy_forces_host(umab_engine* e, const float* pos_host, int32_t n_images,
                                double* energy_host, float* forces_host, void* stream);

/* Analytic Hessian columns (replaces torch.autograd.functional.hessian, pdb2reaction/uma_pysis.py:402-409):
 * every image carries a displacement direction tangent_dev [n_images, n_atoms, 3]; the forward and the
 * hand-written backward run on dual numbers and return dforces_dev = d(forces)/d(eps) = -H.tangent
 * (fp32, eV/A^2 for unit tangents in A).  energy_dev / forces_dev may be NULL.  Device pointers. */
UMAB_API int32_t umab_forces_jvp(umab_engine* e, const float* pos_dev, const float* tangent_dev, int32_t n_images,
                                 double* energy_dev, float* forces_dev, float* dforces_dev, void* stream);

/* Standalone GEMM  C[M,N] = A[M,K] . W[N,K]^T (+bias), device pointers, for kernel unit tests
 * and roofline measurements.  mode: 0 SIMT fp32, 1 tensor-core bf16x3 (activation split in the kernel),
 * 2 same with the bf16 weight planes cached by pointer (timing loops only), 3 / 4 TMA-fed tensor-core
 * bf16x3 kernel (activation pre-split into bf16 hi/lo planes, as the pipeline's producing kernels write
 * it) with a 64 / 32 wide k block, 5 / 6 the CTA-pair (cta_group::2, 256-row tiles) form of the same kernel
 * with a 64 / 32 wide k block. */
UMAB_API int32_t umab_gemm(int32_t mode, const float* a_dev, const float* w_dev, const float* bias_dev, float* c_dev,
                  int64_t m, int32_t n, int32_t k, void* stream);

/* Kernel-only timing of one GEMM shape: `iters` back-to-back launches between two CUDA events on
 * `stream` after one warm-up launch (weight planes, tensor maps and the activation split are prepared
 * outside the timed region).  Modes as umab_gemm.  ms_per_iter: host pointer. */
UMAB_API int32_t umab_gemm_bench(int32_t mode, const float* a_dev, const float* w_dev, float* c_dev, int64_t m, int32_t n,
                        int32_t k, int32_t iters, double* ms_per_iter, void* stream);

/* Debug access (config.debug = 1): device pointer + element count of a named intermediate of
 * the last umab_energy_forces call.  The pointer stays valid until the next call. */
UMAB_API int32_t umab_debug_tensor(umab_engine* e, const char* name, const float** ptr_dev, size_t* numel);

/* Per-kernel-family timing with CUDA events on the launching stream (bench.py's roofline):
 * enable (resets the counters) / disable; read the accumulated device milliseconds, launch count
 * algorithmic work (FLOPs for "gemm", bytes otherwise) and algorithmic HBM bytes of family `cat`;
 * umab_profile_name(cat) is NULL past the last family. */
UMAB_API int32_t umab_profile(umab_engine* e, int32_t enable);
UMAB_API int32_t umab_profile_read(umab_engine* e, int32_t cat, double* ms, int64_t* launches, double* work,
                                   double* bytes);
UMAB_API const char* umab_profile_name(int32_t cat);

/* ---- Hessian assembly / vibrational pre-processing on the device (no engine handle; current CUDA device).
 *
 * umab_hessian_fd_columns: forces_dev [2 * n_cols, dof] fp32 holds F(x + h e_k), F(x - h e_k) for k = dof_idx_dev[q];
 * writes H[:, k] = -(F+ - F-) / (2 h_step) into hessian_dev [dof, ld] (fp64 when is_f64, else fp32).  Replaces the
 * column loop of uma_pysis._build_fd_hessian_gpu (pdb2reaction/uma_pysis.py:652-675).
 *
 * umab_hessian_mw_project: in place, fp64:  H <- sym(P (S H S) P) with S = diag(inv_sqrt_m_dev [n]) and
 * P = I - Q Q^T, q_dev [n, r] row-major orthonormal translation/rotation basis, r <= 6 (r = 0: mass-weighting and
 * symmetrisation only).  Replaces freq._mw_projected_hessian (pdb2reaction/freq.py:159-205) and the PHVA
 * variants (freq.py:290-335).  workspace_dev: at least umab_hessian_mw_workspace(n, r) doubles. */
UMAB_API int32_t umab_hessian_fd_columns(const float* forces_dev, const int32_t* dof_idx_dev, int32_t n_cols, int32_t dof,
                                         double h_step, void* hessian_dev, int64_t ld, int32_t is_f64, void* stream);
UMAB_API int64_t umab_hessian_mw_workspace(int32_t n, int32_t r);
UMAB_API int32_t umab_hessian_mw_project(double* hessian_dev, int32_t n, const double* inv_sqrt_m_dev, const double* q_dev,
                                         int32_t r, double* workspace_dev, int64_t workspace_doubles, void* stream);

/* Free the engine's per-call device memory (edge workspace, per-layer stores, per-edge / per-node state, captured
 * graphs); weights stay, everything re-grows on the next call.  For hosts that keep several engines (compositions)
 * alive on one GPU: the reference re-creates calculators per composition, uma_pysis.py:502-504. */
UMAB_API int32_t umab_release_workspace(umab_engine* e);

/* Counters since creation: kernel launches issued by this library and bytes allocated. */
UMAB_API int32_t umab_stats(umab_engine* e, int64_t* kernel_launches, int64_t* device_bytes);

#ifdef __cplusplus
}
#endif
#endif /* UMAB_H */
