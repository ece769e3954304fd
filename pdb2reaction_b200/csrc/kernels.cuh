// Launchers of every kernel of the pipeline.  `_t<S>` launchers are instantiated for S = float
// (energy / forces) and S = D1 (value + tangent: analytic Hessian columns), see dual.cuh.
#pragma once
#include "dual.cuh"

namespace umab {

// ---- neighbors.cu
void launch_neighbor_count(const float* pos, int n_img, int n_atoms, float cutoff, int cap, int* deg, float* thr, cudaStream_t st);
void launch_neighbor_fill(const float* pos, int n_img, int n_atoms, float cutoff, int cap, const float* thr,
                          const int* row_ptr, int* src, int* tgt, int e_cap, cudaStream_t st);
void launch_scan(const int* in, int* out, int n, cudaStream_t st);
// shared-memory cell-list search (same edge list as the brute-force kernels)
size_t cell_grid_bytes();
void launch_cell_list(const float* pos, int n_img, int n_atoms, float cutoff, int cap_cells, void* grid,
                      int* cell_count, int* atom_cell, int* cell_start, int* cell_atoms, cudaStream_t st);
void launch_neighbor_cell_count(const float* pos, int n_img, int n_atoms, float cutoff, int cap, int cap_cells,
                                const void* grid, const int* cell_start, const int* cell_atoms, int* deg, float* thr,
                                cudaStream_t st);
void launch_neighbor_cell_fill(const float* pos, int n_img, int n_atoms, float cutoff, int cap, int cap_cells,
                               const void* grid, const int* cell_start, const int* cell_atoms, const float* thr,
                               const int* row_ptr, int* src, int* tgt, int e_cap, cudaStream_t st);
void launch_fill_int(int* p, int n, int v, cudaStream_t st);
void launch_edge_status(int* row_ptr, int n_nodes, int e_cap, int* status_dev, cudaStream_t st);
void launch_source_csr(const int* src, int n_edges, const int* n_edges_dev, int n_nodes, int* odeg, int* sptr, int* cursor,
                       int* tmp, int* sedge, cudaStream_t st);

// ---- geometry.cu
template <class S>
void launch_geometry_fwd_t(GP<S> pos, const int* src, const int* tgt, int n_edges, float cutoff, GP<S> vec,
                           GP<S> dist, GP<S> env, GP<S> wig, GP<S> gauss, cudaStream_t st);
template <class S>
void launch_geometry_bwd_t(GP<S> vec, GP<S> dist, GP<S> wig, GP<S> gauss, GP<S> g_gauss, GP<S> g_env, GP<S> g_wig,
                           int n_edges, float cutoff, GP<S> g_vec, cudaStream_t st);
void launch_force_reduce(const float* g_vec, const int* row_ptr, const int* sptr, const int* sedge, int n_nodes,
                         float* forces, cudaStream_t st);

// ---- radial.cu
template <class S>
void launch_ln_silu_fwd_t(GP<S> u, AP<S> h, const float* gamma, const float* beta, const float* bias,
                          const float* t_src, const float* t_tgt, const int* z, const int* src, const int* tgt,
                          int rows, cudaStream_t st, ImgShare sh = ImgShare{0, 0});
template <class S>
void launch_ln_silu_bwd_t(GP<S> u, GP<S> g, AP<S> out, const float* gamma, const float* beta, int rows, cudaStream_t st,
                          ImgShare sh = ImgShare{0, 0}, bool share_u = true);

// ---- edge_ops.cu
template <class S>
void launch_gather_rotate_scale_t(GP<S> x, const int* src, const int* tgt, GP<S> wig, GP<S> rad, long long e0, int n_e,
                                  AP<S> A0, AP<S> A1, AP<S> A2, cudaStream_t st, ImgShare sh = ImgShare{0, 0});
template <class S>
void launch_gather_rotate_bwd_t(GP<S> x, const int* row_ptr, const int* src, GP<S> wig, GP<S> rad, long long e0,
                                int node0, int n_nodes, GP<S> gA0, GP<S> gA1, GP<S> gA2, AP<S> g_rad, GP<S> G,
                                GP<S> g_x, GP<S> g_wig, cudaStream_t st);
template <class S>
void launch_gather_rotate_bwd_closed_t(GP<S> x, const int* row_ptr, const int* sptr, const int* sedge, GP<S> wig, GP<S> rad,
                                       long long e0, int node0, int n_nodes, GP<S> gA0, GP<S> gA1, GP<S> gA2, AP<S> g_rad,
                                       GP<S> g_x, GP<S> g_wig, cudaStream_t st, ImgShare sh = ImgShare{0, 0}, int e_img = 0);
void launch_source_reduce(const float* G, const int* sptr, const int* sedge, int n_nodes, float* g_x, cudaStream_t st);
template <class S>
void launch_combine_gate_fwd_t(GP<S> Y0, GP<S> Y1, GP<S> Y2, int n_e, AP<S> B0, AP<S> B1, AP<S> B2, cudaStream_t st,
                               ImgShare sh = ImgShare{0, 0});
// fused-gate path: m = 0 part of the combine + sigmoid(gates) [n_e, 256] for the GEMM epilogues
void launch_gate_b0(GP<float> Y0, int n_e, AP<float> B0, float* sg, cudaStream_t st);
template <class S>
void launch_combine_gate_bwd_t(GP<S> Y0, GP<S> Y1, GP<S> Y2, int n_e, GP<S> gB0, GP<S> gB1, GP<S> gB2, AP<S> gY0,
                               AP<S> gY1, AP<S> gY2, cudaStream_t st, ImgShare sh = ImgShare{0, 0});
template <class S>
void launch_rotate_back_reduce_t(int mode, GP<S> Z0, GP<S> Z1, GP<S> Z2, const int* row_ptr, GP<S> wig, GP<S> env,
                                 float scale, long long e0, int node0, int n_nodes, GP<S> base, GP<S> out,
                                 cudaStream_t st, ImgShare sh = ImgShare{0, 0}, int e_img = 0);
template <class S>
void launch_rotate_back_bwd_t(int mode, GP<S> Z0, GP<S> Z1, GP<S> Z2, const int* tgt, GP<S> wig, GP<S> env, float scale,
                              long long e0, int n_e, GP<S> g_out, AP<S> gZ0, AP<S> gZ1, AP<S> gZ2, GP<S> g_env,
                              GP<S> g_wig, cudaStream_t st, ImgShare sh = ImgShare{0, 0});

// ---- node_ops.cu
void launch_embed(const float* sphere_emb, const float* csd, const int* z, int n_nodes, float* x, cudaStream_t st);
template <class S>
void launch_rms_fwd_t(GP<S> x, const float* w_aff, const float* b_aff, const float* add0, int n_nodes, GP<S> y,
                      cudaStream_t st);
template <class S>
void launch_rms_bwd_t(GP<S> x, const float* w_aff, GP<S> g_y, GP<S> g_add, int n_nodes, GP<S> g_x, cudaStream_t st);
template <class S> void launch_ffn_gate_fwd_t(GP<S> y1, GP<S> gp, int n_nodes, GP<S> a, cudaStream_t st);
template <class S>
void launch_ffn_gate_bwd_t(GP<S> y1, GP<S> gp, GP<S> g_a, int n_nodes, GP<S> g_y1, GP<S> g_gp, cudaStream_t st);
template <class S> void launch_eltwise_t(int mode, GP<S> a, GP<S> b, long long n, GP<S> out, cudaStream_t st);
template <class S>
void launch_head_final_t(GP<S> p2, const float* w4, const float* b4, int n_nodes, float* node_e, GP<S> g_p2,
                         cudaStream_t st);
void launch_energy_reduce(const float* node_e, int n_img, int n_atoms, double* energy, cudaStream_t st);
void launch_tile_int(const int* in, int n, int reps, int* out, cudaStream_t st);
// Hessian columns of ONE base geometry (all images of a dual-number batch share their value planes):
void launch_replicate_block(float* base, long long block_floats, int reps, cudaStream_t st);   // block 0 -> blocks 1 .. reps-1
void launch_same_images(const float* pos, long long n3, int n_img, int* flag, cudaStream_t st); // flag = 1 unless all images equal image 0

}  // namespace umab
