// placeholder until the tcgen05 path lands: reports "unsupported" so the SIMT kernel runs
#include "common.cuh"
namespace umab {
bool gemm_tc_supported(const GemmArgs&) { return false; }
void gemm_tc(const GemmArgs&, cudaStream_t) { throw CudaError("tensor-core GEMM not built"); }
}  // namespace umab
