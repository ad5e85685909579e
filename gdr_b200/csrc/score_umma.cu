// tcgen05 grouped-GEMM scoring path — placeholder until the TMA/UMMA kernel lands.
#include "gdr_common.cuh"
namespace gdr {
bool umma_make_tensor_map(CUtensorMap *, const void *, int64_t, int) { return false; }
cudaError_t launch_qsplit(const ScoreArgs &, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_score_umma(const ScoreArgs &, const CUtensorMap *, cudaStream_t, int) { return cudaErrorNotSupported; }
}  // namespace gdr
