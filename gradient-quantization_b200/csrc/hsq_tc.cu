// hsq_tc.cu -- tcgen05 (TF32) nearest-codeword search with fp32 rescoring.
// Placeholder until the tensor-core kernel lands: reports "unsupported" so that
// GQ_ALGO_AUTO uses the exact CUDA-core kernel.
#include "gq_internal.cuh"

namespace gq {
bool hsq_tc_supported(int, int, int) { return false; }
size_t hsq_tc_workspace_bytes(int64_t) { return 0; }
int hsq_search_tc(const float *, int64_t, int, const float *, int, void *, int, float *, const int64_t *,
                  int, uint32_t *, void *, size_t, cudaStream_t)
{
    set_error("tcgen05 search not built");
    return GQ_ERR_UNSUPPORTED;
}
}  // namespace gq
