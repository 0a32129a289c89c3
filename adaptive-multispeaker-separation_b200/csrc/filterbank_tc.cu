// tcgen05 Toeplitz implicit-GEMM analysis filterbank (placeholder, see gemm_tc.cu).
#include "common.cuh"
namespace amss {
bool filterbank_analysis_tc_supported(int, int, int, int, int, int) { return false; }
size_t filterbank_analysis_tc_workspace(int, int, int, int, int, int, int) { return 256; }
int filterbank_analysis_tc(const float*, const float*, int, int, int, int, int, int, int, float*, int64_t*, void*,
                           size_t, cudaStream_t) {
    set_error("filterbank_analysis_tc: not built");
    return AMSS_ERR_UNSUPPORTED;
}
}  // namespace amss
