// tcgen05 bf16 GEMM (placeholder until the TMEM/TMA kernel lands: reports "unsupported", the
// caller turns that into a loud AMSS_ERR_UNSUPPORTED -- never a silent fallback).
#include "common.cuh"
namespace amss {
bool gemm_tc_supported(int, int, int, int, int, int, int, int) { return false; }
size_t gemm_tc_workspace(int, int, int, int, int, int) { return 256; }
int gemm_tc(const float*, int, const float*, int, const float*, int, int, int, int, int, int, int, float*, int, int,
            int, void*, size_t, cudaStream_t) {
    set_error("gemm_tc: not built");
    return AMSS_ERR_UNSUPPORTED;
}
}  // namespace amss
