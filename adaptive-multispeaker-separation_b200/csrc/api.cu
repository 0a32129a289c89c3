// Library-level entry points: version, error string, launch counter.
#include "common.cuh"
#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace amss {
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return AMSS_ERR_CUDA;
    }
    return AMSS_OK;
}
}  // namespace amss

extern "C" {
int amss_version(void) { return 100; }
const char* amss_last_error(void) { return amss::g_err; }
uint64_t amss_launch_count(void) { return amss::g_launches.load(); }
}
