// Shared helpers for the libamss_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include "amss.h"

namespace amss {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);

#define AMSS_REQUIRE(cond, ...)                     \
    do {                                            \
        if (!(cond)) {                              \
            amss::set_error(__VA_ARGS__);           \
            return AMSS_ERR_INVALID_ARG;            \
        }                                           \
    } while (0)

#define AMSS_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) {                                              \
            amss::set_error("%s failed: %s", #call, cudaGetErrorString(e__)); \
            return AMSS_ERR_CUDA;                                              \
        }                                                                      \
    } while (0)

// Launch + count + error check.  Usage: AMSS_LAUNCH(kernel, grid, block, smem, stream, args...)
#define AMSS_LAUNCH(kern, grid, block, smem, stream, ...)                 \
    do {                                                                  \
        kern<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__); \
        amss::count_launch();                                             \
        int rc__ = amss::check_launch(#kern);                             \
        if (rc__ != AMSS_OK) return rc__;                                 \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Block-wide sum in a fixed (deterministic) order. `red` needs >= 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = 0.f;
    if (w == 0) {
        r = lane < nw ? red[lane] : 0.f;
        r = warp_sum(r);
        if (lane == 0) red[0] = r;
    }
    __syncthreads();
    r = red[0];
    return r;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

}  // namespace amss
