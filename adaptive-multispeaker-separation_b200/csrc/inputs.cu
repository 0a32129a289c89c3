// Input contract of the path on the device (models/network.py:44-88, data/dataset.py:456-468): a batch is
// (x_mix [B,L], x_non_mix [B,S,L], ind [B,S]) with x_mix = sum_s x_non_mix[:, s].  The reference builds the mixture
// (and, under --dataset_normalize, the per-source zero-mean / unit-variance signals) inside its tf.data graph on the host;
// here the host ships only the sources and ONE kernel builds what the graph expects:
//   optional per-source normalisation (x - mean) / sqrt(var) over the row (tf.nn.moments: population variance), in place,
//   then the mixture as the sequential fp32 sum ((x_0 + x_1) + x_2 ...) -- the order np.sum / tf.reduce_sum over the
//   stacked axis use, and the one the linear-mixture analysis kernel verifies bit for bit.
// HBM-bound: 4*S*L bytes read (+ the same written when normalising) + 4*L written per mixture.
#include "common.cuh"
#include <algorithm>

namespace amss {
namespace {

constexpr int MIX_THREADS = 256;
constexpr int MAX_S = 8;

// grid = (chunks, B): without normalisation the kernel is a pure streaming sum
__global__ void mix_sources_kernel(const float* __restrict__ src, int S, int64_t L, float* __restrict__ mix) {
    const int b = blockIdx.y;
    const float* s0 = src + (size_t)b * S * L;
    float* m = mix + (size_t)b * L;
    const int64_t n4 = (L & 3) == 0 ? (L >> 2) : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 a = __ldcs(reinterpret_cast<const float4*>(s0) + i);
        for (int s = 1; s < S; ++s) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(s0 + (size_t)s * L) + i);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        reinterpret_cast<float4*>(m)[i] = a;
    }
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < L; i += stride) {
        float a = s0[i];
        for (int s = 1; s < S; ++s) a += s0[(size_t)s * L + i];
        m[i] = a;
    }
}

// one CTA per source row: mean and population variance (two passes over the row, fixed-order block reductions),
// stats[row] = (mean, var); the row is normalised in place
__global__ void normalize_rows_kernel(float* __restrict__ x, int64_t L, float* __restrict__ stats) {
    __shared__ float red[32];
    float* r = x + (size_t)blockIdx.x * L;
    float a = 0.f;
    for (int64_t i = threadIdx.x; i < L; i += blockDim.x) a += r[i];
    const float mean = block_sum(a, red) / (float)L;
    float q = 0.f;
    for (int64_t i = threadIdx.x; i < L; i += blockDim.x) { const float d = r[i] - mean; q = fmaf(d, d, q); }
    const float var = block_sum(q, red) / (float)L;
    const float inv = 1.f / sqrtf(var);
    for (int64_t i = threadIdx.x; i < L; i += blockDim.x) r[i] = (r[i] - mean) * inv;
    if (threadIdx.x == 0 && stats) { stats[2 * blockIdx.x] = mean; stats[2 * blockIdx.x + 1] = var; }
}

}  // namespace
}  // namespace amss

using namespace amss;

extern "C" int amss_prepare_inputs(float* x_non_mix, int B, int S, int64_t L, int normalize, float* stats, float* x_mix,
                                   void* stream) {
    AMSS_REQUIRE(x_non_mix && x_mix && B > 0 && S > 0 && S <= MAX_S && L > 0, "prepare_inputs: bad arguments");
    AMSS_REQUIRE((((uintptr_t)x_non_mix | (uintptr_t)x_mix) & 15) == 0, "prepare_inputs: buffers must be 16-byte aligned");
    if (normalize) AMSS_LAUNCH(normalize_rows_kernel, B * S, 512, 0, stream, x_non_mix, L, stats);
    const int chunks = (int)std::min<int64_t>((L / 4 + MIX_THREADS - 1) / MIX_THREADS + 1, 64);
    AMSS_LAUNCH(mix_sources_kernel, dim3(chunks, B), MIX_THREADS, 0, stream, x_non_mix, S, L, x_mix);
    return AMSS_OK;
}
