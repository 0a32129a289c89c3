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

// Separator input options, one CTA per mixture row (256 KB per row: three short passes, L2-resident).
__device__ __forceinline__ float prep_point(float x, int abs_input, int pre_func) {
    if (abs_input) x = fabsf(x);
    if (pre_func == 1) x = sqrtf(x);
    else if (pre_func == 2) x = logf(x + 1e-12f) / logf(10.f);           // utils/ops.py:56-59 log10
    return x;
}
__device__ __forceinline__ float block_minmax(float v, float* red, bool is_max) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const float u = __shfl_xor_sync(0xffffffffu, v, o); v = is_max ? fmaxf(v, u) : fminf(v, u); }
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < nw; ++i) r = is_max ? fmaxf(r, red[i]) : fminf(r, red[i]);
    __syncthreads();
    return r;
}
__global__ void separator_input_prep_kernel(const float* __restrict__ X, int64_t TF, int abs_input, int pre_func,
                                            int normalize, float silence_db, float* __restrict__ out) {
    __shared__ float red[32];
    const float* x = X + (size_t)blockIdx.x * TF;
    float* o = out + (size_t)blockIdx.x * TF;
    float a = 0.f, b = 1.f;                                  // normalised value = (v - a) * b
    float mx = -INFINITY;
    if (normalize == 1 || silence_db > 0.f) {
        float lo = INFINITY, hi = -INFINITY;
        for (int64_t i = threadIdx.x; i < TF; i += blockDim.x) { const float v = prep_point(x[i], abs_input, pre_func); lo = fminf(lo, v); hi = fmaxf(hi, v); }
        lo = block_minmax(lo, red, false);
        hi = block_minmax(hi, red, true);
        if (normalize == 1) { a = lo; b = 1.f / (hi - lo); mx = 1.f; }
        else mx = hi;
    }
    if (normalize == 2) {
        float sum = 0.f;
        for (int64_t i = threadIdx.x; i < TF; i += blockDim.x) sum += prep_point(x[i], abs_input, pre_func);
        const float mean = block_sum(sum, red) / (float)TF;
        float q = 0.f;
        for (int64_t i = threadIdx.x; i < TF; i += blockDim.x) { const float d = prep_point(x[i], abs_input, pre_func) - mean; q = fmaf(d, d, q); }
        const float var = block_sum(q, red) / (float)TF;
        a = mean; b = 1.f / sqrtf(var);
        if (silence_db > 0.f) mx = (mx - a) * b;
    }
    const float thr = silence_db / 20.f;
    for (int64_t i = threadIdx.x; i < TF; i += blockDim.x) {
        float v = (prep_point(x[i], abs_input, pre_func) - a) * b;
        if (silence_db > 0.f && !((mx - v) < thr)) v = 0.f;
        o[i] = v;
    }
}
__global__ void label_weights_kernel(const float* __restrict__ X, int64_t TF, int function_mask, float silence_threshold,
                                     float* __restrict__ w) {
    __shared__ float red[32];
    const float* x = X + (size_t)blockIdx.x * TF;
    float hi = 0.f;
    for (int64_t i = threadIdx.x; i < TF; i += blockDim.x) hi = fmaxf(hi, fabsf(x[i]));
    hi = block_minmax(hi, red, true);
    for (int64_t i = threadIdx.x; i < TF; i += blockDim.x) {
        const float v = fabsf(x[i]);
        float m = 1.f;
        const float r = v / hi;
        if (function_mask == 1) m = r;
        else if (function_mask == 2) m = sqrtf(r);
        else if (function_mask == 3) m = r * r;
        if (silence_threshold > 0.f) m *= (logf(hi / v) / logf(10.f) < silence_threshold) ? 1.f : 0.f;
        w[(size_t)blockIdx.x * TF + i] = m;
    }
}

}  // namespace
}  // namespace amss

using namespace amss;

extern "C" int amss_separator_input_prep(const float* X, int B, int64_t TF, int abs_input, int pre_func, int normalize,
                                         float silence_db, float* out, void* stream) {
    AMSS_REQUIRE(X && out && B > 0 && TF > 0, "separator_input_prep: bad arguments");
    AMSS_REQUIRE(pre_func >= 0 && pre_func <= 2 && normalize >= 0 && normalize <= 2, "separator_input_prep: unknown mode");
    AMSS_LAUNCH(separator_input_prep_kernel, B, 1024, 0, stream, X, TF, abs_input, pre_func, normalize, silence_db, out);
    return AMSS_OK;
}

extern "C" int amss_label_weights(const float* X, int B, int64_t TF, int function_mask, float silence_threshold, float* w,
                                  void* stream) {
    AMSS_REQUIRE(X && w && B > 0 && TF > 0 && function_mask >= 0 && function_mask <= 3, "label_weights: bad arguments");
    AMSS_LAUNCH(label_weights_kernel, B, 1024, 0, stream, X, TF, function_mask, silence_threshold, w);
    return AMSS_OK;
}

extern "C" int amss_prepare_inputs(float* x_non_mix, int B, int S, int64_t L, int normalize, float* stats, float* x_mix,
                                   void* stream) {
    AMSS_REQUIRE(x_non_mix && x_mix && B > 0 && S > 0 && S <= MAX_S && L > 0, "prepare_inputs: bad arguments");
    AMSS_REQUIRE((((uintptr_t)x_non_mix | (uintptr_t)x_mix) & 15) == 0, "prepare_inputs: buffers must be 16-byte aligned");
    if (normalize) AMSS_LAUNCH(normalize_rows_kernel, B * S, 512, 0, stream, x_non_mix, L, stats);
    const int chunks = (int)std::min<int64_t>((L / 4 + MIX_THREADS - 1) / MIX_THREADS + 1, 64);
    AMSS_LAUNCH(mix_sources_kernel, dim3(chunks, B), MIX_THREADS, 0, stream, x_non_mix, S, L, x_mix);
    return AMSS_OK;
}

// ---- box sums: the average-pool front / back end written on the sparse kernels ---------------------------------------------
// tf.layers.average_pooling2d over the stride-1 convolution (models/adapt.py:118-120) is the strided response of the
// BOX-FILTERED signal, and UpSampling2D + conv2d_transpose (:224-243) is the sparse overlap-add with BOX-FILTERED filters, so
// both directions (and all their gradients) run on the kernels of the max-pool path once the operand has been summed over a
// sliding window of P samples:  out[s][u] = scale * sum_{j<P} in[s][u + dir*j]   (in = 0 outside [0, len_in)).
// series s start at in + s*ss / out + s*oss, consecutive elements are es / oes apart (rows of a signal batch: ss = L, es = 1;
// columns of a filter bank [W][N]: ss = 1, es = N).
namespace amss {
namespace {
__global__ void box_sum_kernel(const float* __restrict__ in, int nser, int64_t len_in, int64_t ss, int64_t es, int P, int dir,
                               float scale, int64_t len_out, int64_t oss, int64_t oes, float* __restrict__ out) {
    const int64_t total = (int64_t)nser * len_out;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // consecutive threads take consecutive addresses of the output (and of the input)
        const int64_t s = oes == 1 ? i / len_out : i % nser, u = oes == 1 ? i % len_out : i / nser;
        const float* src = in + s * ss;
        float acc = 0.f;
        for (int j = 0; j < P; ++j) {                       // fixed order: deterministic
            const int64_t v = u + (int64_t)dir * j;
            if (v >= 0 && v < len_in) acc += src[v * es];
        }
        out[s * oss + u * oes] = acc * scale;
    }
}
}  // namespace
}  // namespace amss

extern "C" int amss_box_sum(const float* in, int nser, int64_t len_in, int64_t series_stride, int64_t elem_stride, int P, int dir,
                            float scale, int64_t len_out, int64_t out_series_stride, int64_t out_elem_stride, float* out,
                            void* stream) {
    AMSS_REQUIRE(in && out && nser > 0 && len_in > 0 && len_out > 0 && P > 0 && (dir == 1 || dir == -1), "box_sum: bad arguments");
    const int64_t total = (int64_t)nser * len_out;
    AMSS_LAUNCH(amss::box_sum_kernel, (int)std::min<int64_t>((total + 255) / 256, 32 * amss::kNumSMs), 256, 0, stream, in, nser, len_in,
                series_stride, elem_stride, P, dir, scale, len_out, out_series_stride, out_elem_stride, out);
    return AMSS_OK;
}

