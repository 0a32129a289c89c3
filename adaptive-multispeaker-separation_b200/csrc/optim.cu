// AMSGrad (utils/ops.py:639-704) as one fused pass over the flat parameter buffer, global-norm
// clipping (models/network.py:191-192) and the |window|*bases filter construction of the
// adaptive front/back end (models/adapt.py:106, :234) with its backward.
#include "common.cuh"
#include <algorithm>

namespace amss {
namespace {

// 9 array passes (read p,g,m,v,vhat; write p,m,v,vhat) = 36 B per parameter: HBM-bound.
__global__ void amsgrad_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                               float* __restrict__ v, float* __restrict__ vhat, int64_t n, float lr_t, float b1,
                               float b2, float eps, float gscale, const float* __restrict__ gscale_dev) {
    if (gscale_dev) gscale *= gscale_dev[0];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = n >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        float4 hh = reinterpret_cast<float4*>(vhat)[i];
#define AMS1(c)                                                     \
        {                                                           \
            const float gi = gg.c * gscale;                         \
            mm.c = b1 * mm.c + (1.f - b1) * gi;                     \
            vv.c = b2 * vv.c + (1.f - b2) * gi * gi;                \
            hh.c = fmaxf(hh.c, vv.c);                               \
            pp.c -= lr_t * mm.c / (sqrtf(hh.c) + eps);              \
        }
        AMS1(x) AMS1(y) AMS1(z) AMS1(w)
#undef AMS1
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
        reinterpret_cast<float4*>(vhat)[i] = hh;
    }
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gi = g[i] * gscale;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        const float hi = fmaxf(vhat[i], vi);
        m[i] = mi; v[i] = vi; vhat[i] = hi;
        p[i] -= lr_t * mi / (sqrtf(hi) + eps);
    }
}

// fixed-order two-level sum of squares: part[blockIdx] then a single-thread-block finish
__global__ void sumsq_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ part) {
    __shared__ float red[32];
    float a = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        a = fmaf(g[i], g[i], a);
    a = block_sum(a, red);
    if (threadIdx.x == 0) part[blockIdx.x] = a;
}
__global__ void sumsq_final_kernel(const float* __restrict__ part, int nparts, float* __restrict__ sumsq) {
    __shared__ float red[32];
    float a = 0.f;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) a += part[i];
    a = block_sum(a, red);
    if (threadIdx.x == 0) sumsq[0] += a;
}
__global__ void clip_factor_kernel(const float* __restrict__ sumsq, float clip, float gscale, float* __restrict__ factor) {
    // tf.clip_by_global_norm: g * clip / max(global_norm, clip), the norm being that of gscale * g (gscale = 1 / world size:
    // the buffer holds the SUM of the per-rank gradients, the reference clips the gradient of the batch-mean loss)
    if (threadIdx.x == 0) factor[0] = clip / fmaxf(fabsf(gscale) * sqrtf(sumsq[0]), clip);
}

// tf.train.MomentumOptimizer(lr, momentum) (models/network.py:183): accum = momentum * accum + g ; p -= lr * accum.
// 5 array passes (read p,g,accum; write p,accum) = 20 B per parameter: HBM-bound.
__global__ void momentum_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ acc, int64_t n,
                                float lr, float momentum, float gscale, const float* __restrict__ gscale_dev) {
    if (gscale_dev) gscale *= gscale_dev[0];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = n >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 aa = reinterpret_cast<float4*>(acc)[i];
#define MOM1(c) { aa.c = momentum * aa.c + gg.c * gscale; pp.c -= lr * aa.c; }
        MOM1(x) MOM1(y) MOM1(z) MOM1(w)
#undef MOM1
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(acc)[i] = aa;
    }
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float a = momentum * acc[i] + g[i] * gscale;
        acc[i] = a;
        p[i] -= lr * a;
    }
}

// tf.train.RMSPropOptimizer(lr) (models/network.py:185; TF 1.x defaults decay 0.9, momentum 0, epsilon 1e-10, not centered;
// the `rms` slot starts at ONE, the `momentum` slot at zero -- the caller initialises them):
//   ms = decay * ms + (1 - decay) * g^2 ; mom = momentum * mom + lr * g * rsqrt(ms + eps) ; p -= mom.
// 7 array passes = 28 B per parameter.
__global__ void rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ ms,
                               float* __restrict__ mom, int64_t n, float lr, float decay, float momentum, float eps,
                               float gscale, const float* __restrict__ gscale_dev) {
    if (gscale_dev) gscale *= gscale_dev[0];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = n >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 ss = reinterpret_cast<float4*>(ms)[i];
        float4 mm = reinterpret_cast<float4*>(mom)[i];
#define RMS1(c)                                                     \
        {                                                           \
            const float gi = gg.c * gscale;                         \
            ss.c = decay * ss.c + (1.f - decay) * gi * gi;          \
            mm.c = momentum * mm.c + lr * gi / sqrtf(ss.c + eps);   \
            pp.c -= mm.c;                                           \
        }
        RMS1(x) RMS1(y) RMS1(z) RMS1(w)
#undef RMS1
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(ms)[i] = ss;
        reinterpret_cast<float4*>(mom)[i] = mm;
    }
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gi = g[i] * gscale;
        const float si = decay * ms[i] + (1.f - decay) * gi * gi;
        const float mi = momentum * mom[i] + lr * gi / sqrtf(si + eps);
        ms[i] = si; mom[i] = mi;
        p[i] -= mi;
    }
}

__global__ void make_filter_kernel(const float* __restrict__ window, const float* __restrict__ bases, int W, int N,
                                   float* __restrict__ filt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < W * N) filt[i] = fabsf(window[i / N]) * bases[i];
}
// dbases = |w| * dfilt ; dwindow[k] = sign(w[k]) * sum_n bases[k,n] * dfilt[k,n]
__global__ void make_filter_bwd_kernel(const float* __restrict__ window, const float* __restrict__ bases,
                                       const float* __restrict__ dfilt, int W, int N, float* __restrict__ dwindow,
                                       float* __restrict__ dbases) {
    __shared__ float red[32];
    const int k = blockIdx.x;
    const float w = window[k];
    const float aw = fabsf(w), sg = (w > 0.f) ? 1.f : ((w < 0.f) ? -1.f : 0.f);
    float a = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float d = dfilt[k * N + n];
        if (dbases) dbases[k * N + n] = aw * d;
        a = fmaf(bases[k * N + n], d, a);
    }
    a = block_sum(a, red);
    if (threadIdx.x == 0 && dwindow) dwindow[k] = sg * a;
}

}  // namespace
}  // namespace amss

using namespace amss;

extern "C" int amss_amsgrad_step(float* p, const float* g, float* m, float* v, float* vhat, int64_t n, float lr_t,
                                 float beta1, float beta2, float eps, float grad_scale, const float* grad_scale_dev,
                                 void* stream) {
    AMSS_REQUIRE(p && g && m && v && vhat && n > 0, "amsgrad_step: bad arguments");
    AMSS_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)vhat) & 15) == 0,
                 "amsgrad_step: buffers must be 16-byte aligned");
    const int grid = (int)std::min<int64_t>((n / 4 + 255) / 256 + 1, 8 * kNumSMs);
    AMSS_LAUNCH(amsgrad_kernel, grid, 256, 0, stream, p, g, m, v, vhat, n, lr_t, beta1, beta2, eps, grad_scale,
                grad_scale_dev);
    return AMSS_OK;
}

extern "C" size_t amss_sumsq_workspace_bytes(void) { return 1024 * 4; }
extern "C" int amss_sumsq(const float* g, int64_t n, float* sumsq, void* workspace, void* stream) {
    AMSS_REQUIRE(g && sumsq && workspace && n > 0, "sumsq: bad arguments");
    const int grid = (int)std::min<int64_t>((n + 255) / 256, 1024);
    AMSS_LAUNCH(sumsq_partial_kernel, grid, 256, 0, stream, g, n, (float*)workspace);
    AMSS_LAUNCH(sumsq_final_kernel, 1, 256, 0, stream, (const float*)workspace, grid, sumsq);
    return AMSS_OK;
}
extern "C" int amss_clip_factor(const float* sumsq, float clip, float grad_scale, float* factor, void* stream) {
    AMSS_REQUIRE(sumsq && factor && clip > 0.f, "clip_factor: bad arguments");
    AMSS_LAUNCH(clip_factor_kernel, 1, 32, 0, stream, sumsq, clip, grad_scale, factor);
    return AMSS_OK;
}

extern "C" int amss_momentum_step(float* p, const float* g, float* accum, int64_t n, float lr, float momentum,
                                  float grad_scale, const float* grad_scale_dev, void* stream) {
    AMSS_REQUIRE(p && g && accum && n > 0, "momentum_step: bad arguments");
    AMSS_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)accum) & 15) == 0, "momentum_step: buffers must be 16-byte aligned");
    const int grid = (int)std::min<int64_t>((n / 4 + 255) / 256 + 1, 8 * kNumSMs);
    AMSS_LAUNCH(momentum_kernel, grid, 256, 0, stream, p, g, accum, n, lr, momentum, grad_scale, grad_scale_dev);
    return AMSS_OK;
}

extern "C" int amss_rmsprop_step(float* p, const float* g, float* ms, float* mom, int64_t n, float lr, float decay,
                                 float momentum, float eps, float grad_scale, const float* grad_scale_dev, void* stream) {
    AMSS_REQUIRE(p && g && ms && mom && n > 0, "rmsprop_step: bad arguments");
    AMSS_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)ms | (uintptr_t)mom) & 15) == 0,
                 "rmsprop_step: buffers must be 16-byte aligned");
    const int grid = (int)std::min<int64_t>((n / 4 + 255) / 256 + 1, 8 * kNumSMs);
    AMSS_LAUNCH(rmsprop_kernel, grid, 256, 0, stream, p, g, ms, mom, n, lr, decay, momentum, eps, grad_scale, grad_scale_dev);
    return AMSS_OK;
}

extern "C" int amss_filterbank_make_filter(const float* window, const float* bases, int W, int N, float* filt,
                                           void* stream) {
    AMSS_REQUIRE(window && bases && filt && W > 0 && N > 0, "make_filter: bad arguments");
    AMSS_LAUNCH(make_filter_kernel, (W * N + 255) / 256, 256, 0, stream, window, bases, W, N, filt);
    return AMSS_OK;
}
extern "C" int amss_filterbank_make_filter_bwd(const float* window, const float* bases, const float* dfilt, int W,
                                               int N, float* dwindow, float* dbases, void* stream) {
    AMSS_REQUIRE(window && bases && dfilt, "make_filter_bwd: null pointer");
    AMSS_LAUNCH(make_filter_bwd_kernel, W, 128, 0, stream, window, bases, dfilt, W, N, dwindow, dbases);
    return AMSS_OK;
}
