// Adaptive front/back end: the learnable sparse 1-D conv filterbank of models/adapt.py.
//
//   analysis  (adapt.py:115-117): tf.nn.conv2d(SAME, stride 1) + max_pool_with_argmax(VALID),
//             fused -- the [Bt, L, N] tensor (65.5 MB per 4 s signal) never reaches HBM.
//   synthesis (utils/ops.py:94-120 unpool + adapt.py:241-243 conv2d_transpose), fused as a
//             sparse overlap-add gather: 99.6 % of the unpooled tensor is zeros.
//   backward  w.r.t. the filters, sparse through the arg-max.
//
// This file holds the fp32 SIMT path (AMSS_PREC_FP32, the parity path).  The tcgen05 Toeplitz
// implicit-GEMM analysis kernel lives in filterbank_tc.cu.
#include "common.cuh"
#include "tc.cuh"
#include <algorithm>

namespace amss {
int filterbank_analysis_tc(const float* x, const float* filt, int Bt, int L, int W, int N, int pool, int hop,
                           int precision, float* y, int64_t* argmax, void* workspace, size_t workspace_bytes,
                           cudaStream_t st);
size_t filterbank_analysis_tc_workspace(int Bt, int L, int W, int N, int pool, int hop, int precision);
bool filterbank_analysis_tc_supported(int L, int W, int N, int pool, int hop, int mode);
int filterbank_analysis_mix_tc(const float* x, const float* filt, int B, int S, int L, int W, int N, int pool, int hop,
                               int precision, float* y, int64_t* argmax, void* workspace, size_t workspace_bytes,
                               cudaStream_t st);

namespace {

constexpr int FA_TT = 256;   // time positions per sub-tile
constexpr int FA_NT = 64;    // filters per CTA
constexpr int FA_KC = 32;    // taps per shared-memory chunk
constexpr int FA_THREADS = 256;

__host__ __device__ inline void same_pad(int L, int W, int stride, int* out, int* pl) {
    const int o = (L + stride - 1) / stride;
    int pad = (o - 1) * stride + W - L;
    if (pad < 0) pad = 0;
    *out = o;
    *pl = pad / 2;
}

// MODE 0: max + argmax over [tp*hop, tp*hop+pool) ; MODE 1: mean over [tp*pool, (tp+1)*pool)
// grid (Tp, N/64, Bt).  Thread (tg = tid/8, fg = tid%8) owns times tg*8..+8 and filters fg*8..+8
// of the current 256-sample sub-tile; x slides through registers, filters come as float4.
template <int MODE>
__global__ void __launch_bounds__(FA_THREADS, 2)
analysis_pool_kernel(const float* __restrict__ x, const float* __restrict__ filt, int L, int W, int N, int pool,
                     int hop, int Tp, float* __restrict__ y, int64_t* __restrict__ argmax) {
    extern __shared__ __align__(16) unsigned char fb_smem[];
    const int Wp = (W + FA_KC - 1) / FA_KC * FA_KC;
    float* xs = reinterpret_cast<float*>(fb_smem);            // [FA_TT + Wp + 16]
    float* fs = xs + (FA_TT + Wp + 16);                        // [FA_KC][FA_NT]
    float* redv = fs + FA_KC * FA_NT;                         // [32][FA_NT]
    int* redt = reinterpret_cast<int*>(redv + 32 * FA_NT);    // [32][FA_NT]
    const int tp = blockIdx.x, n0 = blockIdx.y * FA_NT, r = blockIdx.z;
    const int tid = threadIdx.x, tg = tid >> 3, fg = tid & 7;
    const int pl = (W - 1) / 2;   // SAME padding, stride 1
    const int win0 = (MODE == 0) ? tp * hop : tp * pool;

    float best[8];
    int bestt[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = (MODE == 0) ? -INFINITY : 0.f; bestt[j] = 0; }

    for (int sub = 0; sub < pool; sub += FA_TT) {
        const int t0 = win0 + sub;                       // first time position of this sub-tile
        __syncthreads();
        for (int i = tid; i < FA_TT + Wp + 16; i += FA_THREADS) {
            const int s = t0 + i - pl;
            xs[i] = (s >= 0 && s < L && i < FA_TT + W) ? x[(size_t)r * L + s] : 0.f;
        }
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

        for (int k0 = 0; k0 < W; k0 += FA_KC) {
            __syncthreads();
            for (int i = tid; i < FA_KC * FA_NT; i += FA_THREADS) {
                const int kk = i / FA_NT, nn = i - kk * FA_NT;
                fs[i] = (k0 + kk < W && n0 + nn < N) ? filt[(size_t)(k0 + kk) * N + n0 + nn] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int k8 = 0; k8 < FA_KC; k8 += 8) {
                float xw[16];
                const float4* xp = reinterpret_cast<const float4*>(xs + tg * 8 + k0 + k8);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = xp[q];
                    xw[q * 4 + 0] = v.x; xw[q * 4 + 1] = v.y; xw[q * 4 + 2] = v.z; xw[q * 4 + 3] = v.w;
                }
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const float4 f0 = *reinterpret_cast<const float4*>(fs + (k8 + kk) * FA_NT + fg * 8);
                    const float4 f1 = *reinterpret_cast<const float4*>(fs + (k8 + kk) * FA_NT + fg * 8 + 4);
                    const float fv[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xw[kk + i], fv[j], acc[i][j]);
                }
            }
        }
        // fold this sub-tile into the running window statistic (ascending time, first max wins)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int tl = sub + tg * 8 + i;   // offset inside the pooling window
            if (tl < pool) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (MODE == 0) { if (acc[i][j] > best[j]) { best[j] = acc[i][j]; bestt[j] = win0 + tl; } }
                    else best[j] += acc[i][j];
                }
            }
        }
    }
    // cross-thread reduction over the 32 time groups, ascending
#pragma unroll
    for (int j = 0; j < 8; ++j) { redv[tg * FA_NT + fg * 8 + j] = best[j]; redt[tg * FA_NT + fg * 8 + j] = bestt[j]; }
    __syncthreads();
    if (tid < FA_NT && n0 + tid < N) {
        float b = redv[tid];
        int bt = redt[tid];
        for (int g = 1; g < 32; ++g) {
            const float v = redv[g * FA_NT + tid];
            if (MODE == 0) { if (v > b) { b = v; bt = redt[g * FA_NT + tid]; } }
            else b += v;
        }
        const size_t o = ((size_t)r * Tp + tp) * N + n0 + tid;
        if (MODE == 0) { y[o] = b; if (argmax) argmax[o] = (int64_t)bt * N + n0 + tid; }
        else y[o] = b / (float)pool;
    }
}

// strided conv, SAME padding (adapt.py:121-122): y[r,tp,n] = sum_k x[tp*hop + k - pl] filt[k,n]
__global__ void analysis_stride_kernel(const float* __restrict__ x, const float* __restrict__ filt, int L, int W,
                                       int N, int hop, int Tp, int pl, float* __restrict__ y) {
    extern __shared__ __align__(16) unsigned char fb_smem[];
    float* xs = reinterpret_cast<float*>(fb_smem);   // [W]
    const int tp = blockIdx.x, r = blockIdx.y;
    for (int k = threadIdx.x; k < W; k += blockDim.x) {
        const int s = tp * hop + k - pl;
        xs[k] = (s >= 0 && s < L) ? x[(size_t)r * L + s] : 0.f;
    }
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float a = 0.f;
        for (int k = 0; k < W; ++k) a = fmaf(xs[k], filt[(size_t)k * N + n], a);
        y[((size_t)r * Tp + tp) * N + n] = a;
    }
}

__global__ void transpose_filter_kernel(const float* __restrict__ f, int W, int N, float* __restrict__ fT) {
    __shared__ float tile[32][33];
    const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int k = k0 + i, n = n0 + threadIdx.x;
        tile[i][threadIdx.x] = (k < W && n < N) ? f[(size_t)k * N + n] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int n = n0 + i, k = k0 + threadIdx.x;
        if (n < N && k < W) fT[(size_t)n * W + k] = tile[threadIdx.x][i];
    }
}

// Sparse overlap-add gather.  grid (ceil(L/256), R).  Thread u accumulates, in a fixed atom order
// (frame ascending, filter ascending), every atom whose filter support covers sample u.
__global__ void __launch_bounds__(256)
synthesis_fwd_kernel(const float* __restrict__ vals, const int64_t* __restrict__ argmax,
                     const float* __restrict__ filtT, int S, int L, int W, int N, int Tp, int hop_lo, int pool_hi,
                     float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char fb_smem[];
    float* sv = reinterpret_cast<float*>(fb_smem);    // [N]
    int* sp = reinterpret_cast<int*>(sv + N);         // [N]
    const int r = blockIdx.y, b = r / S, u0 = blockIdx.x * 256, u = u0 + threadIdx.x;
    const int pl = (W - 1) / 2;
    // atoms of frame tp sit at pos in [tp*hop_lo, tp*hop_lo + pool_hi); k = u - pos + pl in [0, W)
    const int pos_min = u0 - (W - 1 - pl), pos_max = u0 + 255 + pl;
    int tp_lo = (pos_min - pool_hi + 1);
    tp_lo = tp_lo <= 0 ? 0 : (tp_lo + hop_lo - 1) / hop_lo;
    int tp_hi = pos_max / hop_lo;
    if (tp_hi > Tp - 1) tp_hi = Tp - 1;
    float acc = 0.f;
    for (int tp = tp_lo; tp <= tp_hi; ++tp) {
        __syncthreads();
        for (int n = threadIdx.x; n < N; n += blockDim.x) {
            sv[n] = vals[((size_t)r * Tp + tp) * N + n];
            sp[n] = (int)(argmax[((size_t)b * Tp + tp) * N + n] / N);
        }
        __syncthreads();
        for (int n = 0; n < N; ++n) {
            const int k = u - sp[n] + pl;
            if (k >= 0 && k < W) acc = fmaf(sv[n], __ldg(filtT + (size_t)n * W + k), acc);
        }
    }
    if (u < L) out[(size_t)r * L + u] = acc;
}

// ---- filter-stationary sparse overlap-add (W = 1024 taps, pool = hop = 256) ---------------------------------------------------
// The gather above reads one filter tap from L1 / L2 per multiply-add.  Here a CTA keeps SF_F filters in shared memory
// (64 KB, staged by TMA bulk copies) and walks the frames of ONE mixture in time order for all of its S rows at once (they share the arg-max
// positions, adapt.py:212-218): an atom (frame tp, filter f) at pos = 256 tp + off touches the samples
// [256 tp - 511, 256 tp + 767], i.e. five 256-sample blocks starting at block tp - 2.  Thread t owns sample t of each of
// the five blocks; with kk = t - off - 1 its taps are kk + 256 m for the blocks m = 1..3 and the single tap kk & 1023 for
// block 0 (kk >= 0) or block 4 (kk < 0): four shared-memory loads (consecutive threads, consecutive taps) feed 5 S
// multiply-adds.  After frame tp block tp - 2 is complete for this filter group: it is stored and the five accumulators
// slide by one block.  Partial outputs [N / SF_F][R][L] are summed in filter-group order by a second kernel, so the result does
// not depend on the launch geometry.  A CTA owns SF_Q consecutive blocks (it starts two frames early and ends two frames late).
constexpr int SF_F = 16;                 // filters per CTA (64 KB of taps)
constexpr int SF_Q = 50;                 // 256-sample output blocks per CTA

template <int S>
__global__ void __launch_bounds__(256, 3)
synthesis_fwd_fs_kernel(const float* __restrict__ vals, const int64_t* __restrict__ argmax, const float* __restrict__ filtT,
                        int L, int N, int Tp, float* __restrict__ part) {
    constexpr int W = 1024;
    extern __shared__ __align__(16) unsigned char fb_smem[];
    float* ws = reinterpret_cast<float*>(fb_smem);                    // [SF_F][W]
    __shared__ int s_pos[2][SF_F];
    __shared__ float s_val[2][S][SF_F];
    const int g = blockIdx.x, b = blockIdx.y, q0 = blockIdx.z * SF_Q, t = threadIdx.x, n0 = g * SF_F;
    const int nblk = (L + 255) / 256, q1 = min(q0 + SF_Q, nblk);
    // taps of the group: ws[f][k] = filtT[n0 + f][k], one 4 KB TMA bulk copy per filter from the transposed bank (completion
    // on an mbarrier); filters beyond N are zero-filled by the threads
    __shared__ __align__(8) uint64_t tap_bar;
    const uint32_t bar = tc::smem_u32(&tap_bar);
    const int nf = min(SF_F, N - n0);
    if (t == 0) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
    for (int i = t; i < (SF_F - nf) * W; i += 256) ws[nf * W + i] = 0.f;
    __syncthreads();
    if (t == 0) {
        tc::mbar_expect_tx(bar, (uint32_t)nf * W * 4);
        for (int f = 0; f < nf; ++f) tc::bulk_g2s(tc::smem_u32(ws + f * W), filtT + (size_t)(n0 + f) * W, W * 4, bar);
    }
    const int tp0 = max(q0 - 2, 0), tp1 = min(q1 + 2, Tp);             // frames that touch the blocks [q0, q1)
    auto stage = [&](int tp, int buf) {                               // positions / values of frame tp for the group's filters
        if (t < SF_F) {
            const int n = n0 + t;
            // an arg-max of the pooling window lies in [256 tp, 256 tp + 256); the clamp keeps the tap indices inside the staged
            // filters when the caller hands over garbage (e.g. the arg-max of an all-NaN window after a diverged step)
            const int off = n < N ? (int)((uint32_t)argmax[((size_t)b * Tp + tp) * N + n] / (uint32_t)N) - 256 * tp : 0;
            s_pos[buf][t] = min(max(off, 0), 255);
        } else if (t < SF_F * (S + 1)) {
            const int s = t / SF_F - 1, f = t - (s + 1) * SF_F, n = n0 + f;
            s_val[buf][s][f] = n < N ? vals[(((size_t)b * S + s) * Tp + tp) * N + n] : 0.f;
        }
    };
    float acc[S][5];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int m = 0; m < 5; ++m) acc[s][m] = 0.f;
    if (tp0 < tp1) stage(tp0, 0);
    tc::mbar_wait(bar, 0);                                            // the taps have landed
    __syncthreads();
    for (int tp = tp0; tp < tp1; ++tp) {
        const int buf = (tp - tp0) & 1;
        if (tp + 1 < tp1) stage(tp + 1, buf ^ 1);                     // next frame's atoms while this one is accumulated
#pragma unroll 4
        for (int f = 0; f < SF_F; ++f) {
            const int kk = t - s_pos[buf][f] - 1;                     // off in [0, 256): kk in [-256, 255]
            const float* wf = ws + f * W;
            const float x04 = wf[kk & 1023], x1 = wf[kk + 256], x2 = wf[kk + 512], x3 = wf[kk + 768];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const float v = s_val[buf][s][f];
                const float v0 = kk >= 0 ? v : 0.f;
                acc[s][0] = fmaf(v0, x04, acc[s][0]);
                acc[s][1] = fmaf(v, x1, acc[s][1]);
                acc[s][2] = fmaf(v, x2, acc[s][2]);
                acc[s][3] = fmaf(v, x3, acc[s][3]);
                acc[s][4] = fmaf(v - v0, x04, acc[s][4]);
            }
        }
        const int q = tp - 2;                                         // block completed by this frame
        if (q >= q0 && q < q1) {
            const int u = 256 * q + t;
            if (u < L) {
#pragma unroll
                for (int s = 0; s < S; ++s) part[((size_t)g * gridDim.y * S + (size_t)b * S + s) * L + u] = acc[s][0];
            }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
#pragma unroll
            for (int m = 0; m < 4; ++m) acc[s][m] = acc[s][m + 1];
            acc[s][4] = 0.f;
        }
        __syncthreads();                                              // staged frame visible; this frame's buffer free
    }
    // frames beyond Tp do not exist: the blocks still held by the accumulators are complete (slot m = block tp1 - 2 + m after
    // the last shift); blocks no frame reaches are zero
    for (int q = max(tp1 - 2, q0); q < q1; ++q) {
        const int m = q - (tp1 - 2);
        const int u = 256 * q + t;
        if (u < L) {
#pragma unroll
            for (int s = 0; s < S; ++s) {
                float v = 0.f;
#pragma unroll
                for (int mm = 0; mm < 5; ++mm) v = mm == m ? acc[s][mm] : v;
                part[((size_t)g * gridDim.y * S + (size_t)b * S + s) * L + u] = v;
            }
        }
    }
}

// out[i] = sum_g part[g][i] in group order (i over R * L), 16-byte vectors when n % 4 == 0
__global__ void synthesis_sum_groups_kernel(const float* __restrict__ part, int groups, int64_t n, float* __restrict__ out) {
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(part) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        for (int64_t i = i0; i < n / 4; i += stride) {
            float4 a = __ldcs(reinterpret_cast<const float4*>(part) + i);
            for (int g = 1; g < groups; ++g) {
                const float4 v = __ldcs(reinterpret_cast<const float4*>(part + (size_t)g * n) + i);
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
            reinterpret_cast<float4*>(out)[i] = a;
        }
    } else {
        for (int64_t i = i0; i < n; i += stride) {
            float a = part[i];
            for (int g = 1; g < groups; ++g) a += part[(size_t)g * n + i];
            out[i] = a;
        }
    }
}

// dvals[r,tp,n] = sum_k dout[r, pos + k - pl] * filt2[k,n] : one warp per atom, lanes over taps.
__global__ void __launch_bounds__(256)
synthesis_bwd_vals_kernel(const float* __restrict__ dout, const int64_t* __restrict__ argmax,
                          const float* __restrict__ filtT, int S, int L, int W, int N, int Tp,
                          float* __restrict__ dvals) {
    const int tp = blockIdx.x, r = blockIdx.y, b = r / S;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int pl = (W - 1) / 2;
    for (int n = w; n < N; n += nw) {
        const int pos = (int)(argmax[((size_t)b * Tp + tp) * N + n] / N);
        float a = 0.f;
        for (int k = lane; k < W; k += 32) {
            const int s = pos + k - pl;
            if (s >= 0 && s < L) a = fmaf(dout[(size_t)r * L + s], filtT[(size_t)n * W + k], a);
        }
        a = warp_sum(a);
        if (lane == 0) dvals[((size_t)r * Tp + tp) * N + n] = a;
    }
}

// dfilt[k,n] = sum_{r,tp} vals[r,tp,n] * sig[r, pos(r,tp,n) + k - pl].  grid (N, chunks): a CTA
// walks its slice of the (r,tp) atom list in order; partials [chunks][N][W] are summed in order.
// argdiv: rows of argmax = r / argdiv (S for the synthesis: the mixture's argmax; 1 for analysis).
__global__ void __launch_bounds__(256)
sparse_filter_grad_kernel(const float* __restrict__ vals, const int64_t* __restrict__ argmax,
                          const float* __restrict__ sig, int R, int argdiv, int L, int W, int N, int Tp,
                          float* __restrict__ part) {
    const int n = blockIdx.x, chunk = blockIdx.y, chunks = gridDim.y;
    const int pl = (W - 1) / 2;
    const int64_t natoms = (int64_t)R * Tp;
    const int64_t a0 = natoms * chunk / chunks, a1 = natoms * (chunk + 1) / chunks;
    for (int kb = 0; kb < W; kb += 256 * 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int64_t a = a0; a < a1; ++a) {
            const int r = (int)(a / Tp), tp = (int)(a - (int64_t)r * Tp);
            const float v = vals[((size_t)r * Tp + tp) * N + n];
            const int pos = (int)(argmax[((size_t)(r / argdiv) * Tp + tp) * N + n] / N);
            if (v != 0.f) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int k = kb + q * 256 + threadIdx.x;
                    const int s = pos + k - pl;
                    if (k < W && s >= 0 && s < L) acc[q] = fmaf(v, sig[(size_t)r * L + s], acc[q]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = kb + q * 256 + threadIdx.x;
            if (k < W) part[((size_t)chunk * N + n) * W + k] = acc[q];
        }
    }
}
// ---- filter-stationary versions of the two sparse reductions (W <= 1024) -------------------------------------------------
// A CTA owns SG_F consecutive filters and a slice of the (row, frame) pairs.  For one (row, frame) the arg-max positions of
// all filters lie inside one pooling window, so the sample windows [pos - pl, pos - pl + W) of the CTA's filters overlap
// almost completely: the union (<= pool + W - 1 samples) is staged in shared memory ONCE and every filter reads its W taps
// from there at its own offset -- consecutive threads read consecutive words.  The per-atom global traffic of the
// one-warp-per-atom kernels above (2 x 4 KB from L2 per atom) becomes one 5 KB span per SG_F atoms.  The (row, frame) list
// is walked in blocks of SG_AB pairs: positions / values of a block are fetched up front, and the spans of the next SG_NS - 2
// pairs are in flight (cp.async ring, zero-filled outside the signal) while pair i is accumulated: one __syncthreads per pair
// (with a single span in flight every pair paid most of an L2 round trip: 3.2 -> 1.x ms for the analysis backward).  Thread t owns taps k = t, t + 256, t + 512, t + 768.
constexpr int SG_F = 8;
constexpr int SG_SPAN = 2560;            // staged samples per pair (a longer span takes the direct global path)
constexpr int SG_AB = 64;                // (row, frame) pairs per block
constexpr int SG_NS = 4;                 // span buffers: spans i+1 .. i+SG_NS-2 are in flight while pair i is accumulated

struct __align__(16) SgBlock {
    int pos[SG_AB][SG_F];        // 16-byte aligned rows (SG_F = 8): a pair's positions / values are read as two vectors each
    float val[SG_AB][SG_F];
    int pmin[SG_AB], len[SG_AB], row[SG_AB], frame[SG_AB];   // (row, frame) split once per pair: no 64-bit division in the loops
};

// positions (and values) of the CTA's filters for pairs [a, a + cnt): coalesced over the SG_F consecutive filters
__device__ __forceinline__ void sg_load_block(SgBlock& blk, const float* __restrict__ vals, const int64_t* __restrict__ argmax,
                                              int64_t a, int cnt, int argdiv, int W, int N, int Tp, int n0) {
    const uint32_t a32 = (uint32_t)a, N32 = (uint32_t)N;         // R * Tp and the per-sample index L * N fit 32 bits (checked on the host)
    for (int u = threadIdx.x; u < cnt * SG_F; u += blockDim.x) {
        const int i = u / SG_F, f = u - i * SG_F, n = n0 + f;
        const uint32_t at = a32 + (uint32_t)i;
        const int r = (int)(at / (uint32_t)Tp), tp = (int)(at - (uint32_t)r * (uint32_t)Tp);
        const bool ok = n < N;
        blk.pos[i][f] = ok ? (int)((uint32_t)argmax[((size_t)(r / argdiv) * Tp + tp) * N + n] / N32) : -1;
        if (vals) blk.val[i][f] = ok ? vals[((size_t)r * Tp + tp) * N + n] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const uint32_t at = a32 + (uint32_t)i;
        blk.row[i] = (int)(at / (uint32_t)Tp);
        blk.frame[i] = (int)(at - (uint32_t)blk.row[i] * (uint32_t)Tp);
        int lo = 0x7fffffff, hi = -1;
#pragma unroll
        for (int f = 0; f < SG_F; ++f) { const int q = blk.pos[i][f]; if (q >= 0) { lo = min(lo, q); hi = max(hi, q); } }
        blk.pmin[i] = hi < 0 ? 0 : lo;
        blk.len[i] = hi < 0 ? 0 : hi - lo + W;
    }
    __syncthreads();
}

// asynchronous copy of sig[r, pmin - pl .. + len) into xs (zero-filled outside [0, L)); one commit group per call.
// vec16 (L % 4 == 0, 16-byte aligned rows): the span is fetched as 16-byte chunks starting at the 4-sample boundary below its
// first sample -- a chunk is then either entirely inside or entirely outside the row -- and the consumer reads at
// xs[sg_adj(pmin - pl) + ...]; otherwise sample by sample.
__device__ __forceinline__ int sg_adj(int base, bool vec16) { return vec16 ? (base & 3) : 0; }
__device__ __forceinline__ void sg_issue_span(float* xs, const float* __restrict__ sig, int r, int L, int pl, int pmin, int len,
                                              bool vec16) {
    if (len > 0 && len <= SG_SPAN) {
        const int base = pmin - pl;
        const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(xs);
        if (vec16) {
            const int s0 = base - (base & 3), n4 = (len + (base & 3) + 3) >> 2;
            for (int i = threadIdx.x; i < n4; i += blockDim.x) {
                const int s = s0 + 4 * i;
                const bool ok = s >= 0 && s < L;
                const float* src = sig + (size_t)r * L + (ok ? s : 0);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + 16u * i), "l"(src), "r"(ok ? 16u : 0u) : "memory");
            }
        } else {
            for (int i = threadIdx.x; i < len; i += blockDim.x) {
                const int s = base + i;
                const bool ok = s >= 0 && s < L;
                const float* src = sig + (size_t)r * L + (ok ? s : 0);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst0 + 4u * i), "l"(src), "r"(ok ? 4u : 0u) : "memory");
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N_> __device__ __forceinline__ void sg_wait_pending() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// dfilt partials: part[chunk][n][k] = sum over the chunk's (r,tp) of vals[r,tp,n] * sig[r, pos + k - pl]
template <bool FULLW>   // W == 1024: thread t owns exactly the taps t, t + 256, t + 512, t + 768
__global__ void __launch_bounds__(256)
sparse_filter_grad_fs_kernel(const float* __restrict__ vals, const int64_t* __restrict__ argmax,
                             const float* __restrict__ sig, int R, int argdiv, int L, int W, int N, int Tp,
                             float* __restrict__ part) {
    __shared__ SgBlock blk;
    __shared__ __align__(16) float xs[SG_NS][SG_SPAN + 8];
    const int n0 = blockIdx.x * SG_F, chunk = blockIdx.y, chunks = gridDim.y, tid = threadIdx.x, pl = (W - 1) / 2;
    const bool vec16 = (L & 3) == 0 && (reinterpret_cast<uintptr_t>(sig) & 15) == 0;
    const int64_t natoms = (int64_t)R * Tp;
    const int64_t a0 = natoms * chunk / chunks, a1 = natoms * (chunk + 1) / chunks;
    float acc[SG_F][4];
#pragma unroll
    for (int f = 0; f < SG_F; ++f)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[f][j] = 0.f;
    for (int64_t ab = a0; ab < a1; ab += SG_AB) {
        const int cnt = (int)min((int64_t)SG_AB, a1 - ab);
        __syncthreads();                                        // the previous block's entries / spans are no longer read
        sg_load_block(blk, vals, argmax, ab, cnt, argdiv, W, N, Tp, n0);
        for (int i = 0; i < SG_NS - 1; ++i) {                   // prologue: spans 0 .. SG_NS-2 (empty groups past the end)
            if (i < cnt) sg_issue_span(xs[i], sig, blk.row[i], L, pl, blk.pmin[i], blk.len[i], vec16);
            else asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int i = 0; i < cnt; ++i) {
            const int r = blk.row[i];
            sg_wait_pending<SG_NS - 2>();                       // this thread's copies of span i have landed
            __syncthreads();                                    // span i visible to all; buffer (i-1) % SG_NS free again
            if (i + SG_NS - 1 < cnt)
                sg_issue_span(xs[(i + SG_NS - 1) % SG_NS], sig, blk.row[i + SG_NS - 1], L, pl, blk.pmin[i + SG_NS - 1],
                              blk.len[i + SG_NS - 1], vec16);
            else asm volatile("cp.async.commit_group;" ::: "memory");
            const int pmin = blk.pmin[i], len = blk.len[i];
            const float* xb = xs[i % SG_NS] + sg_adj(pmin - pl, vec16) - pmin + tid;
            int posv[SG_F];
            float valv[SG_F];
#pragma unroll
            for (int f4 = 0; f4 < SG_F; f4 += 4) {              // the pair's positions / values: vector loads (broadcast)
                const int4 pq = *reinterpret_cast<const int4*>(&blk.pos[i][f4]);
                const float4 vq = *reinterpret_cast<const float4*>(&blk.val[i][f4]);
                posv[f4] = pq.x; posv[f4 + 1] = pq.y; posv[f4 + 2] = pq.z; posv[f4 + 3] = pq.w;
                valv[f4] = vq.x; valv[f4 + 1] = vq.y; valv[f4 + 2] = vq.z; valv[f4 + 3] = vq.w;
            }
            if (FULLW && len <= SG_SPAN) {                      // W = 1024 taps, span staged: 4 loads + 4 FMAs per filter, no tests
#pragma unroll
                for (int f = 0; f < SG_F; ++f) {
                    const float* xf = xb + posv[f];
                    const float v = valv[f];
                    if (v != 0.f) {                             // warp-uniform (a filter beyond N carries value 0)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[f][j] = fmaf(v, xf[256 * j], acc[f][j]);
                    }
                }
            } else {
#pragma unroll
                for (int f = 0; f < SG_F; ++f) {
                    const float v = valv[f];
                    if (v != 0.f) {                             // warp-uniform
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int k = tid + 256 * j;
                            if (k < W) {
                                float x;
                                if (len <= SG_SPAN) x = xb[posv[f] + 256 * j];
                                else { const int s = posv[f] + k - pl; x = (s >= 0 && s < L) ? sig[(size_t)r * L + s] : 0.f; }
                                acc[f][j] = fmaf(v, x, acc[f][j]);
                            }
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int f = 0; f < SG_F; ++f)
        if (n0 + f < N)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = tid + 256 * j;
                if (k < W) part[((size_t)chunk * N + n0 + f) * W + k] = acc[f][j];
            }
}

// dvals[r,tp,n] = sum_k dout[r, pos + k - pl] * filt2[k,n]: the CTA's filter taps live in registers for the whole kernel
template <bool FULLW>
__global__ void __launch_bounds__(256)
synthesis_bwd_vals_fs_kernel(const float* __restrict__ dout, const int64_t* __restrict__ argmax,
                             const float* __restrict__ filt2, int R, int S, int L, int W, int N, int Tp,
                             float* __restrict__ dvals) {
    __shared__ SgBlock blk;
    __shared__ __align__(16) float xs[SG_NS][SG_SPAN + 8];
    __shared__ float red[2][8][SG_F];
    const int n0 = blockIdx.x * SG_F, chunk = blockIdx.y, chunks = gridDim.y, tid = threadIdx.x, pl = (W - 1) / 2;
    const int lane = tid & 31, warp = tid >> 5;
    const bool vec16 = (L & 3) == 0 && (reinterpret_cast<uintptr_t>(dout) & 15) == 0;
    float w[SG_F][4];
#pragma unroll
    for (int f = 0; f < SG_F; ++f)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = tid + 256 * j;
            w[f][j] = (k < W && n0 + f < N) ? __ldg(filt2 + (size_t)k * N + n0 + f) : 0.f;
        }
    const int64_t natoms = (int64_t)R * Tp;
    const int64_t a0 = natoms * chunk / chunks, a1 = natoms * (chunk + 1) / chunks;
    for (int64_t ab = a0; ab < a1; ab += SG_AB) {
        const int cnt = (int)min((int64_t)SG_AB, a1 - ab);
        __syncthreads();
        sg_load_block(blk, nullptr, argmax, ab, cnt, S, W, N, Tp, n0);
        for (int i = 0; i < SG_NS - 1; ++i) {
            if (i < cnt) sg_issue_span(xs[i], dout, blk.row[i], L, pl, blk.pmin[i], blk.len[i], vec16);
            else asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int i = 0; i < cnt; ++i) {
            const int r = blk.row[i];
            sg_wait_pending<SG_NS - 2>();
            __syncthreads();                                    // span i visible; red[(i-1)&1] complete; buffer (i-1) % SG_NS free
            if (i > 0 && tid < SG_F && n0 + tid < N) {          // finish pair i-1: fixed-order sum of the 8 warp partials
                const int rp = blk.row[i - 1], tpp = blk.frame[i - 1];
                float t = 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) t += red[(i - 1) & 1][q][tid];
                dvals[((size_t)rp * Tp + tpp) * N + n0 + tid] = t;
            }
            if (i + SG_NS - 1 < cnt)
                sg_issue_span(xs[(i + SG_NS - 1) % SG_NS], dout, blk.row[i + SG_NS - 1], L, pl, blk.pmin[i + SG_NS - 1],
                              blk.len[i + SG_NS - 1], vec16);
            else asm volatile("cp.async.commit_group;" ::: "memory");
            const int pmin = blk.pmin[i], len = blk.len[i];
            const float* xb = xs[i % SG_NS] + sg_adj(pmin - pl, vec16) - pmin + tid;
            int posv[SG_F];
#pragma unroll
            for (int f4 = 0; f4 < SG_F; f4 += 4) {
                const int4 pq = *reinterpret_cast<const int4*>(&blk.pos[i][f4]);
                posv[f4] = pq.x; posv[f4 + 1] = pq.y; posv[f4 + 2] = pq.z; posv[f4 + 3] = pq.w;
            }
            float pacc[SG_F];
#pragma unroll
            for (int f = 0; f < SG_F; ++f) {
                pacc[f] = 0.f;
                if (posv[f] >= 0) {                             // warp-uniform
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = tid + 256 * j;
                        if (FULLW || k < W) {
                            float x;
                            if (len <= SG_SPAN) x = xb[posv[f] + 256 * j];
                            else { const int s = posv[f] + k - pl; x = (s >= 0 && s < L) ? dout[(size_t)r * L + s] : 0.f; }
                            pacc[f] = fmaf(w[f][j], x, pacc[f]);
                        }
                    }
                }
            }
            {   // warp reduction of the SG_F = 8 partial sums together: a halving butterfly (4 + 2 + 1 exchanges) leaves lane l with
                // the 4-lane partial of filter l >> 2, two more exchanges finish it: 9 shuffles instead of 8 x 5
                static_assert(SG_F == 8, "the butterfly below reduces exactly 8 values");
                float q4[4], q2[2];
                const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float keep = b4 ? pacc[j + 4] : pacc[j], give = b4 ? pacc[j] : pacc[j + 4];
                    q4[j] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float keep = b3 ? q4[j + 2] : q4[j], give = b3 ? q4[j] : q4[j + 2];
                    q2[j] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
                }
                float q1 = (b2 ? q2[1] : q2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? q2[0] : q2[1], 4);
                q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
                q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
                if ((lane & 3) == 0) red[i & 1][warp][lane >> 2] = q1;
            }
        }
        __syncthreads();                                        // the last pair of the block
        if (tid < SG_F && n0 + tid < N) {
            const int rp = blk.row[cnt - 1], tpp = blk.frame[cnt - 1];
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) t += red[(cnt - 1) & 1][q][tid];
            dvals[((size_t)rp * Tp + tpp) * N + n0 + tid] = t;
        }
    }
}

// dfilt[k][n] (+)= sum_chunks part[chunk][n][k]
__global__ void filter_grad_final_kernel(const float* __restrict__ part, int chunks, int W, int N, int accumulate,
                                         float* __restrict__ dfilt) {
    __shared__ float tile[32][33];
    const int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int n = n0 + i, k = k0 + threadIdx.x;
        float a = 0.f;
        if (n < N && k < W)
            for (int c = 0; c < chunks; ++c) a += part[((size_t)c * N + n) * W + k];
        tile[i][threadIdx.x] = a;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int k = k0 + i, n = n0 + threadIdx.x;
        if (k < W && n < N) {
            const float v = tile[threadIdx.x][i];
            dfilt[(size_t)k * N + n] = accumulate ? dfilt[(size_t)k * N + n] + v : v;
        }
    }
}

// the filter-stationary kernels index (row, frame) pairs and arg-max values with 32-bit arithmetic
static bool sg_fits_32bit(int R, int Tp, int L, int N) { return (int64_t)R * Tp < (1ll << 31) && (int64_t)L * N < (1ll << 32); }
constexpr int FG_CHUNKS = 32;     // slices of the (row, frame) list per filter (gather kernels)
// Slices per filter GROUP of the filter-stationary kernels: four CTAs of 256 threads are resident per SM (64 registers, 46 KB of
// shared memory), so the grid is cut to fill whole waves of 4 x SMs CTAs: 256 filters = 32 groups x 37 slices = 1184 = two full
// waves (32 x 32 = 1024 ran as one full wave + one at 73 %).
constexpr int FG_CHUNKS_MAX = 96;
inline int fg_chunks_fs(int N) {
    const int groups = (N + 7) / 8, resident = 4 * kNumSMs;
    int c = 2 * resident / groups;
    if (c > FG_CHUNKS_MAX) c = resident / groups;
    return std::max(1, std::min(c, FG_CHUNKS_MAX));
}

}  // namespace
}  // namespace amss

using namespace amss;

extern "C" int amss_filterbank_analysis_out_frames(int L, int W, int pool, int hop, int mode) {
    if (mode == AMSS_POOL_MAX) return L >= pool ? (L - pool) / hop + 1 : 0;
    if (mode == AMSS_POOL_AVG) return L / pool;
    int out, pl;
    same_pad(L, W, hop, &out, &pl);
    return out;
}

extern "C" size_t amss_filterbank_analysis_workspace_bytes(int Bt, int L, int W, int N, int pool, int hop, int mode,
                                                           int precision) {
    if (filterbank_analysis_tc_supported(L, W, N, pool, hop, mode) && precision != AMSS_PREC_FP32)
        return filterbank_analysis_tc_workspace(Bt, L, W, N, pool, hop, precision);
    return 256;
}

extern "C" int amss_filterbank_analysis_fwd(const float* x, const float* filt, int Bt, int L, int W, int N, int pool,
                                            int hop, int mode, int precision, float* y, int64_t* argmax,
                                            void* workspace, size_t workspace_bytes, void* stream) {
    AMSS_REQUIRE(x && filt && y, "filterbank_analysis_fwd: null pointer");
    AMSS_REQUIRE(Bt > 0 && L > 0 && W > 0 && N > 0 && hop > 0, "filterbank_analysis_fwd: bad sizes");
    AMSS_REQUIRE(mode >= 0 && mode <= 2, "filterbank_analysis_fwd: bad mode %d", mode);
    const int Tp = amss_filterbank_analysis_out_frames(L, W, pool, hop, mode);
    AMSS_REQUIRE(Tp > 0, "filterbank_analysis_fwd: no output frames (L=%d pool=%d)", L, pool);
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == AMSS_POOL_STRIDE) {
        int out, pl;
        same_pad(L, W, hop, &out, &pl);
        dim3 grid(Tp, Bt);
        AMSS_LAUNCH(analysis_stride_kernel, grid, 256, align_up((size_t)W * 4, 16), st,   // ptxas widens the tail reads to LDS.128
                    x, filt, L, W, N, hop, Tp, pl, y);
        return AMSS_OK;
    }
    AMSS_REQUIRE(pool > 0, "filterbank_analysis_fwd: pool must be positive");
    if (mode == AMSS_POOL_MAX && precision != AMSS_PREC_FP32 &&
        filterbank_analysis_tc_supported(L, W, N, pool, hop, mode))
        return filterbank_analysis_tc(x, filt, Bt, L, W, N, pool, hop, precision, y, argmax, workspace,
                                      workspace_bytes, st);
    const int Wp = (W + FA_KC - 1) / FA_KC * FA_KC;
    const size_t smem = ((size_t)(FA_TT + Wp + 16) + FA_KC * FA_NT + 32 * FA_NT * 2) * 4;
    dim3 grid(Tp, (N + FA_NT - 1) / FA_NT, Bt);
    if (mode == AMSS_POOL_MAX) {
        AMSS_CUDA(cudaFuncSetAttribute(analysis_pool_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        AMSS_LAUNCH(analysis_pool_kernel<0>, grid, FA_THREADS, smem, st, x, filt, L, W, N, pool, hop, Tp, y, argmax);
    } else {
        AMSS_CUDA(cudaFuncSetAttribute(analysis_pool_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        AMSS_LAUNCH(analysis_pool_kernel<1>, grid, FA_THREADS, smem, st, x, filt, L, W, N, pool, hop, Tp, y, argmax);
    }
    return AMSS_OK;
}

extern "C" int amss_filterbank_analysis_mix_fwd(const float* x, const float* filt, int B, int S, int L, int W, int N,
                                                int pool, int hop, int mode, int precision, float* y, int64_t* argmax,
                                                void* workspace, size_t workspace_bytes, void* stream) {
    AMSS_REQUIRE(B > 0 && S > 0, "filterbank_analysis_mix_fwd: bad batch B=%d S=%d", B, S);
    if (x && filt && y && mode == AMSS_POOL_MAX && precision != AMSS_PREC_FP32 && pool > 0 &&
        filterbank_analysis_tc_supported(L, W, N, pool, hop, mode))
        return filterbank_analysis_mix_tc(x, filt, B, S, L, W, N, pool, hop, precision, y, argmax, workspace,
                                          workspace_bytes, (cudaStream_t)stream);
    return amss_filterbank_analysis_fwd(x, filt, B * (S + 1), L, W, N, pool, hop, mode, precision, y, argmax, workspace,
                                        workspace_bytes, stream);
}

extern "C" size_t amss_filterbank_grad_workspace_bytes(int W, int N) {
    return align_up((size_t)W * N * 4, 256) + (size_t)std::max(FG_CHUNKS, fg_chunks_fs(N)) * N * W * 4;
}

extern "C" int amss_filterbank_analysis_bwd(const float* x, const float* dy, const int64_t* argmax, int Bt, int L,
                                            int W, int N, int Tp, int accumulate, float* dfilt, void* workspace,
                                            size_t workspace_bytes, void* stream) {
    AMSS_REQUIRE(x && dy && argmax && dfilt && workspace, "filterbank_analysis_bwd: null pointer");
    if (workspace_bytes < amss_filterbank_grad_workspace_bytes(W, N)) { set_error("filterbank_analysis_bwd: workspace too small"); return AMSS_ERR_WORKSPACE; }
    float* part = (float*)((char*)workspace + align_up((size_t)W * N * 4, 256));
    int chunks = FG_CHUNKS;
    if (W <= 1024 && sg_fits_32bit(Bt, Tp, L, N)) {
        chunks = fg_chunks_fs(N);
        dim3 grid((N + SG_F - 1) / SG_F, chunks);
        if (W == 1024) AMSS_LAUNCH(sparse_filter_grad_fs_kernel<true>, grid, 256, 0, stream, dy, argmax, x, Bt, 1, L, W, N, Tp, part);
        else AMSS_LAUNCH(sparse_filter_grad_fs_kernel<false>, grid, 256, 0, stream, dy, argmax, x, Bt, 1, L, W, N, Tp, part);
    } else {
        dim3 grid(N, FG_CHUNKS);
        AMSS_LAUNCH(sparse_filter_grad_kernel, grid, 256, 0, stream, dy, argmax, x, Bt, 1, L, W, N, Tp, part);
    }
    dim3 g2((N + 31) / 32, (W + 31) / 32), b2(32, 8);
    AMSS_LAUNCH(filter_grad_final_kernel, g2, b2, 0, stream, part, chunks, W, N, accumulate, dfilt);
    return AMSS_OK;
}

// the filter-stationary forward: the reference geometry only (other shapes take the gather kernel)
static bool synthesis_fs_ok(int S, int W, int N, int pool, int hop, int L) {
    return W == 1024 && pool == 256 && hop == 256 && S >= 1 && S <= 4 && N % 4 == 0 && (int64_t)L * N < (1ll << 32);
}
static size_t synthesis_fs_bytes(int B, int S, int L, int N) {
    return (size_t)((N + SF_F - 1) / SF_F) * (size_t)B * S * L * 4;
}

extern "C" size_t amss_filterbank_synthesis_workspace_bytes(int B, int S, int L, int W, int N, int Tp) {
    (void)Tp;
    // (pool / hop are not known here: the partial-output buffer of the filter-stationary forward is reserved whenever the
    // rest of the geometry allows it)
    const size_t fs = synthesis_fs_ok(S, W, N, 256, 256, L) ? align_up((size_t)W * N * 4, 256) + synthesis_fs_bytes(B, S, L, N) : 0;
    return std::max(amss_filterbank_grad_workspace_bytes(W, N), fs);
}

extern "C" int amss_filterbank_synthesis_fwd(const float* vals, const int64_t* argmax, const float* filt2, int B,
                                             int S, int L, int W, int N, int Tp, int pool, int hop, float* out,
                                             void* workspace, size_t workspace_bytes, void* stream) {
    AMSS_REQUIRE(vals && argmax && filt2 && out && workspace, "filterbank_synthesis_fwd: null pointer");
    AMSS_REQUIRE(B > 0 && S > 0 && pool > 0 && hop > 0, "filterbank_synthesis_fwd: bad sizes");
    if (workspace_bytes < (size_t)W * N * 4) { set_error("filterbank_synthesis_fwd: workspace too small"); return AMSS_ERR_WORKSPACE; }
    float* fT = (float*)workspace;
    const size_t fs_off = align_up((size_t)W * N * 4, 256);
    const bool no_fs = getenv("AMSS_SYNTHESIS_GATHER") != nullptr;             // A/B and parity tests: force the gather kernel
    if (!no_fs && synthesis_fs_ok(S, W, N, pool, hop, L) && workspace_bytes >= fs_off + synthesis_fs_bytes(B, S, L, N)) {
        float* part = (float*)((char*)workspace + fs_off);
        {   // the transposed bank fT[n][k] = filt2[k][n]: a filter's 1024 taps become one contiguous 4 KB bulk-copy source
            dim3 gt((W + 31) / 32, (N + 31) / 32), bt(32, 8);
            AMSS_LAUNCH(transpose_filter_kernel, gt, bt, 0, stream, filt2, W, N, fT);
        }
        const int groups = (N + SF_F - 1) / SF_F, nblk = (L + 255) / 256;
        dim3 grid(groups, B, (nblk + SF_Q - 1) / SF_Q);
        const size_t smem = (size_t)SF_F * 1024 * 4;
#define AMSS_SYN_FS(SS)                                                                                                         \
    do {                                                                                                                        \
        AMSS_CUDA(cudaFuncSetAttribute(synthesis_fwd_fs_kernel<SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
        AMSS_LAUNCH(synthesis_fwd_fs_kernel<SS>, grid, 256, smem, stream, vals, argmax, fT, L, N, Tp, part);                    \
    } while (0)
        if (S == 1) AMSS_SYN_FS(1); else if (S == 2) AMSS_SYN_FS(2); else if (S == 3) AMSS_SYN_FS(3); else AMSS_SYN_FS(4);
#undef AMSS_SYN_FS
        const int64_t n = (int64_t)B * S * L;
        AMSS_LAUNCH(synthesis_sum_groups_kernel, 8 * kNumSMs, 256, 0, stream, part, groups, n, out);
        return AMSS_OK;
    }
    dim3 gt((W + 31) / 32, (N + 31) / 32), bt(32, 8);
    AMSS_LAUNCH(transpose_filter_kernel, gt, bt, 0, stream, filt2, W, N, fT);
    dim3 grid((L + 255) / 256, B * S);
    AMSS_LAUNCH(synthesis_fwd_kernel, grid, 256, (size_t)N * 8, stream, vals, argmax, fT, S, L, W, N, Tp, hop, pool,
                out);
    return AMSS_OK;
}

extern "C" int amss_filterbank_synthesis_bwd(const float* dout, const float* vals, const int64_t* argmax,
                                             const float* filt2, int B, int S, int L, int W, int N, int Tp,
                                             float* dvals, float* dfilt2, void* workspace, size_t workspace_bytes,
                                             void* stream) {
    AMSS_REQUIRE(dout && vals && argmax && filt2 && workspace, "filterbank_synthesis_bwd: null pointer");
    if (workspace_bytes < amss_filterbank_grad_workspace_bytes(W, N)) { set_error("filterbank_synthesis_bwd: workspace too small"); return AMSS_ERR_WORKSPACE; }
    float* fT = (float*)workspace;
    float* part = (float*)((char*)workspace + align_up((size_t)W * N * 4, 256));
    const bool fs = W <= 1024 && sg_fits_32bit(B * S, Tp, L, N);
    if (dvals && fs) {
        dim3 grid((N + SG_F - 1) / SG_F, fg_chunks_fs(N));
        if (W == 1024) AMSS_LAUNCH(synthesis_bwd_vals_fs_kernel<true>, grid, 256, 0, stream, dout, argmax, filt2, B * S, S, L, W, N, Tp, dvals);
        else AMSS_LAUNCH(synthesis_bwd_vals_fs_kernel<false>, grid, 256, 0, stream, dout, argmax, filt2, B * S, S, L, W, N, Tp, dvals);
    } else if (dvals) {
        dim3 gt((W + 31) / 32, (N + 31) / 32), bt(32, 8);
        AMSS_LAUNCH(transpose_filter_kernel, gt, bt, 0, stream, filt2, W, N, fT);
        dim3 grid(Tp, B * S);
        AMSS_LAUNCH(synthesis_bwd_vals_kernel, grid, 256, 0, stream, dout, argmax, fT, S, L, W, N, Tp, dvals);
    }
    if (dfilt2 && fs) {
        const int chunks = fg_chunks_fs(N);
        dim3 grid((N + SG_F - 1) / SG_F, chunks);
        if (W == 1024) AMSS_LAUNCH(sparse_filter_grad_fs_kernel<true>, grid, 256, 0, stream, vals, argmax, dout, B * S, S, L, W, N, Tp, part);
        else AMSS_LAUNCH(sparse_filter_grad_fs_kernel<false>, grid, 256, 0, stream, vals, argmax, dout, B * S, S, L, W, N, Tp, part);
        dim3 g2((N + 31) / 32, (W + 31) / 32), b2(32, 8);
        AMSS_LAUNCH(filter_grad_final_kernel, g2, b2, 0, stream, part, chunks, W, N, 0, dfilt2);
    } else if (dfilt2) {
        dim3 grid(N, FG_CHUNKS);
        AMSS_LAUNCH(sparse_filter_grad_kernel, grid, 256, 0, stream, vals, argmax, dout, B * S, S, L, W, N, Tp, part);
        dim3 g2((N + 31) / 32, (W + 31) / 32), b2(32, 8);
        AMSS_LAUNCH(filter_grad_final_kernel, g2, b2, 0, stream, part, FG_CHUNKS, W, N, 0, dfilt2);
    }
    return AMSS_OK;
}
