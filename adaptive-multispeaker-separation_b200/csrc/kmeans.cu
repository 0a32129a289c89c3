// In-graph multi-try hard/soft k-means (reference: models/Kmeans_2.py:14-188) and the mask
// application of Separator.separate (models/network.py:554-582).
//
// Design: the reference tiles X `tries` times and materialises [B*tries, L, K, E] broadcast
// temporaries per iteration.  Here X[B,L,E] is read ONCE per iteration: a CTA stages a tile of
// points in shared memory, every point is labelled against all tries' centroids (phase 1), and
// the per-(try, cluster) sums are accumulated by one owner thread per (try, feature) in a fixed
// order (phase 2), so the result is deterministic (no floating-point atomics anywhere).
// Per-CTA partials are reduced in a fixed order by a small finalize kernel.
//
// HBM traffic: the input is L2-normalised ON THE FLY while a tile is staged (tf.nn.l2_normalize, Kmeans_2.py:40-41): no
// normalised copy is written or re-read.  The batch is processed in GROUPS of mixtures: all iterations + the inertia pass +
// the final assignment of a group run before the next group starts.  A group that fits the 126 MB L2 (9 mixtures of 10.24 MB
// at TF = 64000, E = 40) fetches X from HBM once; since the tensor-core pass became issue-bound rather than byte-bound the
// default group is larger (km_group: fewer, longer passes amortise the fixed cost of a pass) and X streams from HBM.
#include "common.cuh"
#include "kmeans_pieces.cuh"
#include <algorithm>
#include <cstdlib>

namespace amss {

// tensor-core pass (kmeans_tc.cu): hard assignments, E = 40, K in {2,3,4}, tries*K <= 32, no silence gate
bool kmeans_tc_supported(int E, int K, int tries, bool soft, bool gated);
void kmeans_tc_set_profile(long long* dev_buf);
void kmeans_tc_geometry(int Bg, int64_t L, int* G_out, int* pmax_out);
int kmeans_pass_tc(const float* X, const float* cent, const float* prev_part, float* cent_out, int Bg, int64_t L, int K, int tries,
                   int chunks, int normalize, int mode, float* part, cudaStream_t st);

namespace {

constexpr int KM_TILE = 256;     // points per tile
constexpr int KM_THREADS = 256;
constexpr int KM_MAXK = 8;
constexpr int KM_MAXG = 8;

enum { KM_UPDATE = 0, KM_INERTIA = 1 };

struct KmSmemLayout {
    size_t xs, cs, lab, w, wv, ssum, total;
    int G, NF;
};

__host__ __device__ inline KmSmemLayout km_layout(int E, int K, int tries, int mode, int soft) {
    KmSmemLayout s;
    const int EE = (mode == KM_UPDATE) ? E : 1;
    s.NF = EE + 1;  // + the count feature
    int G = KM_THREADS / (tries * s.NF);
    if (G < 1) G = 1;
    if (G > KM_MAXG) G = KM_MAXG;
    s.G = G;
    size_t off = 0;
    s.xs = off;  off += (size_t)KM_TILE * (E + 1) * 4;
    s.cs = off;  off += (size_t)((tries * K * E + 3) / 4 * 4) * 4;
    s.lab = off; off += (size_t)((tries * KM_TILE + 15) / 16 * 16);            // u8 labels (hard)
    s.w = off;   off += soft ? (size_t)tries * KM_TILE * K * 4 : 0;             // soft weights
    s.wv = off;  off += (mode == KM_INERTIA) ? (size_t)tries * KM_TILE * (soft ? K : 1) * 4 : 0;
    s.ssum = off; off += (size_t)G * tries * K * s.NF * 4;
    s.total = off;
    return s;
}

// One pass over X for all tries.  grid = (chunks, B).  cent[R,K,E], R = B*tries (row r=b*tries+t).
// part[b][chunk][t][k][f], f in [0,NF): f<EE feature sums, f==EE the count / weight sum.
template <int MODE, int SOFT>
__global__ void __launch_bounds__(KM_THREADS)
kmeans_pass_kernel(const float* __restrict__ X, const float* __restrict__ cent,
                   const uint8_t* __restrict__ notsilent, int B, int b_off, int normalize, int64_t L, int E, int K,
                   int tries, float beta, float* __restrict__ part) {
    extern __shared__ __align__(16) unsigned char km_smem[];
    const KmSmemLayout lay = km_layout(E, K, tries, MODE, SOFT);
    float* xs = reinterpret_cast<float*>(km_smem + lay.xs);
    float* cs = reinterpret_cast<float*>(km_smem + lay.cs);
    uint8_t* lab = km_smem + lay.lab;
    float* wsm = reinterpret_cast<float*>(km_smem + lay.w);
    float* wv = reinterpret_cast<float*>(km_smem + lay.wv);
    float* ssum = reinterpret_cast<float*>(km_smem + lay.ssum);
    const int G = lay.G, NF = lay.NF, EE = NF - 1;
    const int b = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
    const int tid = threadIdx.x;
    const int EP = E + 1;

    for (int i = tid; i < tries * K * E; i += KM_THREADS) cs[i] = cent[(size_t)b * tries * K * E + i];
    for (int i = tid; i < G * tries * K * NF; i += KM_THREADS) ssum[i] = 0.f;
    __syncthreads();

    const int64_t ntiles = (L + KM_TILE - 1) / KM_TILE;
    for (int64_t tile = chunk; tile < ntiles; tile += chunks) {
        const int64_t p0 = tile * KM_TILE;
        const int np = (int)((L - p0) < KM_TILE ? (L - p0) : KM_TILE);
        // stage the tile: coalesced read of np*E contiguous floats
        const float* src = X + ((size_t)b * L + p0) * E;
        for (int i = tid; i < np * E; i += KM_THREADS) {
            const int p = i / E, e = i - p * E;
            xs[p * EP + e] = src[i];
        }
        __syncthreads();
        if (normalize) {      // x * rsqrt(max(sum x^2, 1e-12)) per point (Kmeans_2.py:40-41), in place in the staged tile
            if (tid < np) {
                float* xp = xs + tid * EP;
                float ss = 0.f;
                for (int e = 0; e < E; ++e) ss = fmaf(xp[e], xp[e], ss);
                const float inv = rsqrtf(fmaxf(ss, 1e-12f));
                for (int e = 0; e < E; ++e) xp[e] *= inv;
            }
            __syncthreads();
        }
        // ---- phase 1: one thread per point, all tries -------------------------------------
        if (tid < np) {
            const float* xp = xs + tid * EP;
            for (int t = 0; t < tries; ++t) {
                // the reference tiles notsilent try-major while X is batch-major
                // (Kmeans_2.py:80 vs :48-52): row r of [B*tries] uses mask row r % B (b counts from the start of the
                // whole batch: the kernel sees a group of it, b_off mixtures in).
                const int r = (b_off + b) * tries + t;
                float ns = 1.f;
                if (notsilent) ns = notsilent[(size_t)(r % B) * L + p0 + tid] ? 1.f : 0.f;
                float d2u[KM_MAXK];
#pragma unroll
                for (int k = 0; k < KM_MAXK; ++k) {
                    if (k < K) {
                        const float* c = cs + (t * K + k) * E;
                        float a = 0.f;
                        for (int e = 0; e < E; ++e) { const float d = xp[e] - c[e]; a = fmaf(d, d, a); }
                        d2u[k] = a;
                    }
                }
                if (SOFT) {
                    float ev[KM_MAXK], tot = 0.f;
#pragma unroll
                    for (int k = 0; k < KM_MAXK; ++k)
                        if (k < K) { ev[k] = expf(-1.f * beta * (d2u[k] * ns)); tot += ev[k]; }
#pragma unroll
                    for (int k = 0; k < KM_MAXK; ++k)
                        if (k < K) {
                            const float w = ev[k] / tot;
                            wsm[(t * KM_TILE + tid) * K + k] = w;
                            if (MODE == KM_INERTIA) wv[(t * KM_TILE + tid) * K + k] = d2u[k] * w;
                        }
                } else {
                    int best = 0;
                    float bd = sqrtf(d2u[0] * ns);
#pragma unroll
                    for (int k = 1; k < KM_MAXK; ++k)
                        if (k < K) { const float d = sqrtf(d2u[k] * ns); if (d < bd) { bd = d; best = k; } }
                    lab[t * KM_TILE + tid] = (uint8_t)best;
                    if (MODE == KM_INERTIA) {
                        float sel = d2u[0];
#pragma unroll
                        for (int k = 1; k < KM_MAXK; ++k) if (k < K && k == best) sel = d2u[k];
                        wv[t * KM_TILE + tid] = sel;
                    }
                }
            }
        }
        __syncthreads();
        // ---- phase 2: owner thread per (group, try, feature), fixed point order -------------
        const int nitems = G * tries * NF;
        for (int item = tid; item < nitems; item += KM_THREADS) {
            const int f = item % NF;
            const int t = (item / NF) % tries;
            const int g = item / (NF * tries);
            const int r = (b_off + b) * tries + t;
            const uint8_t* nsrow = notsilent ? notsilent + (size_t)(r % B) * L + p0 : nullptr;
            float acc[KM_MAXK];
#pragma unroll
            for (int k = 0; k < KM_MAXK; ++k) acc[k] = 0.f;
            for (int p = g; p < np; p += G) {
                float val;
                if (f == EE) val = 1.f;                       // count / weight-sum feature
                else if (MODE == KM_UPDATE) {
                    val = xs[p * EP + f];
                    if (nsrow && !nsrow[p]) val = 0.f;        // X * notsilent (Kmeans_2.py:148)
                } else val = 0.f;                             // inertia: value depends on k (below)
                if (SOFT) {
                    const float* wp = wsm + (t * KM_TILE + p) * K;
                    const float* wvp = wv + (t * KM_TILE + p) * K;
#pragma unroll
                    for (int k = 0; k < KM_MAXK; ++k)
                        if (k < K) {
                            if (MODE == KM_INERTIA && f == 0) acc[k] += wvp[k];
                            else acc[k] = fmaf(wp[k], val, acc[k]);
                        }
                } else {
                    const int l = lab[t * KM_TILE + p];
                    if (MODE == KM_INERTIA && f == 0) val = wv[t * KM_TILE + p];
#pragma unroll
                    for (int k = 0; k < KM_MAXK; ++k)
                        if (k < K) acc[k] += (l == k) ? val : 0.f;
                }
            }
#pragma unroll
            for (int k = 0; k < KM_MAXK; ++k)
                if (k < K) ssum[((g * tries + t) * K + k) * NF + f] += acc[k];
        }
        __syncthreads();
    }
    // combine groups in fixed order, write this CTA's partial
    float* dst = part + ((size_t)b * chunks + chunk) * tries * K * NF;
    for (int i = tid; i < tries * K * NF; i += KM_THREADS) {
        float a = 0.f;
        for (int g = 0; g < G; ++g) a += ssum[(size_t)g * tries * K * NF + i];
        dst[i] = a;
    }
}

// cent[r][k][e] = sum_chunks part_sum / sum_chunks part_cnt   (Kmeans_2.py:158-165 / :151-155)
__global__ void kmeans_finalize_kernel(const float* __restrict__ part, int B, int chunks, int tries, int K,
                                       int E, float* __restrict__ cent) {
    const int NF = E + 1;
    const int64_t n = (int64_t)B * tries * K * E;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(i % E);
        const int64_t rk = i / E;               // (b*tries+t)*K+k
        const int b = (int)(rk / ((int64_t)tries * K));
        const int tk = (int)(rk % ((int64_t)tries * K));
        float s = 0.f, c = 0.f;
        for (int ch = 0; ch < chunks; ++ch) {
            const float* p = part + (((size_t)b * chunks + ch) * tries * K + tk) * NF;
            s += p[e];
            c += p[E];
        }
        cent[i] = s / c;   // empty cluster -> 0/0 = NaN, as in the reference
    }
}

// inertia[b][t] = sum_k tot_k / cnt_k ; best = argmin_t (first minimum; a NaN inertia -- empty
// cluster -- is never selected, as with tf.argmin's `<` reducer) ; centroids_out[b] = cent[b*tries+best]   (Kmeans_2.py:99-104)
__global__ void kmeans_select_kernel(const float* __restrict__ part, const float* __restrict__ cent, int B,
                                     int chunks, int tries, int K, int E, float* __restrict__ inertia_out,
                                     int* __restrict__ best_out, float* __restrict__ cent_out, int64_t nt, int64_t Tt, int G) {
    const int b = blockIdx.x;
    if (G > 0) {                // partials of the tensor-core pass: mixture b has kt_pieces() of the `chunks` slots filled
        int c0, n;
        kt_pieces(b, nt, Tt, G, c0, n);
        const int pitch = chunks;
        chunks = n;
        part += (size_t)b * (pitch - n) * tries * K * 2;        // rows of mixture b start at b * pitch
    }
    __shared__ int sbest;
    __shared__ float ratio[128];
    // one thread per (try, cluster): the chunks in sequence, as before; thread 0 adds the clusters of a try in order
    for (int tk = threadIdx.x; tk < tries * K && tk < 128; tk += blockDim.x) {
        float s = 0.f, c = 0.f;
        for (int ch = 0; ch < chunks; ++ch) {
            const float* p = part + (((size_t)b * chunks + ch) * tries * K + tk) * 2;
            s += p[0];
            c += p[1];
        }
        ratio[tk] = s / c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int best = 0;
        float bv = 0.f;
        for (int t = 0; t < tries; ++t) {
            float in = 0.f;
            for (int k = 0; k < K; ++k) {
                float q;
                if (t * K + k < 128) q = ratio[t * K + k];
                else {
                    float s = 0.f, c = 0.f;
                    for (int ch = 0; ch < chunks; ++ch) {
                        const float* p = part + (((size_t)b * chunks + ch) * tries * K + t * K + k) * 2;
                        s += p[0];
                        c += p[1];
                    }
                    q = s / c;
                }
                in += q;
            }
            if (inertia_out) inertia_out[b * tries + t] = in;
            if (t == 0) { bv = isnan(in) ? INFINITY : in; best = 0; }
            else if (in < bv) { bv = in; best = t; }   // NaN never wins, first minimum wins ties
        }
        sbest = best;
        best_out[b] = best;
    }
    __syncthreads();
    const int best = sbest;
    for (int i = threadIdx.x; i < K * E; i += blockDim.x)
        cent_out[(size_t)b * K * E + i] = cent[((size_t)b * tries + best) * K * E + i];
}

// centroids0[r][k] = X[b][init_idx[r][k]]   (Kmeans_2.py:66-71)
// (normalised on the fly like every other read of X: one warp per initial row)
__global__ void kmeans_gather_init_kernel(const float* __restrict__ X, const int* __restrict__ idx, int B,
                                          int64_t L, int E, int K, int tries, int normalize, float* __restrict__ cent) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t rows = (int64_t)B * tries * K;
    for (int64_t rk = warp; rk < rows; rk += nwarps) {
        const int b = (int)(rk / ((int64_t)tries * K));
        const float* x = X + ((size_t)b * L + idx[rk]) * E;
        float inv = 1.f;
        if (normalize) {                    // same sequential order as the tile path: bit-identical rows
            float ss = 0.f;
            for (int e = 0; e < E; ++e) ss = fmaf(x[e], x[e], ss);
            inv = rsqrtf(fmaxf(ss, 1e-12f));
        }
        for (int e = lane; e < E; e += 32) cent[rk * E + e] = normalize ? x[e] * inv : x[e];
    }
}

// Final assignment against one centroid set per batch row   (Kmeans_2.py:106-107, 169-188)
template <int SOFT>
__global__ void __launch_bounds__(KM_THREADS)
kmeans_assign_kernel(const float* __restrict__ X, const float* __restrict__ cent, const uint8_t* __restrict__ notsilent,
                     const int* __restrict__ best, int B, int b_off, int normalize, int64_t L, int E, int K, int tries,
                     float beta, int* __restrict__ labels, float* __restrict__ soft) {
    extern __shared__ __align__(16) unsigned char km_smem[];
    float* xs = reinterpret_cast<float*>(km_smem);
    float* cs = xs + KM_TILE * (E + 1);
    const int b = blockIdx.y, tid = threadIdx.x, EP = E + 1;
    for (int i = tid; i < K * E; i += KM_THREADS) cs[i] = cent[(size_t)b * K * E + i];
    const uint8_t* nsrow = nullptr;
    if (notsilent) nsrow = notsilent + (size_t)(((b_off + b) * tries + best[b]) % B) * L;
    const int64_t ntiles = (L + KM_TILE - 1) / KM_TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p0 = tile * KM_TILE;
        const int np = (int)((L - p0) < KM_TILE ? (L - p0) : KM_TILE);
        const float* src = X + ((size_t)b * L + p0) * E;
        __syncthreads();
        for (int i = tid; i < np * E; i += KM_THREADS) {
            const int p = i / E, e = i - p * E;
            xs[p * EP + e] = src[i];
        }
        __syncthreads();
        if (tid < np) {
            float* xw = xs + tid * EP;
            if (normalize) {
                float ss = 0.f;
                for (int e = 0; e < E; ++e) ss = fmaf(xw[e], xw[e], ss);
                const float inv = rsqrtf(fmaxf(ss, 1e-12f));
                for (int e = 0; e < E; ++e) xw[e] *= inv;
            }
            const float* xp = xw;
            const float ns = nsrow ? (nsrow[p0 + tid] ? 1.f : 0.f) : 1.f;
            float d2[KM_MAXK];
#pragma unroll
            for (int k = 0; k < KM_MAXK; ++k)
                if (k < K) {
                    const float* c = cs + k * E;
                    float a = 0.f;
                    for (int e = 0; e < E; ++e) { const float d = xp[e] - c[e]; a = fmaf(d, d, a); }
                    d2[k] = a * ns;
                }
            if (SOFT) {
                float ev[KM_MAXK], tot = 0.f;
#pragma unroll
                for (int k = 0; k < KM_MAXK; ++k) if (k < K) { ev[k] = expf(-1.f * beta * d2[k]); tot += ev[k]; }
#pragma unroll
                for (int k = 0; k < KM_MAXK; ++k) if (k < K) soft[((size_t)b * L + p0 + tid) * K + k] = ev[k] / tot;
            } else {
                int bi = 0;
                float bd = sqrtf(d2[0]);
#pragma unroll
                for (int k = 1; k < KM_MAXK; ++k)
                    if (k < K) { const float d = sqrtf(d2[k]); if (d < bd) { bd = d; bi = k; } }
                labels[(size_t)b * L + p0 + tid] = bi;
            }
        }
    }
}

// Hard final assignment, E = 40, no silence gate (the inference path of config 5): the arithmetic of kmeans_assign_kernel<0>
// per point (sequential sums, sqrt, first minimum), with the tile staged by coalesced 16-byte loads into a 44-float pitch
// (conflict-free LDS.128 of a row) and the row held in registers.
constexpr int KA_E = 40, KA_PITCH = 44, KA_TILE = 256;
__global__ void __launch_bounds__(KA_TILE)
kmeans_assign_hard40_kernel(const float* __restrict__ X, const float* __restrict__ cent, int normalize, int64_t L, int K,
                            int* __restrict__ labels) {
    __shared__ __align__(16) float xs[KA_TILE * KA_PITCH];
    __shared__ __align__(16) float cs[KM_MAXK * KA_E];
    const int b = blockIdx.y, tid = threadIdx.x;
    for (int i = tid; i < K * KA_E; i += KA_TILE) cs[i] = cent[(size_t)b * K * KA_E + i];
    const int64_t ntiles = (L + KA_TILE - 1) / KA_TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p0 = tile * KA_TILE;
        const int np = (int)((L - p0) < KA_TILE ? (L - p0) : KA_TILE);
        const float4* src = reinterpret_cast<const float4*>(X + ((size_t)b * L + p0) * KA_E);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < KA_E / 4; ++j) {
            const int u = tid + KA_TILE * j, rr = u / (KA_E / 4), c = u - rr * (KA_E / 4);
            if (u < np * (KA_E / 4)) *reinterpret_cast<float4*>(xs + rr * KA_PITCH + c * 4) = __ldg(src + u);
        }
        __syncthreads();
        if (tid < np) {
            float x[KA_E];
#pragma unroll
            for (int c = 0; c < KA_E / 4; ++c) {
                const float4 v = *reinterpret_cast<const float4*>(xs + tid * KA_PITCH + c * 4);
                x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
            }
            if (normalize) {
                float ss = 0.f;
#pragma unroll
                for (int e = 0; e < KA_E; ++e) ss = fmaf(x[e], x[e], ss);
                const float inv = rsqrtf(fmaxf(ss, 1e-12f));
#pragma unroll
                for (int e = 0; e < KA_E; ++e) x[e] *= inv;
            }
            int bi = 0;
            float bd = 0.f;
#pragma unroll
            for (int k = 0; k < KM_MAXK; ++k)
                if (k < K) {
                    float a = 0.f;
#pragma unroll
                    for (int c = 0; c < KA_E / 4; ++c) {
                        const float4 cv = *reinterpret_cast<const float4*>(cs + k * KA_E + c * 4);
                        float d = x[4 * c] - cv.x; a = fmaf(d, d, a);
                        d = x[4 * c + 1] - cv.y; a = fmaf(d, d, a);
                        d = x[4 * c + 2] - cv.z; a = fmaf(d, d, a);
                        d = x[4 * c + 3] - cv.w; a = fmaf(d, d, a);
                    }
                    const float dk = sqrtf(a);
                    if (k == 0 || dk < bd) { bd = dk; bi = k; }
                }
            labels[(size_t)b * L + p0 + tid] = bi;
        }
    }
}

// max over L of latent per batch row, then notsilent = log10(max/latent) < threshold (Kmeans_2.py:76-79)
__global__ void row_max_kernel(const float* __restrict__ x, int64_t L, float* __restrict__ out) {
    __shared__ float red[32];
    const int b = blockIdx.x;
    float m = -INFINITY;
    for (int64_t i = threadIdx.x; i < L; i += blockDim.x) m = fmaxf(m, x[(size_t)b * L + i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
        m = warp_max(m);
        if (threadIdx.x == 0) out[b] = m;
    }
}
__global__ void silence_mask_kernel(const float* __restrict__ x, const float* __restrict__ rowmax, int64_t L,
                                    int64_t n, float threshold, uint8_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = logf(rowmax[i / L] / x[i]) / logf(10.f);   // utils/ops.py:56-59 log10
        out[i] = v < threshold ? 1 : 0;
    }
}

// separated[b*S+s][i] = X_input[b][i] * mask[b][i][s]   (network.py:567-580)
__global__ void apply_masks_kernel(const float* __restrict__ X, const int* __restrict__ labels,
                                   const float* __restrict__ soft, int B, int S, int64_t TF, float* __restrict__ out) {
    const int64_t n = (int64_t)B * TF;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / TF, j = i - b * TF;
        const float x = X[i];
        for (int s = 0; s < S; ++s) {
            const float m = labels ? (labels[i] == s ? 1.f : 0.f) : soft[i * S + s];
            out[((size_t)b * S + s) * TF + j] = x * m;
        }
    }
}

int km_chunks(int B, int64_t L) {
    int64_t ntiles = (L + KM_TILE - 1) / KM_TILE;
    int64_t want = (2 * kNumSMs + B - 1) / B;
    if (want < 1) want = 1;
    return (int)(ntiles < want ? ntiles : want);
}

// The tensor-core pass runs ONE CTA per SM (512 TMEM columns, ~210 KB of shared memory) in ONE wave: the Bg * ntiles tiles of
// the group are cut into min(SMs, tiles) equal contiguous ranges, whatever Bg is (kmeans_pieces.cuh; with whole chunks per
// mixture 64 mixtures filled 128 of the 148 SMs, and a group of 9 launched 297 CTAs = two waves + ONE straggler CTA).
// km_chunks_tc = the most pieces a mixture is cut into = the row pitch of the partial sums.
int km_chunks_tc(int Bg, int64_t L) {
    int G, pmax;
    kmeans_tc_geometry(Bg, L, &G, &pmax);
    return pmax;
}

// Mixtures per group: as many as keep the group's X inside ~3/4 of the 126 MB L2 (at least 1), then evened out over the
// groups (64 mixtures = 8 groups of 8 rather than 7 of 9 and a group of one).
int km_group(int B, int64_t L, int E) {
    const double per = (double)L * E * 4.0;
    // The passes are bound by instruction / tensor issue (~1700 clk per 128-point tile, 20 KB), not by bytes: 64 mixtures in
    // one group stream 655 MB per pass from HBM at ~2.8 TB/s and finish sooner (2.85 ms) than eight L2-resident groups of
    // 8 (3.61 ms), because every pass of every group pays ~14 us of launch / prologue / drain.  AMSS_KMEANS_GROUP_MB=96
    // restores the L2-resident schedule (12x fewer DRAM bytes).
    double budget = 700.0e6;
    if (const char* e = getenv("AMSS_KMEANS_GROUP_MB")) { const double v = atof(e); if (v > 0.0) budget = v * 1.0e6; }
    int g = (int)(budget / per);
    if (g < 1) g = 1;
    if (g >= B) return B;
    const int ngroups = (B + g - 1) / g;
    return (B + ngroups - 1) / ngroups;
}

struct KmWs {
    float *cent, *part, *part2;
    size_t total;
};
KmWs km_ws(void* base, int B, int64_t L, int E, int K, int tries) {
    KmWs w;
    size_t off = 0;
    char* p = (char*)base;
    const int G = km_group(B, L, E);
    w.cent = (float*)(p + off); off += align_up((size_t)B * tries * K * E * 4, 256);
    // Bg * km_chunks(Bg) <= 2 * SMs + Bg - 1 for every group size Bg <= G (a smaller last group gets more chunks per mixture)
    // (and Bg * pieces <= SMs + 2 * Bg for the flat decomposition of the tensor-core pass)
    w.part = (float*)(p + off); off += align_up((size_t)(2 * kNumSMs + 2 * G) * tries * K * (E + 1) * 4, 256);
    // second buffer: the tensor-core pass reduces the previous pass's partial sums in its prologue while it writes its own
    w.part2 = (float*)(p + off); off += align_up((size_t)(2 * kNumSMs + 2 * G) * tries * K * (E + 1) * 4, 256);
    w.total = off;
    return w;
}

template <int MODE, int SOFT>
int launch_pass(const float* X, const float* cent, const uint8_t* ns, int Bg, int B, int b_off, int normalize, int64_t L,
                int E, int K, int tries, float beta, float* part, cudaStream_t st) {
    const KmSmemLayout lay = km_layout(E, K, tries, MODE, SOFT);
    AMSS_REQUIRE(lay.total <= 227 * 1024, "kmeans: shared memory %zu B exceeds 227 KB (E=%d K=%d tries=%d)",
                 lay.total, E, K, tries);
    AMSS_CUDA(cudaFuncSetAttribute(kmeans_pass_kernel<MODE, SOFT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)lay.total));
    dim3 grid(km_chunks(Bg, L), Bg);
    AMSS_LAUNCH((kmeans_pass_kernel<MODE, SOFT>), grid, KM_THREADS, lay.total, st, X, cent, ns, B, b_off, normalize, L, E, K,
                tries, beta, part);
    return AMSS_OK;
}

}  // namespace
}  // namespace amss

using namespace amss;

// Diagnostics: clock64() stamps of the tensor-core update pass (CTA 0, tiles 8..11, 8 slots per tile for the loader, the MMA
// issuer and the first epilogue thread) are written to dev_buf (>= 96 int64) by subsequent amss_kmeans_fit calls; NULL = off.
extern "C" int amss_debug_kmeans_profile(long long* dev_buf) {
    kmeans_tc_set_profile(dev_buf);
    return AMSS_OK;
}

extern "C" size_t amss_kmeans_workspace_bytes(int B, int64_t L, int E, int K, int tries) {
    return km_ws(nullptr, B, L, E, K, tries).total;
}

extern "C" int amss_kmeans_fit(const float* X, const int32_t* init_idx, const uint8_t* notsilent, int B, int64_t L,
                               int E, int K, int tries, int iters, float beta, int normalize_input,
                               int assign_at_end, float* centroids, int32_t* labels, float* soft, float* inertia,
                               int32_t* best_try, void* workspace, size_t workspace_bytes, void* stream) {
    AMSS_REQUIRE(X && init_idx && centroids && workspace && best_try, "kmeans_fit: null pointer");
    AMSS_REQUIRE(B > 0 && L > 0 && E > 0 && tries > 0 && iters >= 0, "kmeans_fit: bad sizes");
    AMSS_REQUIRE(K >= 1 && K <= KM_MAXK, "kmeans_fit: K=%d outside [1,%d]", K, KM_MAXK);
    AMSS_REQUIRE(K <= L, "kmeans_fit: K > L");
    const bool is_soft = !isnan(beta);
    AMSS_REQUIRE(is_soft ? soft != nullptr : labels != nullptr, "kmeans_fit: output for the chosen mode is null");
    KmWs w = km_ws(workspace, B, L, E, K, tries);
    if (workspace_bytes < w.total) {
        set_error("kmeans_fit: workspace %zu < %zu", workspace_bytes, w.total);
        return AMSS_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int G = km_group(B, L, E);
    // the update / inertia passes run on the tensor cores when the shape allows (kmeans_tc.cu); AMSS_KMEANS_SIMT=1 forces
    // the fp32 SIMT kernels of this file (the parity path for every other shape)
    const bool use_tc = kmeans_tc_supported(E, K, tries, is_soft, notsilent != nullptr);
    const size_t smem = ((size_t)KM_TILE * (E + 1) + (size_t)K * E) * 4;
    if (is_soft) AMSS_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else AMSS_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // one group of mixtures at a time: every pass over the group's X after the first one is served by the L2
    for (int b0 = 0; b0 < B; b0 += G) {
        const int Bg = std::min(G, B - b0);
        const float* Xg = X + (size_t)b0 * L * E;
        float* centg = w.cent + (size_t)b0 * tries * K * E;
        // (the inertia pass of the tensor-core kernel counts a CTA's tiles in 16-bit counters)
        const bool tcp = use_tc && (((uintptr_t)Xg & 15) == 0) && (L + 127) / 128 * Bg / kNumSMs < 65000;
        const int chunks = tcp ? km_chunks_tc(Bg, L) : km_chunks(Bg, L);
        AMSS_LAUNCH(kmeans_gather_init_kernel, 64, 256, 0, st, Xg, init_idx + (size_t)b0 * tries * K, Bg, L, E, K, tries,
                    normalize_input, centg);
        int rc;
        const float* inertia_part = w.part;
        if (tcp) {
            // update passes chained through their partial sums (no finalize launch in between); the inertia pass reduces the
            // last ones and leaves the final centroids of every try in centg for the selection
            float* cur = w.part;
            const float* prev = nullptr;
            for (int it = 0; it < iters; ++it) {
                rc = kmeans_pass_tc(Xg, centg, prev, nullptr, Bg, L, K, tries, chunks, normalize_input, KM_UPDATE, cur, st);
                if (rc != AMSS_OK) return rc;
                prev = cur;
                cur = cur == w.part ? w.part2 : w.part;
            }
            rc = kmeans_pass_tc(Xg, centg, prev, centg, Bg, L, K, tries, chunks, normalize_input, KM_INERTIA, cur, st);
            if (rc != AMSS_OK) return rc;
            inertia_part = cur;
        } else {
            for (int it = 0; it < iters; ++it) {
                rc = is_soft ? launch_pass<KM_UPDATE, 1>(Xg, centg, notsilent, Bg, B, b0, normalize_input, L, E, K, tries, beta, w.part, st)
                             : launch_pass<KM_UPDATE, 0>(Xg, centg, notsilent, Bg, B, b0, normalize_input, L, E, K, tries, beta, w.part, st);
                if (rc != AMSS_OK) return rc;
                AMSS_LAUNCH(kmeans_finalize_kernel, 64, 256, 0, st, w.part, Bg, chunks, tries, K, E, centg);
            }
            rc = is_soft ? launch_pass<KM_INERTIA, 1>(Xg, centg, notsilent, Bg, B, b0, normalize_input, L, E, K, tries, beta, w.part, st)
                         : launch_pass<KM_INERTIA, 0>(Xg, centg, notsilent, Bg, B, b0, normalize_input, L, E, K, tries, beta, w.part, st);
            if (rc != AMSS_OK) return rc;
        }
        float* centroids_g = centroids + (size_t)b0 * K * E;
        int tcG = 0, tcP = 0;
        if (tcp) kmeans_tc_geometry(Bg, L, &tcG, &tcP);
        AMSS_LAUNCH(kmeans_select_kernel, Bg, 128, 0, st, inertia_part, centg, Bg, chunks, tries, K, E,
                    inertia ? inertia + (size_t)b0 * tries : nullptr, best_try + b0, centroids_g, (L + 127) / 128,
                    (int64_t)Bg * ((L + 127) / 128), tcG);
        // final labels: un-gated X if assign_at_end (Kmeans_2.py:106-107), else the best try's gated labels
        const uint8_t* ns_final = assign_at_end ? nullptr : notsilent;
        dim3 grid((unsigned)std::min<int64_t>((L + KM_TILE - 1) / KM_TILE, (int64_t)std::max(1, 4 * kNumSMs / Bg)), Bg);
        if (!is_soft && ns_final == nullptr && E == KA_E && (((uintptr_t)Xg & 15) == 0)) {
            dim3 gridf((unsigned)std::min<int64_t>((L + KA_TILE - 1) / KA_TILE, (int64_t)std::max(1, 4 * kNumSMs / Bg)), Bg);
            AMSS_LAUNCH(kmeans_assign_hard40_kernel, gridf, KA_TILE, 0, st, Xg, centroids_g, normalize_input, L, K,
                        labels + (size_t)b0 * L);
        } else if (is_soft) {
            AMSS_LAUNCH(kmeans_assign_kernel<1>, grid, KM_THREADS, smem, st, Xg, centroids_g, ns_final, best_try + b0, B, b0,
                        normalize_input, L, E, K, tries, beta, (int*)nullptr, soft + (size_t)b0 * L * K);
        } else {
            AMSS_LAUNCH(kmeans_assign_kernel<0>, grid, KM_THREADS, smem, st, Xg, centroids_g, ns_final, best_try + b0, B, b0,
                        normalize_input, L, E, K, tries, beta, labels + (size_t)b0 * L, (float*)nullptr);
        }
    }
    return AMSS_OK;
}

extern "C" int amss_kmeans_silence_mask(const float* latent, int B, int64_t L, float threshold, uint8_t* notsilent,
                                        void* workspace, void* stream) {
    AMSS_REQUIRE(latent && notsilent && workspace, "kmeans_silence_mask: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    float* rowmax = (float*)workspace;  // B floats
    AMSS_LAUNCH(row_max_kernel, B, 256, 0, st, latent, L, rowmax);
    AMSS_LAUNCH(silence_mask_kernel, 2 * kNumSMs, 256, 0, st, latent, rowmax, L, (int64_t)B * L, threshold, notsilent);
    return AMSS_OK;
}

extern "C" int amss_apply_masks(const float* X_input, const int32_t* labels, const float* soft, int B, int S,
                                int64_t TF, float* separated, void* stream) {
    AMSS_REQUIRE(X_input && separated && ((labels != nullptr) != (soft != nullptr)),
                 "apply_masks: need exactly one of labels / soft");
    AMSS_LAUNCH(apply_masks_kernel, 4 * kNumSMs, 256, 0, (cudaStream_t)stream, X_input, labels, soft, B, S, TF,
                separated);
    return AMSS_OK;
}
