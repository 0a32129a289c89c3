// k-means pass on the 5th-gen tensor cores (models/Kmeans_2.py:145-188): one pass over X labels every point against the
// centroids of ALL tries and accumulates the per-(try, cluster) sums -- both as small GEMMs whose operands are bf16
// SPLITS of the fp32 data (x = hi + mid + lo, three bf16 terms = 24 mantissa bits, exact), so the results carry fp32
// accuracy although the products run on tcgen05:
//   phase 1  dot[p][(t,k)] = sum_e x[p][e] c[(t,k)][e]   as 6 split-pair products (hi.hi, hi.mid, mid.hi, hi.lo, mid.mid,
//            lo.hi; the dropped terms are below 2^-24 |x||c|), M = 128 points, N = 32 centroid columns, K = 48 per split;
//            d^2 = |x|^2 - 2 dot + |c|^2, arg-min over the clusters of each try in the epilogue thread that owns the
//            point (TMEM lane = point: all tries x clusters of a point sit in one thread's registers, no shuffles);
//   phase 2  sum[(t,k)][e] = sum_p onehot[p][(t,k)] x[p][e]  with the one-hot matrix (exact in bf16) as the MN-major A
//            operand and the SAME operand tile of x splits as the MN-major B operand (K = points); a ones column in the
//            tile's padding yields the counts.  The accumulator stays in TMEM for all tiles of the CTA.
// Phase 1 reads its A operand (the x splits, K-major) from TENSOR MEMORY: the loader thread that owns a point's (half) row
// also parks the packed splits in its TMEM lane (tcgen05.st); with both operands in shared memory every one of the 18
// small-N MMAs of a tile paid ~100 clk of operand fetch (5000 clk per tile measured, unchanged by more loader warps or a
// deeper ring), from TMEM they issue at ~36 clk.
// The operand tile of 128 points is built once per tile: coalesced float4 fetch (software-pipelined in registers) ->
// fp32 rows in shared memory -> the row owner normalises (tf.nn.l2_normalize, Kmeans_2.py:40-41), splits and writes 18
// 16-byte units (3 splits x 48 features) in the canonical core-matrix layout: read K-major by phase 1 (LBO = 128,
// SBO = 2304) and MN-major by phase 2 (LBO = 2304, SBO = 128).
// Warp roles: 0-7 loaders (a point's row is split between two threads: features [0,24) and [24,48), because one warp per
// scheduler left the normalise / split chain latency-bound: ncu 14 % active warps, 5000 clk per tile), 8 MMA issuer
// (+TMEM), 9-12 epilogue.  HBM/L2-bound by design: 4E bytes per point and pass.
// Restrictions (anything else takes the SIMT kernels of kmeans.cu): hard assignments, E == 40, tries*K <= 32, no silence
// gate.
#include "common.cuh"
#include "tc.cuh"
#include <algorithm>
#include <cstdlib>

namespace amss {
namespace {

using namespace tc;

constexpr int KT_THREADS = 416;
constexpr int KT_LOADERS = 256;
constexpr int KT_E = 40;
constexpr int KT_PITCH = 44;                     // fp32 staging pitch: LDS.128 of a row per thread is conflict-free
constexpr int KT_NCH = 18;                       // 16-byte units per point: 3 splits x 6 chunks of 8 features (48 >= E + 1)
constexpr uint32_t KT_RG = KT_NCH * 128;         // bytes per group of 8 points
constexpr uint32_t KT_X3 = 16 * KT_RG;           // operand tile of 128 points
constexpr uint32_t KT_OHG = 4 * 128;             // one-hot tile: 4 units (32 columns) per point, 8 points per group
constexpr uint32_t KT_OH = 16 * KT_OHG;
constexpr int KT_N1 = 32;                        // centroid columns (tries * K <= 32)
constexpr int KT_N2 = 144;                       // 3 x 48 feature columns
constexpr uint32_t KT_C3 = KT_NCH * (KT_N1 / 8) * 128;
constexpr uint32_t KT_ACOL = 256;                // TMEM: D1 ring [0,96), D2 [96,240), A ring [256, 256 + 72*NBUF)
constexpr int KT_NBUF = 3;                      // operand / one-hot / distance-accumulator ring: the loaders (most of the
                                                 // instructions) run two tiles ahead of the MMA -> epilogue -> MMA chain

enum { KT_UPDATE = 0, KT_INERTIA = 1 };

struct KtParams {
    const float* X;        // [Bg][L][E]
    const float* cent;     // [Bg][tries][K][E]
    float* part;           // UPDATE: [Bg][chunks][tries*K][E+1]; INERTIA: [Bg][chunks][tries*K][2]
    int64_t L, ntiles;
    int K, tries, chunks, normalize;
};

__device__ __forceinline__ void kt_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// x -> (hi, mid, lo), each representable in bf16, with hi + mid + lo == x exactly: hi and mid by truncation of the fp32
// mantissa (one LOP each), the remainders by exact subtractions; lo has at most 8 significant bits left
__device__ __forceinline__ void split3(float x, float& hi, float& mid, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFF0000u);
    const float r1 = x - hi;
    mid = __uint_as_float(__float_as_uint(r1) & 0xFFFF0000u);
    lo = r1 - mid;
}
// two floats that are exactly representable in bf16 -> one packed word (byte permute, no conversion)
__device__ __forceinline__ uint32_t pack_trunc(float lo, float hi) {
    return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632);
}

// KC = clusters per try (compile time: column m = t*KC + k is then a static register index in the epilogue)
template <int MODE, int KC>
__global__ void __launch_bounds__(KT_THREADS, 1) kmeans_pass_tc_kernel(KtParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // x3_full[NBUF], x3_empty[NBUF], d1_full[NBUF], d1_empty[NBUF], oh_full[NBUF], oh_empty[NBUF], done
    __shared__ __align__(8) uint64_t bars[6 * KT_NBUF + 1];
    __shared__ uint32_t tmem_base_s;
    __shared__ float cc_s[KT_N1];
    __shared__ float xx_s[KT_NBUF][128];
    __shared__ float ss_s[2][128];                 // per-half partial sums of squares of the raw row
    __shared__ float xn_s[2][128];                 // ... and of the normalised row
    __shared__ float fin_s[4][32][2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x / p.chunks, chunk = blockIdx.x % p.chunks;
    const int TK = p.tries * KC;
    uint8_t* x3_s = smem;                                        // [NBUF][KT_X3]
    uint8_t* oh_s = x3_s + KT_NBUF * KT_X3;                      // [NBUF][KT_OH]  (the MMA also reads the 1.5 KB behind a tile:
    uint8_t* c3_s = oh_s + KT_NBUF * KT_OH;                      //  rows >= 32 of the M = 128 product, never used)
    float* vs = reinterpret_cast<float*>(c3_s + KT_C3);          // [128][KT_PITCH] fp32 staging
    const uint32_t x3_full = smem_u32(&bars[0]), x3_empty = x3_full + 8 * KT_NBUF, d1_full = x3_full + 16 * KT_NBUF,
                   d1_empty = x3_full + 24 * KT_NBUF, oh_full = x3_full + 32 * KT_NBUF, oh_empty = x3_full + 40 * KT_NBUF,
                   done = x3_full + 48 * KT_NBUF;
    if (tid == 0) {
        for (int i = 0; i < KT_NBUF; ++i) {
            mbar_init(x3_full + 8 * i, KT_LOADERS); mbar_init(x3_empty + 8 * i, 1);
            mbar_init(d1_full + 8 * i, 1);   mbar_init(d1_empty + 8 * i, 128);
            mbar_init(oh_full + 8 * i, 128); mbar_init(oh_empty + 8 * i, 1);
        }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc(smem_u32(&tmem_base_s), 512);
    // centroid operand (K-major B, N = 32 columns n = t*K + k; unit (n, c) at (c*4 + n/8)*128 + (n%8)*16) + |c|^2
    for (int u = tid; u < KT_N1 * KT_NCH; u += KT_THREADS) {
        const int n = u % KT_N1, c = u / KT_N1, sp = c / 6, kc = c % 6;
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v2[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int e = kc * 8 + 2 * j + h;
                const float x = (n < TK && e < KT_E) ? p.cent[((size_t)b * TK + n) * KT_E + e] : 0.f;
                float s0, s1, s2;
                split3(x, s0, s1, s2);
                v2[h] = sp == 0 ? s0 : (sp == 1 ? s1 : s2);
            }
            w[j] = pack_trunc(v2[0], v2[1]);
        }
        *reinterpret_cast<uint4*>(c3_s + (size_t)(c * (KT_N1 / 8) + (n >> 3)) * 128 + (n & 7) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (tid < KT_N1) {
        float a = 0.f;
        if (tid < TK)
            for (int e = 0; e < KT_E; ++e) { const float x = p.cent[((size_t)b * TK + tid) * KT_E + e]; a = fmaf(x, x, a); }
        cc_s[tid] = a;
    }
    for (uint32_t i = tid * 16; i < KT_NBUF * KT_OH; i += KT_THREADS * 16) *reinterpret_cast<uint4*>(oh_s + i) = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int64_t t0 = p.ntiles * chunk / p.chunks, t1 = p.ntiles * (chunk + 1) / p.chunks;
    const uint32_t ntile = (uint32_t)(t1 - t0);

    if (warp < 8) {
        // ================= loaders: tile -> fp32 rows -> normalise -> 3 bf16 splits in the operand layout =================
        // thread = (row r, half h): h = 0 owns features [0,24) = chunks 0-2, h = 1 features [24,48) = chunks 3-5 (16 real
        // features + the ones column at feature 40)
        const int r = tid & 127, h = tid >> 7;
        float4 tilev[5];
        auto fetch = [&](int64_t tile) {
            const int64_t p0 = tile * 128;
            const int np = (int)min((int64_t)128, p.L - p0);
            const float4* src = reinterpret_cast<const float4*>(p.X + ((size_t)b * p.L + p0) * KT_E);
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int u = tid + KT_LOADERS * j;
                tilev[j] = u < np * 10 ? __ldg(src + u) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        if (ntile) fetch(t0);
        for (uint32_t i = 0; i < ntile; ++i) {
            const uint32_t buf = i % KT_NBUF, ph = (i / KT_NBUF) & 1;
            const int64_t p0 = (t0 + i) * 128;
            const bool valid = p0 + r < p.L;
            kt_sync(1, KT_LOADERS);                              // the rows of the previous tile have been consumed
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int u = tid + KT_LOADERS * j, rr = u / 10, c = u - rr * 10;
                *reinterpret_cast<float4*>(vs + rr * KT_PITCH + c * 4) = tilev[j];
            }
            if (i + 1 < ntile) fetch(t0 + i + 1);
            kt_sync(1, KT_LOADERS);
            float x[24];
            const float* row = vs + r * KT_PITCH + 24 * h;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (24 * h + 4 * c < KT_E) v = *reinterpret_cast<const float4*>(row + c * 4);
                x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
            }
            float ss = 0.f;
#pragma unroll
            for (int e = 0; e < 24; ++e) ss = fmaf(x[e], x[e], ss);
            ss_s[h][r] = ss;
            kt_sync(1, KT_LOADERS);
            float inv = 1.f;
            if (p.normalize) inv = rsqrtf(fmaxf(ss_s[0][r] + ss_s[1][r], 1e-12f));
            float xx = 0.f;
#pragma unroll
            for (int e = 0; e < 24; ++e) { x[e] *= inv; xx = fmaf(x[e], x[e], xx); }
            mbar_wait(x3_empty + 8 * buf, ph ^ 1);               // the MMAs of tile i-2 have finished with x3[buf]
            mbar_wait(d1_empty + 8 * buf, ph ^ 1);               // ... and its epilogue has read xx_s[buf]
            if (h == 1) x[16] = valid ? 1.f : 0.f;               // the ones column (feature 40): exact in bf16
            uint8_t* dst = x3_s + buf * KT_X3 + (size_t)(r >> 3) * KT_RG + (r & 7) * 16;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float sp3[3][8];
#pragma unroll
                for (int j = 0; j < 8; ++j) split3(x[c * 8 + j], sp3[0][j], sp3[1][j], sp3[2][j]);
#pragma unroll
                for (int sp = 0; sp < 3; ++sp) {
                    const uint32_t w0 = pack_trunc(sp3[sp][0], sp3[sp][1]), w1 = pack_trunc(sp3[sp][2], sp3[sp][3]),
                                   w2 = pack_trunc(sp3[sp][4], sp3[sp][5]), w3 = pack_trunc(sp3[sp][6], sp3[sp][7]);
                    *reinterpret_cast<uint4*>(dst + (sp * 6 + 3 * h + c) * 128) = make_uint4(w0, w1, w2, w3);   // phase-2 operand
                    // phase-1 A operand: TMEM lane = point, 24 columns per split (two k per column), chunk = 4 columns
                    tmem_st4(tmem + ((uint32_t)((warp & 3) * 32) << 16) + KT_ACOL + buf * 72 + sp * 24 + (3 * h + c) * 4, w0, w1, w2, w3);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            // |x|^2 of the normalised row: the two halves are added by the h = 1 thread after a second exchange
            xn_s[h][r] = xx;
            fence_async_smem();
            kt_sync(1, KT_LOADERS);
            if (h == 1) xx_s[buf][r] = xn_s[0][r] + xn_s[1][r];
            mbar_arrive(x3_full + 8 * buf);
        }
    } else if (warp == 8) {
        // ================= MMA issuer =================
        const uint32_t idesc1 = idesc_bf16(128, KT_N1, 0, 0), idesc2 = idesc_bf16(128, KT_N2, 1, 1);
        const bool leader = elect_one();
        const uint32_t cbase = smem_u32(c3_s);
        auto phase2 = [&](uint32_t j) {                           // sums of tile j: onehot^T [x splits | ones]
            const uint32_t bj = j % KT_NBUF, pj = (j / KT_NBUF) & 1;
            mbar_wait(oh_full + 8 * bj, pj);
            tc_fence_after();
            const uint32_t oa = smem_u32(oh_s + bj * KT_OH), xa = smem_u32(x3_s + bj * KT_X3);
            for (int kk = 0; kk < 8; ++kk) {
                const uint64_t ad = smem_desc(oa + kk * 2 * KT_OHG, KT_OHG, 128);
                const uint64_t bd = smem_desc(xa + kk * 2 * KT_RG, KT_RG, 128);
                if (leader) mma_bf16(tmem + KT_NBUF * KT_N1, ad, bd, idesc2, (j | (uint32_t)kk) != 0);
            }
            if (leader) { mma_commit(x3_empty + 8 * bj); mma_commit(oh_empty + 8 * bj); }
        };
        for (uint32_t i = 0; i < ntile; ++i) {
            const uint32_t buf = i % KT_NBUF, ph = (i / KT_NBUF) & 1;
            mbar_wait(x3_full + 8 * buf, ph);
            mbar_wait(d1_empty + 8 * buf, ph ^ 1);               // the epilogue of tile i-2 has drained D1[buf]
            tc_fence_after();
            uint32_t acc = 0;
#pragma unroll
            for (int term = 0; term < 6; ++term) {
                const int si = term == 0 ? 0 : term == 1 ? 0 : term == 2 ? 1 : term == 3 ? 0 : term == 4 ? 1 : 2;
                const int sj = term == 0 ? 0 : term == 1 ? 1 : term == 2 ? 0 : term == 3 ? 2 : term == 4 ? 1 : 0;
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) {
                    const uint64_t bd = smem_desc(cbase + (sj * 6 + 2 * kk) * (KT_N1 / 8) * 128, (KT_N1 / 8) * 128, 128);
                    if (leader) mma_bf16_ts(tmem + buf * KT_N1, tmem + KT_ACOL + buf * 72 + si * 24 + kk * 8, bd, idesc1, acc);
                    acc = 1;
                }
            }
            if (leader) {
                mma_commit(d1_full + 8 * buf);
                if (MODE == KT_INERTIA) mma_commit(x3_empty + 8 * buf);
            }
            if (MODE == KT_UPDATE && i > 0) phase2(i - 1);
        }
        if (MODE == KT_UPDATE && ntile) { phase2(ntile - 1); if (leader) mma_commit(done); }
    } else {
        // ================= epilogue: thread = point =================
        const int q = warp & 3, r = q * 32 + lane;
        float tot = 0.f, cnt = 0.f;                              // INERTIA: column m = lane of this warp's points
        for (uint32_t i = 0; i < ntile; ++i) {
            const uint32_t buf = i % KT_NBUF, ph = (i / KT_NBUF) & 1;
            const int64_t p0 = (t0 + i) * 128;
            const bool valid = p0 + r < p.L;
            mbar_wait(x3_full + 8 * buf, ph);                     // acquire |x|^2 written by the loaders
            mbar_wait(d1_full + 8 * buf, ph);
            tc_fence_after();
            uint32_t v[KT_N1];
            const float xx = xx_s[buf][r];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + buf * KT_N1, v);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(d1_empty + 8 * buf);
            uint32_t mask = 0;                                   // bit m = 1: this point belongs to column m = t*KC + k
            constexpr int TMAX = KT_N1 / KC;
            float dmin[TMAX];
#pragma unroll
            for (int t = 0; t < TMAX; ++t) {
                dmin[t] = 0.f;
                if (t < p.tries) {                               // warp-uniform
                    float bd = 0.f;
                    int bk = 0;
#pragma unroll
                    for (int k = 0; k < KC; ++k) {
                        const float d2 = fmaxf(xx - 2.f * __uint_as_float(v[t * KC + k]) + cc_s[t * KC + k], 0.f);
                        if (k == 0 || d2 < bd) { bd = d2; bk = k; }   // first minimum wins ties (tf.argmin)
                    }
                    dmin[t] = bd;
                    if (valid) mask |= 1u << (t * KC + bk);
                }
            }
            if (MODE == KT_UPDATE) {
                mbar_wait(oh_empty + 8 * buf, ph ^ 1);           // phase 2 of tile i-2 has finished with oh[buf]
                uint8_t* dst = oh_s + buf * KT_OH + (size_t)(r >> 3) * KT_OHG + (r & 7) * 16;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t bits = (mask >> (g * 8 + 2 * j)) & 3u;
                        w[j] = ((bits & 1u) ? 0x3F80u : 0u) | ((bits & 2u) ? 0x3F800000u : 0u);
                    }
                    *reinterpret_cast<uint4*>(dst + g * 128) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                fence_async_smem();
                mbar_arrive(oh_full + 8 * buf);
            } else {
                // inertia: per column m the sum of the selected squared distances and the count (fixed shuffle order)
#pragma unroll
                for (int t = 0; t < TMAX; ++t) {
                    if (t < p.tries) {
#pragma unroll
                        for (int k = 0; k < KC; ++k) {
                            const bool mine = (mask >> (t * KC + k)) & 1u;
                            const float s = warp_sum(mine ? dmin[t] : 0.f);
                            const float c = (float)__popc(__ballot_sync(0xffffffffu, mine));
                            if (lane == t * KC + k) { tot += s; cnt += c; }
                        }
                    }
                }
            }
        }
        if (MODE == KT_UPDATE) {
            // final: rows m = lane of TMEM quadrant 0 (warp 8) hold sum_p onehot[p][m] * [x splits | ones]
            if (q == 0) {
                float* dst = p.part + (((size_t)b * p.chunks + chunk) * TK + lane) * (KT_E + 1);
                float sum[KT_E];
#pragma unroll
                for (int e = 0; e < KT_E; ++e) sum[e] = 0.f;
                float count = 0.f;
                if (ntile) { mbar_wait(done, 0); tc_fence_after(); }
#pragma unroll
                for (int c = 0; c < KT_N2 / 16; ++c) {
                    uint32_t v[16];
                    if (ntile) { tmem_ld16(tmem + KT_NBUF * KT_N1 + c * 16, v); tmem_ld_wait(); }
                    else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = 0u;
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = c * 16 + j, e = n % 48;
                        if (e < KT_E) sum[e] += __uint_as_float(v[j]);
                        else if (n == KT_E) count = __uint_as_float(v[j]);
                    }
                }
                if (lane < TK) {
#pragma unroll
                    for (int e = 0; e < KT_E; ++e) dst[e] = sum[e];
                    dst[KT_E] = count;
                }
            }
        } else {
            fin_s[q][lane][0] = tot; fin_s[q][lane][1] = cnt;
            kt_sync(2, 128);
            if (q == 0 && lane < TK) {
                float* dst = p.part + (((size_t)b * p.chunks + chunk) * TK + lane) * 2;
                dst[0] = fin_s[1][lane][0] + fin_s[2][lane][0] + fin_s[3][lane][0] + fin_s[0][lane][0];
                dst[1] = fin_s[1][lane][1] + fin_s[2][lane][1] + fin_s[3][lane][1] + fin_s[0][lane][1];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 512);
}

constexpr size_t KT_SMEM = KT_NBUF * (size_t)KT_X3 + KT_NBUF * (size_t)KT_OH + KT_C3 + (size_t)128 * KT_PITCH * 4 + 2048;

}  // namespace

bool kmeans_tc_supported(int E, int K, int tries, bool soft, bool gated) {
    if (const char* e = getenv("AMSS_KMEANS_SIMT")) { if (atoi(e)) return false; }
    return !soft && !gated && E == KT_E && K >= 2 && K <= 4 && tries >= 1 && tries * K <= KT_N1;
}

int kmeans_pass_tc(const float* X, const float* cent, int Bg, int64_t L, int K, int tries, int chunks, int normalize, int mode,
                   float* part, cudaStream_t st) {
    KtParams p;
    p.X = X; p.cent = cent; p.part = part; p.L = L; p.ntiles = (L + 127) / 128; p.K = K; p.tries = tries; p.chunks = chunks;
    p.normalize = normalize;
#define KT_LAUNCH(MODE_, KC_)                                                                                                   \
    do {                                                                                                                        \
        AMSS_CUDA(cudaFuncSetAttribute((kmeans_pass_tc_kernel<MODE_, KC_>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KT_SMEM)); \
        AMSS_LAUNCH((kmeans_pass_tc_kernel<MODE_, KC_>), Bg * chunks, KT_THREADS, KT_SMEM, st, p);                              \
    } while (0)
    if (mode == KT_UPDATE) {
        if (K == 2) KT_LAUNCH(KT_UPDATE, 2); else if (K == 3) KT_LAUNCH(KT_UPDATE, 3); else KT_LAUNCH(KT_UPDATE, 4);
    } else {
        if (K == 2) KT_LAUNCH(KT_INERTIA, 2); else if (K == 3) KT_LAUNCH(KT_INERTIA, 3); else KT_LAUNCH(KT_INERTIA, 4);
    }
#undef KT_LAUNCH
    return AMSS_OK;
}

}  // namespace amss
