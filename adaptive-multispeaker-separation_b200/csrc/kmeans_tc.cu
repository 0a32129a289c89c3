// k-means pass on the 5th-gen tensor cores (models/Kmeans_2.py:145-188): one pass over X labels every point against the
// centroids of ALL tries and accumulates the per-(try, cluster) sums -- both as small GEMMs whose operands are bf16
// SPLITS of the fp32 data (x = hi + mid + lo, three bf16 terms = 24 mantissa bits, exact), so the results carry fp32
// accuracy although the products run on tcgen05:
//   phase 1  dot[p][(t,k)] = sum_e x[p][e] c[(t,k)][e]   as 6 split-pair products (hi.hi, hi.mid, mid.hi, hi.lo, mid.mid,
//            lo.hi; the dropped terms are below 2^-24 |x||c|), M = 128 points, N = 32 centroid columns, K = 48 per split;
//            d^2 = |x|^2 - 2 dot + |c|^2, arg-min over the clusters of each try in the epilogue thread that owns the
//            point (TMEM lane = point: all tries x clusters of a point sit in one thread's registers, no shuffles);
//   phase 2  D2[feature row][(t,k)] = sum_p xsplit[p][feature] onehot[p][(t,k)]: the x splits are the MN-major A operand
//            (M = 128 rows: hi 0-39, mid 40-79, lo 80-119, the ones column 120 -> counts), the one-hot tile (exact in bf16)
//            the MN-major B operand (N = 32), K = points; the accumulator (32 TMEM columns) stays resident for all tiles
//            of the CTA and is reduced to sum[(t,k)][e] = (hi + mid) + lo once at the end.
// Phase 1 reads its A operand (the x splits, K-major) from TENSOR MEMORY: the loader thread that owns a point's row also parks
// the packed splits in its TMEM lane (tcgen05.st); with both operands in shared memory every one of the 18 small-N MMAs of a
// tile paid ~100 clk of operand fetch, from TMEM they issue at ~36 clk.  (Phase 2's 8 MMAs still pay it: their A operand
// would have to sit feature-per-lane in TMEM, and a thread can only write its own lane.)
// Raw tiles (128 points x 160 B, contiguous in X) arrive by ONE TMA bulk copy each into a 4-deep ring (a producer warp; 128
// row copies into a padded pitch cost ~50 clk each on the TMA engine).  ONE loader thread owns a row: conflict-free LDS.128
// of the dense rows (the upper half of a quarter-warp reads its chunks rotated by one), normalise (tf.nn.l2_normalize,
// Kmeans_2.py:40-41), split, 16 stores of 16 bytes in the canonical core-matrix layout (phase-2 operand) + 18 tcgen05.st
// (phase-1 operand); the two loader groups (4 warps each) take alternate tiles; no CTA-wide barrier anywhere in the loop.
// The epilogue thread owns a point: |c|^2 in registers, branch-free arg-min over the clusters of every try, one-hot row.
// Measured per tile (tools/kmeans_tile_profile.py, profiles/r02l_kmeans_tile_profile.txt): 2900 clk -> 1740 clk; what is left
// is the tensor pipe's fixed cost per small MMA (18 x ~36 + 8 x ~100 clk) and the SM's instruction issue.
// Warp roles: 0-7 loaders, 8 MMA issuer (+TMEM), 9-12 epilogue, 13 TMA producer.
// Restrictions (anything else takes the SIMT kernels of kmeans.cu): hard assignments, E == 40, tries*K <= 32, no silence
// gate.
#include "common.cuh"
#include "tc.cuh"
#include "kmeans_pieces.cuh"
#include <algorithm>
#include <cstdlib>

namespace amss {
namespace {

using namespace tc;

constexpr int KT_THREADS = 448;
constexpr int KT_NRAW = 4;                       // raw fp32 tiles in flight (one TMA bulk copy each, dense 160-byte rows)
constexpr int KT_E = 40;
constexpr int KT_NCH = 18;                       // 16-byte units per point: 3 splits x 6 chunks of 8 features (48 >= E + 1)
constexpr uint32_t KT_RG = 16 * 128;             // phase-2 operand tile: bytes per group of 8 points = 16 feature chunks: hi 0-4,
                                                 // mid 0-4, lo 0-4 (5 chunks of 8 features per split) + the ones chunk -> M = 128 rows
constexpr uint32_t KT_X3 = 16 * KT_RG;           // ... of 128 points
constexpr uint32_t KT_OHG = 4 * 128;             // one-hot tile: 4 units (32 columns) per point, 8 points per group
constexpr uint32_t KT_OH = 16 * KT_OHG;
constexpr int KT_N1 = 32;                        // centroid columns (tries * K <= 32)
constexpr int KT_ONES = 120;                     // row of the phase-2 product that holds the counts (chunk 15, feature 40)
constexpr uint32_t KT_C3 = KT_NCH * (KT_N1 / 8) * 128;
constexpr uint32_t KT_ACOL = 256;                // TMEM: D1 ring [0,96), D2 [96,128), A ring [256, 256 + 72*NBUF)
constexpr int KT_NBUF = 3;                      // operand / one-hot / distance-accumulator ring: the loaders (most of the
                                                 // instructions) run two tiles ahead of the MMA -> epilogue -> MMA chain

enum { KT_UPDATE = 0, KT_INERTIA = 1 };

// Optional tile profile (diagnostics, amss_debug_kmeans_profile): clock64() stamps of CTA 0, tiles [KT_PROF_T0, +4), 8 slots
// per tile and role (0 loader thread 0, 1 MMA issuer, 2 first epilogue thread): dev_buf[(role*4 + tile)*8 + slot].
long long* g_kt_prof = nullptr;
constexpr uint32_t KT_PROF_T0 = 8;
#define KPROF(role, k) do { if (prof && i >= KT_PROF_T0 && i < KT_PROF_T0 + 4) prof[((role) * 4 + (i - KT_PROF_T0)) * 8 + (k)] = clock64(); } while (0)

struct KtParams {
    long long* prof;
    const float* X;        // [Bg][L][E]
    const float* cent;     // [Bg][tries][K][E]  (read when prev_part is null)
    const float* prev_part;  // UPDATE partial sums of the previous pass [Bg][chunks][tries*K][E+1]: the centroids are reduced from
                           // them in the prologue (the arithmetic of kmeans_finalize_kernel, no launch in between) ...
    float* cent_out;       // ... and written here by the first CTA of every mixture (may be null)
    float* part;           // UPDATE: [Bg][chunks][tries*K][E+1]; INERTIA: [Bg][chunks][tries*K][2]
    int64_t L, ntiles, Tt;   // points and 128-point tiles per mixture, tiles of the group
    int K, tries, chunks, normalize;   // chunks = row pitch of `part` in pieces per mixture (>= the most pieces any mixture has)
    int G;                 // CTAs = equal contiguous ranges of the group's flat tile list (kmeans_pieces.cuh)
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
    // x3_full[NBUF], x3_empty[NBUF], d1_full[NBUF], d1_empty[NBUF], oh_full[NBUF], oh_empty[NBUF], done, raw_full[NRAW], raw_empty[NRAW]
    __shared__ __align__(8) uint64_t bars[6 * KT_NBUF + 1 + 2 * KT_NRAW];
    __shared__ uint32_t tmem_base_s;
    __shared__ float cc_s[KT_N1];
    __shared__ float xx_s[KT_NBUF][128];
    __shared__ float fin_s[4][32][2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int TK = p.tries * KC;
    uint8_t* x3_s = smem;                                        // [NBUF][KT_X3]
    uint8_t* oh_s = x3_s + KT_NBUF * KT_X3;                      // [NBUF][KT_OH]  (the MMA also reads the 1.5 KB behind a tile:
    uint8_t* c3_s = oh_s + KT_NBUF * KT_OH;                      //  rows >= 32 of the M = 128 product, never used)
    float* vs = reinterpret_cast<float*>(c3_s + KT_C3);          // [NRAW][128][KT_E] raw fp32 tiles
    const uint32_t x3_full = smem_u32(&bars[0]), x3_empty = x3_full + 8 * KT_NBUF, d1_full = x3_full + 16 * KT_NBUF,
                   d1_empty = x3_full + 24 * KT_NBUF, oh_full = x3_full + 32 * KT_NBUF, oh_empty = x3_full + 40 * KT_NBUF,
                   done = x3_full + 48 * KT_NBUF, raw_full = done + 8, raw_empty = raw_full + 8 * KT_NRAW;
    if (tid == 0) {
        for (int i = 0; i < KT_NBUF; ++i) {
            mbar_init(x3_full + 8 * i, 128); mbar_init(x3_empty + 8 * i, 1);
            mbar_init(d1_full + 8 * i, 1);   mbar_init(d1_empty + 8 * i, 128);
            mbar_init(oh_full + 8 * i, 128); mbar_init(oh_empty + 8 * i, 1);
        }
        mbar_init(done, 1);
        for (int i = 0; i < KT_NRAW; ++i) { mbar_init(raw_full + 8 * i, 1); mbar_init(raw_empty + 8 * i, 128); }
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc(smem_u32(&tmem_base_s), 512);
    for (uint32_t i = tid * 16; i < KT_NBUF * KT_OH; i += KT_THREADS * 16) *reinterpret_cast<uint4*>(oh_s + i) = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    __shared__ float cent_s[KT_N1 * KT_E];
    // this CTA's range of the group's flat tile list, walked mixture by mixture ("segments"; the rings and their mbarrier
    // phases run on through the segments: ibase = tiles of the earlier segments)
    const int64_t f0 = kt_start(blockIdx.x, p.Tt, p.G), f1 = kt_start(blockIdx.x + 1, p.Tt, p.G);
    uint32_t ibase = 0, seg = 0;
    for (int64_t fs = f0; fs < f1; ++seg) {
    const int b = (int)(fs / p.ntiles);
    const int64_t t0 = fs - (int64_t)b * p.ntiles;
    const int64_t fe = min(f1, (int64_t)(b + 1) * p.ntiles);
    const uint32_t ntile = (uint32_t)(fe - fs);
    fs = fe;
    int c_first, npieces;
    kt_pieces(b, p.ntiles, p.Tt, p.G, c_first, npieces);
    const int chunk = (int)blockIdx.x - c_first;                 // which piece of mixture b this segment is
    // centroids of this mixture: cent[m][e] = sum_pieces part_sum / sum_pieces part_cnt (Kmeans_2.py:158-165; pieces in
    // sequence, an empty cluster gives 0/0 = NaN as in the reference), or the given ones
    for (int idx = tid; idx < TK * KT_E; idx += KT_THREADS) {
        float v;
        if (p.prev_part) {
            const int m = idx / KT_E, e = idx - m * KT_E;
            float sacc = 0.f, cacc = 0.f;
            const float* q0 = p.prev_part + ((size_t)b * p.chunks * TK + m) * (KT_E + 1);
            const size_t qs = (size_t)TK * (KT_E + 1);
            int ch = 0;
            for (; ch + 6 <= npieces; ch += 6) {                 // six pieces' loads in flight, added in piece order
                float sv[6], cv[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) { sv[j] = q0[(ch + j) * qs + e]; cv[j] = q0[(ch + j) * qs + KT_E]; }
#pragma unroll
                for (int j = 0; j < 6; ++j) { sacc += sv[j]; cacc += cv[j]; }
            }
            for (; ch < npieces; ++ch) { sacc += q0[ch * qs + e]; cacc += q0[ch * qs + KT_E]; }
            v = sacc / cacc;
            if (chunk == 0 && p.cent_out) p.cent_out[(size_t)b * TK * KT_E + idx] = v;
        } else {
            v = p.cent[(size_t)b * TK * KT_E + idx];
        }
        cent_s[idx] = v;
    }
    __syncthreads();
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
                const float x = (n < TK && e < KT_E) ? cent_s[n * KT_E + e] : 0.f;
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
            for (int e = 0; e < KT_E; ++e) { const float x = cent_s[tid * KT_E + e]; a = fmaf(x, x, a); }
        cc_s[tid] = a;
    }
    fence_async_smem();
    __syncthreads();

    if (warp < 8) {
        // ================= loaders: tile -> fp32 rows -> normalise -> 3 bf16 splits in the operand layout =================
        // thread = (row r, group g): ONE thread owns a whole row (the norm needs no exchange and no barrier), the two groups of
        // four warps take alternate tiles.  (Two threads per row, each computing the row norm redundantly, cost 1100 issue
        // slots per tile and scheduler against the epilogue's 450: the SM's instruction issue bounds this kernel.)
        const int r = tid & 127, g = tid >> 7;
        long long* prof = (blockIdx.x == 0 && r == 0) ? p.prof : nullptr;
        const int rot = (r >> 2) & 1;
        for (uint32_t i = (g + ibase) & 1; i < ntile; i += 2) {  // group g takes the tiles whose running index is g mod 2
            const uint32_t ig = ibase + i;
            const uint32_t buf = ig % KT_NBUF, ph = (ig / KT_NBUF) & 1;
            const uint32_t slot = ig % KT_NRAW, rph = (ig / KT_NRAW) & 1;
            const int64_t p0 = (t0 + i) * 128;
            const bool valid = p0 + r < p.L;
            KPROF(0, 0);
            mbar_wait(raw_full + 8 * slot, rph);                 // the producer's bulk copy of this tile has landed
            KPROF(0, 1);
            // The tile lies dense (160-byte rows): rows r and r + 4 start in the same bank, so the upper half of every
            // quarter-warp reads its 16-byte chunks rotated by one (conflict-free LDS.128) and un-rotates in registers
            float x[48];
            const float* row = vs + ((size_t)slot * 128 + r) * KT_E;
            {
                float4 rv[KT_E / 4];
#pragma unroll
                for (int c = 0; c < KT_E / 4; ++c) {
                    int cc = c + rot; if (cc == KT_E / 4) cc = 0;
                    rv[c] = valid ? *reinterpret_cast<const float4*>(row + cc * 4) : make_float4(0.f, 0.f, 0.f, 0.f);   // rows past the end are not copied
                }
#pragma unroll
                for (int c = 0; c < KT_E / 4; ++c) {
                    const float4 a = rv[c], bq = rv[(c + KT_E / 4 - 1) % (KT_E / 4)];      // chunk c sits in rv[c] (rot 0) or rv[c - 1] (rot 1)
                    x[4 * c] = rot ? bq.x : a.x; x[4 * c + 1] = rot ? bq.y : a.y; x[4 * c + 2] = rot ? bq.z : a.z; x[4 * c + 3] = rot ? bq.w : a.w;
                }
            }
            // sums of squares as two sequential chains over features [0,24) and [24,40), added at the end
            float ss0 = 0.f, ss1 = 0.f;
#pragma unroll
            for (int e = 0; e < 24; ++e) ss0 = fmaf(x[e], x[e], ss0);
#pragma unroll
            for (int e = 24; e < KT_E; ++e) ss1 = fmaf(x[e], x[e], ss1);
            const float ss = ss0 + ss1;
            mbar_arrive(raw_empty + 8 * slot);
            KPROF(0, 2);
            float inv = 1.f;
            if (p.normalize) inv = rsqrtf(fmaxf(ss, 1e-12f));
            float xx0 = 0.f, xx1 = 0.f;
#pragma unroll
            for (int e = 0; e < 24; ++e) { x[e] *= inv; xx0 = fmaf(x[e], x[e], xx0); }
#pragma unroll
            for (int e = 24; e < KT_E; ++e) { x[e] *= inv; xx1 = fmaf(x[e], x[e], xx1); }
            const float xx = xx0 + xx1;
            x[KT_E] = valid ? 1.f : 0.f;                         // the ones column (feature 40): exact in bf16
#pragma unroll
            for (int e = KT_E + 1; e < 48; ++e) x[e] = 0.f;
            KPROF(0, 3);
            mbar_wait(x3_empty + 8 * buf, ph ^ 1);               // the MMAs of tile i-3 have finished with x3[buf]
            mbar_wait(d1_empty + 8 * buf, ph ^ 1);               // ... and its epilogue has read xx_s[buf]
            KPROF(0, 4);
            uint8_t* dst = x3_s + buf * KT_X3 + (size_t)(r >> 3) * KT_RG + (r & 7) * 16;
            const uint32_t acol = tmem + ((uint32_t)((warp & 3) * 32) << 16) + KT_ACOL + buf * 72;
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                float sp3[3][8];
#pragma unroll
                for (int j = 0; j < 8; ++j) split3(x[c * 8 + j], sp3[0][j], sp3[1][j], sp3[2][j]);
#pragma unroll
                for (int sp = 0; sp < 3; ++sp) {
                    // the padding chunk (features 40-47) carries the ones column in the hi split and zeros in the others:
                    // those zeros are written on the first use of a buffer only (nothing else touches them)
                    if (c == 5 && sp > 0 && ig >= KT_NBUF) continue;
                    const uint32_t w0 = pack_trunc(sp3[sp][0], sp3[sp][1]), w1 = pack_trunc(sp3[sp][2], sp3[sp][3]),
                                   w2 = pack_trunc(sp3[sp][4], sp3[sp][5]), w3 = pack_trunc(sp3[sp][6], sp3[sp][7]);
                    // phase-2 operand: feature chunk c of split sp -> chunk sp*5 + c, the ones chunk -> chunk 15
                    if (c < 5) *reinterpret_cast<uint4*>(dst + (sp * 5 + c) * 128) = make_uint4(w0, w1, w2, w3);
                    else if (sp == 0) *reinterpret_cast<uint4*>(dst + 15 * 128) = make_uint4(w0, w1, w2, w3);
                    // phase-1 A operand: TMEM lane = point, 24 columns per split (two k per column), chunk = 4 columns
                    tmem_st4(acol + sp * 24 + c * 4, w0, w1, w2, w3);
                }
            }
            KPROF(0, 5);
            tmem_st_wait();
            tc_fence_before();
            KPROF(0, 6);
            xx_s[buf][r] = xx;                                   // |x|^2 of the normalised row
            fence_async_smem();
            mbar_arrive(x3_full + 8 * buf);
            KPROF(0, 7);
        }
    } else if (warp == 8) {
        // ================= MMA issuer =================
        const uint32_t idesc1 = idesc_bf16(128, KT_N1, 0, 0), idesc2 = idesc_bf16(128, KT_N1, 1, 1);
        const bool leader = elect_one();
        const uint32_t cbase = smem_u32(c3_s);
        long long* prof = (blockIdx.x == 0 && leader) ? p.prof : nullptr;
        auto phase2 = [&](uint32_t j) {                           // sums of tile j: [x splits | ones]^T onehot  (M = 128 feature rows,
                                                                  // N = 32 columns, K = points: 8 small MMAs instead of 8 of N = 144)
            const uint32_t bj = (ibase + j) % KT_NBUF, pj = ((ibase + j) / KT_NBUF) & 1;
            mbar_wait(oh_full + 8 * bj, pj);
            tc_fence_after();
            { const uint32_t i = j + 1; KPROF(1, 4); }
            const uint32_t oa = smem_u32(oh_s + bj * KT_OH), xa = smem_u32(x3_s + bj * KT_X3);
            for (int kk = 0; kk < 8; ++kk) {
                const uint64_t ad = smem_desc(xa + kk * 2 * KT_RG, KT_RG, 128);
                const uint64_t bd = smem_desc(oa + kk * 2 * KT_OHG, KT_OHG, 128);
                if (leader) mma_bf16(tmem + KT_NBUF * KT_N1, ad, bd, idesc2, (j | (uint32_t)kk) != 0);
            }
            if (leader) { mma_commit(x3_empty + 8 * bj); mma_commit(oh_empty + 8 * bj); }
            { const uint32_t i = j + 1; KPROF(1, 5); }
        };
        for (uint32_t i = 0; i < ntile; ++i) {
            const uint32_t buf = (ibase + i) % KT_NBUF, ph = ((ibase + i) / KT_NBUF) & 1;
            KPROF(1, 0);
            mbar_wait(x3_full + 8 * buf, ph);
            KPROF(1, 1);
            mbar_wait(d1_empty + 8 * buf, ph ^ 1);               // the epilogue of tile i-2 has drained D1[buf]
            tc_fence_after();
            KPROF(1, 2);
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
            KPROF(1, 3);
            if (MODE == KT_UPDATE && i > 0) phase2(i - 1);
        }
        if (MODE == KT_UPDATE && ntile) { phase2(ntile - 1); if (leader) mma_commit(done); }
    } else if (warp == 13) {
        // ================= producer: one TMA bulk copy per tile (128 row copies into a padded pitch were tried: ~50 clk per
        // copy on the TMA engine, 6300 clk per tile) =================
        for (uint32_t i = 0; i < ntile; ++i) {
            const uint32_t slot = (ibase + i) % KT_NRAW, rph = ((ibase + i) / KT_NRAW) & 1;
            const int64_t p0 = (t0 + i) * 128;
            const int np = (int)min((int64_t)128, p.L - p0);
            mbar_wait(raw_empty + 8 * slot, rph ^ 1);            // every loader has read its row of tile i - NRAW
            if (lane == 0) {                                     // the tile is one contiguous block of np * 160 bytes
                const float* src = p.X + ((size_t)b * p.L + p0) * KT_E;
                mbar_expect_tx(raw_full + 8 * slot, (uint32_t)np * KT_E * 4);
                bulk_g2s(smem_u32(vs + (size_t)slot * 128 * KT_E), src, (uint32_t)np * KT_E * 4, raw_full + 8 * slot);
            }
            __syncwarp();
        }
    } else {
        // ================= epilogue: thread = point =================
        const int q = warp & 3, r = q * 32 + lane;
        // INERTIA: per-thread sums of the selected squared distances and counts per column (this thread's points, in tile
        // order), reduced over the lanes / warps once at the end
        constexpr int TMAXI = MODE == KT_INERTIA ? KT_N1 / KC : 1;
        float tot[TMAXI][KC];
        uint32_t cpk[(KT_N1 + 1) / 2];                           // counts, two 16-bit counters per word (ntile < 65536, checked by the host)
#pragma unroll
        for (int t = 0; t < TMAXI; ++t)
#pragma unroll
            for (int k = 0; k < KC; ++k) tot[t][k] = 0.f;
#pragma unroll
        for (int j = 0; j < (KT_N1 + 1) / 2; ++j) cpk[j] = 0u;
        // |c|^2 of every column stays in registers for all tiles (columns >= tries*K hold zero centroids: computed, masked off):
        // the epilogue is branch-free -- per-try uniform branches and shared-memory reads of |c|^2 made a serial
        // S2UR -> LDS -> FADD chain per try (1800-2300 clk of the 2900 clk a tile took, tools/kmeans_tile_profile.py)
        constexpr int NCC = MODE == KT_UPDATE ? KT_N1 : 1;       // (the inertia pass keeps its sums in registers instead)
        float ccr[NCC];
#pragma unroll
        for (int m = 0; m < NCC; ++m) ccr[m] = cc_s[m];
        const uint32_t colmask = TK >= 32 ? 0xFFFFFFFFu : ((1u << TK) - 1u);
        long long* prof = (blockIdx.x == 0 && r == 0) ? p.prof : nullptr;
        for (uint32_t i = 0; i < ntile; ++i) {
            const uint32_t buf = (ibase + i) % KT_NBUF, ph = ((ibase + i) / KT_NBUF) & 1;
            const int64_t p0 = (t0 + i) * 128;
            const bool valid = p0 + r < p.L;
            KPROF(2, 0);
            mbar_wait(x3_full + 8 * buf, ph);                     // acquire |x|^2 written by the loaders
            KPROF(2, 1);
            mbar_wait(d1_full + 8 * buf, ph);
            tc_fence_after();
            KPROF(2, 2);
            uint32_t v[KT_N1];
            const float xx = xx_s[buf][r];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + buf * KT_N1, v);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(d1_empty + 8 * buf);
            KPROF(2, 3);
            uint32_t mask = 0;                                   // bit m = 1: this point belongs to column m = t*KC + k
            constexpr int TMAX = KT_N1 / KC;
            float dmin[TMAX];
#pragma unroll
            for (int t = 0; t < TMAX; ++t) {
                float bd = 0.f;
                uint32_t bit = 0;
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const float cc = MODE == KT_UPDATE ? ccr[(t * KC + k) % NCC] : cc_s[t * KC + k];
                    // clamp rounding negatives, but keep NaN (an empty cluster's 0/0 centroid): fmaxf(NaN, 0) = 0 would make
                    // the NaN cluster everybody's nearest; `NaN < bd` is false, as in the SIMT kernels and tf.argmin
                    const float dr = xx - 2.f * __uint_as_float(v[t * KC + k]) + cc;
                    const float d2 = dr < 0.f ? 0.f : dr;
                    if (k == 0 || d2 < bd) { bd = d2; bit = 1u << (t * KC + k); }   // first minimum wins ties (tf.argmin)
                }
                dmin[t] = bd;
                mask |= bit;
            }
            mask = valid ? (mask & colmask) : 0u;
            KPROF(2, 4);
            if (MODE == KT_UPDATE) {
                mbar_wait(oh_empty + 8 * buf, ph ^ 1);           // phase 2 of tile i-2 has finished with oh[buf]
                KPROF(2, 5);
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
                KPROF(2, 6);
            } else {
                // inertia: per column m the sum of the selected squared distances and the count
#pragma unroll
                for (int t = 0; t < TMAX; ++t)
#pragma unroll
                    for (int k = 0; k < KC; ++k) {
                        const bool mine = (mask >> (t * KC + k)) & 1u;
                        tot[t % TMAXI][k] += mine ? dmin[t] : 0.f;
                    }
#pragma unroll
                for (int j = 0; j < (KT_N1 + 1) / 2; ++j)        // bits 2j, 2j+1 of the mask -> the two halves of counter word j
                    cpk[j] += ((mask >> (2 * j)) & 1u) | (((mask >> (2 * j + 1)) & 1u) << 16);
            }
        }
        if (MODE == KT_UPDATE) {
            // final: TMEM lane f = feature row (hi 0-39, mid 40-79, lo 80-119, ones 120), column m = (try, cluster): through
            // shared memory (the fp32 staging block is free by now), then sum[m][e] = (hi + mid) + lo, coalesced store
            float* fin = vs;                                     // [128][33]
            if (ntile) { mbar_wait(done, seg & 1); tc_fence_after(); }
            {
                uint32_t v[KT_N1];
                if (ntile) { tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + KT_NBUF * KT_N1, v); tmem_ld_wait(); }
                else {
#pragma unroll
                    for (int j = 0; j < KT_N1; ++j) v[j] = 0u;
                }
#pragma unroll
                for (int j = 0; j < KT_N1; ++j) fin[r * 33 + j] = __uint_as_float(v[j]);
            }
            kt_sync(2, 128);
            float* dst = p.part + ((size_t)b * p.chunks + chunk) * TK * (KT_E + 1);
            for (int idx = r; idx < TK * (KT_E + 1); idx += 128) {
                const int m = idx / (KT_E + 1), e = idx - m * (KT_E + 1);
                dst[idx] = e < KT_E ? (fin[e * 33 + m] + fin[(KT_E + e) * 33 + m]) + fin[(2 * KT_E + e) * 33 + m]
                                    : fin[KT_ONES * 33 + m];
            }
        } else {
            // lanes in the fixed butterfly order of warp_sum, then the four warps in a fixed order
            float mytot = 0.f, mycnt = 0.f;
#pragma unroll
            for (int t = 0; t < TMAXI; ++t)
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const int m = t * KC + k;
                    const float s = warp_sum(tot[t][k]);
                    const float c = warp_sum((float)((cpk[m / 2] >> (16 * (m & 1))) & 0xFFFFu));   // exact: integers < 2^24
                    if (lane == m) { mytot = s; mycnt = c; }
                }
            fin_s[q][lane][0] = mytot; fin_s[q][lane][1] = mycnt;
            kt_sync(2, 128);
            if (q == 0 && lane < TK) {
                float* dst = p.part + (((size_t)b * p.chunks + chunk) * TK + lane) * 2;
                dst[0] = fin_s[1][lane][0] + fin_s[2][lane][0] + fin_s[3][lane][0] + fin_s[0][lane][0];
                dst[1] = fin_s[1][lane][1] + fin_s[2][lane][1] + fin_s[3][lane][1] + fin_s[0][lane][1];
            }
        }
    }
    // end of the segment: every role has finished with the centroid operand, |c|^2 and the staging block
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    ibase += ntile;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 512);
}

constexpr size_t KT_SMEM = KT_NBUF * (size_t)KT_X3 + KT_NBUF * (size_t)KT_OH + KT_C3 + (size_t)KT_NRAW * 128 * KT_E * 4 + 2048;

}  // namespace

void kmeans_tc_set_profile(long long* dev_buf) { g_kt_prof = dev_buf; }

bool kmeans_tc_supported(int E, int K, int tries, bool soft, bool gated) {
    if (const char* e = getenv("AMSS_KMEANS_SIMT")) { if (atoi(e)) return false; }
    return !soft && !gated && E == KT_E && K >= 2 && K <= 4 && tries >= 1 && tries * K <= KT_N1;
}

// CTAs of a pass over a group of Bg mixtures and the most pieces any mixture is cut into (the row pitch of `part`)
void kmeans_tc_geometry(int Bg, int64_t L, int* G_out, int* pmax_out) {
    const int64_t nt = (L + 127) / 128, Tt = nt * Bg;
    const int G = (int)std::min<int64_t>(kNumSMs, Tt);
    int pmax = 1;
    for (int b = 0; b < Bg; ++b) {
        int c0, n;
        kt_pieces(b, nt, Tt, G, c0, n);
        pmax = std::max(pmax, n);
    }
    *G_out = G; *pmax_out = pmax;
}

int kmeans_pass_tc(const float* X, const float* cent, const float* prev_part, float* cent_out, int Bg, int64_t L, int K, int tries,
                   int chunks, int normalize, int mode, float* part, cudaStream_t st) {
    KtParams p;
    int G, pmax;
    kmeans_tc_geometry(Bg, L, &G, &pmax);
    if (chunks < pmax) { set_error("kmeans_pass_tc: part pitch %d < %d pieces", chunks, pmax); return AMSS_ERR_INVALID_ARG; }
    p.G = G; p.Tt = (int64_t)Bg * ((L + 127) / 128);
    p.prev_part = prev_part; p.cent_out = cent_out;
    p.prof = mode == KT_UPDATE ? g_kt_prof : nullptr;
    p.X = X; p.cent = cent; p.part = part; p.L = L; p.ntiles = (L + 127) / 128; p.K = K; p.tries = tries; p.chunks = chunks;
    p.normalize = normalize;
#define KT_LAUNCH(MODE_, KC_)                                                                                                   \
    do {                                                                                                                        \
        AMSS_CUDA(cudaFuncSetAttribute((kmeans_pass_tc_kernel<MODE_, KC_>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KT_SMEM)); \
        AMSS_LAUNCH((kmeans_pass_tc_kernel<MODE_, KC_>), G, KT_THREADS, KT_SMEM, st, p);                                        \
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
