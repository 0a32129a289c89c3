// Dense GEMM on the 5th-gen tensor cores: C[M,N] (+)= op(A) op(B) (+ bias), bf16 operands, fp32 accumulation in
// TMEM, fp32 out.  Used (AMSS_PREC_BF16) for the hoisted BLSTM input projections, the embedding head
// (utils/ops.py:501-503) and every backward GEMM of those.
//
//   1. operands are plain ROW-MAJOR bf16 matrices (fp32 sources are converted once by convert_bf16_kernel; producers
//      that already hold bf16 call gemm_bf16 directly).  Each operand is either K-major (X[r][k] = src[r*ld + k]) or
//      MN-major (X[r][k] = src[k*ld + r]) -- the transposed GEMMs of the backward pass (dW = X^T dZ, dX = dZ W^T) need
//      no transpose pass, the UMMA descriptors read both majors;
//   2. TMA tensor maps (cuTensorMapEncodeTiled, SWIZZLE_128B) move 64-wide boxes global -> shared: K-major one box
//      {64 k, 128|256 rows}, MN-major boxes {64 mn, 64 k}; out-of-range rows / k are zero filled by the TMA unit, so
//      ragged M, N, K need no padding.  Descriptor fields pinned by tools/tma_probe.cu;
//   3. persistent CTAs (one per SM) walk (tile, k-split) work items; ONE thread feeds a 4-stage mbarrier ring
//      (48 KB per stage, ~190 KB of loads in flight per SM, no per-element load instructions);
//   4. one warp issues tcgen05.mma 128x256x16 (converged loop, elected lane);
//   5. TMEM accumulators are double buffered (2 x 256 columns): 8 epilogue warps drain tile i (bias, optional fused
//      l2-normalise over column groups, row remap, accumulate / split-K red.global.add) while tile i+1 is multiplied.
#include "common.cuh"
#include "tc.cuh"
#include <cuda.h>
#include <algorithm>
#include <cstdlib>

namespace amss {
namespace {

using namespace tc;

constexpr int GT_BM = 128, GT_BN = 256, GT_BK = 64;
constexpr int GT_THREADS = 320;                      // warp 0 loader, 1 MMA (+TMEM alloc), 2-9 epilogue
constexpr int GT_A_BYTES = GT_BM * GT_BK * 2;
constexpr int GT_EMAX = 48;                         // widest l2-normalised group the fused epilogue handles
constexpr int GT_STG_BYTES = 32 * GT_EMAX * 4;      // one warp's staging block: 32 rows x (32 | E) fp32 columns
// CTAS = 1: a CTA owns a 128 x 256 tile (B stage 32 KB, 4 stages).  CTAS = 2: a CTA PAIR (cluster of 2, cta_group::2)
// owns a 256 x 256 tile; each CTA stages its own 128 rows of A and HALF of B (16 KB) and the pair's MMA reads both
// halves -> a third less L2 -> SM operand traffic per flop, which is what bounds these GEMMs; 6 stages.
template <int CTAS> struct GtCfg {
    static constexpr int B_ROWS = GT_BN / CTAS;
    static constexpr int B_BYTES = B_ROWS * GT_BK * 2;
    static constexpr int STAGE_BYTES = GT_A_BYTES + B_BYTES;
    static constexpr int STAGES = CTAS == 1 ? 3 : 5;
    // ring + per-epilogue-warp output staging (TMA stores) + slack to align everything to the 1024-B swizzle atom
    static constexpr int SMEM = STAGES * STAGE_BYTES + 8 * GT_STG_BYTES + 1024;
};
constexpr int GT_MAX_STAGES = 6;

struct GtParams {
    CUtensorMap mapC;                 // fp32 output, box {32 cols, 32 rows} SWIZZLE_128B (plain) or {norm_E, 32} (fused normalise)
    int c_tma;                        // 1: the epilogue stores / reduces through mapC; 0: direct per-thread stores (row remap, odd ldc)
    CUtensorMap mapA, mapB;           // mapB box: 256 rows (CTAS = 1) or 128 rows (CTAS = 2) for a K-major B
    const float* bias;
    float* C;
    float* inv;                       // norm_E > 0: [M][N / norm_E] reciprocal norms
    int ldc, M, N, K, KS, accumulate, swapB, swapT, a_mn, b_mn;
    int bn, norm_E;                   // tile width along N (256, or the largest multiple of norm_E below it)
    int tm, tn, ksplit, sper;         // tiles, k-splits, stages per split
};

// fp32 [rows, cols] (ld) -> bf16 [rows, ldd] (ldd % 8 == 0, zero padded): one 16-byte store per thread
__global__ void convert_bf16_kernel(const float* __restrict__ src, int rows, int cols, int ld, int ldd, uint4* __restrict__ dst) {
    const int upr = ldd >> 3;
    const int64_t units = (int64_t)rows * upr;
    const bool vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < units; u += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = u / upr;
        const int c0 = (int)(u % upr) * 8;
        const float* s = src + r * ld + c0;
        float v[8];
        if (c0 + 8 <= cols && vec_ok) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(s)), b = __ldg(reinterpret_cast<const float4*>(s) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = c0 + e < cols ? __ldg(s + e) : 0.f;
        }
        dst[u] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}

// ---- CTA-pair (cta_group::2) variants: both CTAs of the pair issue the loads, completion lands on the LEADER's (even
// rank) barrier -- a shared::cluster address with the pair-rank bit cleared; only the leader issues MMAs and commits.
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar & kPeerMask)
                 : "memory");
}
__device__ __forceinline__ void mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once the issued MMAs have completed
__device__ __forceinline__ void mma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
// Relaxed: the arrival only tells the MMA warp that this warp's TMEM reads are done (ordered by the tcgen05 fence that
// precedes it); a release arrival made every epilogue warp wait for its outstanding global stores (MEMBAR + ERRBAR were 17 %
// of the head GEMM's stall samples).
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerMask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t result_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(result_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ uint32_t pair_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void pair_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// shared -> global tile store / fp32 add-reduce through a tensor map (rows and columns outside the tensor are clipped)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src, bool add) {
    if (add)
        asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0),
                     "r"(c1), "r"(src)
                     : "memory");
    else
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1),
                     "r"(src)
                     : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// E (multiple of 8, <= GT_EMAX) accumulator columns of this thread's row -> registers (no wait)
__device__ __forceinline__ void load_group(uint32_t taddr, int E, uint32_t (&dst)[GT_EMAX]) {
#pragma unroll
    for (int c = 0; c < GT_EMAX / 8; ++c)
        if (c * 8 < E) tmem_ld8(taddr + c * 8, &dst[c * 8]);
}

template <int CTAS>
__device__ __forceinline__ void gemm_tc_body(const GtParams& p) {
    using Cfg = GtCfg<CTAS>;
    constexpr int GT_STAGES = Cfg::STAGES, GT_STAGE_BYTES = Cfg::STAGE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bars[2 * GT_MAX_STAGES + 4];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = CTAS == 2 ? pair_rank() : 0u;            // 0 = leader of the pair
    const uint32_t full = smem_u32(&bars[0]), empty = smem_u32(&bars[GT_STAGES]);
    const uint32_t tfull = smem_u32(&bars[2 * GT_STAGES]), tempty = tfull + 16;
    if (tid == 0) {
        // full: the leader's arrive.expect_tx (covers the bytes of both CTAs of a pair); empty / tfull: one
        // tcgen05.commit; tempty: one arrival per epilogue warp of every CTA of the pair
        for (int s = 0; s < GT_STAGES; ++s) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull + 8 * b, 1); mbar_init(tempty + 8 * b, 8 * CTAS); }
        mbar_fence_init();
    }
    if (warp == 1) { if (CTAS == 2) tmem_alloc_pair(smem_u32(&tmem_base_s), 512); else tmem_alloc(smem_u32(&tmem_base_s), 512); }
    tc_fence_before();
    if (CTAS == 2) pair_sync(); else __syncthreads();              // barriers of BOTH CTAs initialised before any remote signal
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int items = p.tm * p.tn * p.ksplit;
    const int first = CTAS == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int stride = CTAS == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int hb = p.bn / CTAS;                                    // B rows (output columns) staged by one CTA

    if (warp == 0) {
        // ---------------- loader: TMA box loads, one lane ----------------
        if (lane == 0) {
            uint32_t g = 0;
            auto load = [&](uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
                if (CTAS == 2) tma_load_2d_pair(dst, map, c0, c1, bar); else tma_load_2d(dst, map, c0, c1, bar);
            };
            for (int it = first; it < items; it += stride) {
                const int split = it % p.ksplit, tile = it / p.ksplit;
                const int n0 = (tile % p.tn) * p.bn + (int)rank * hb;                    // this CTA's share of B
                const int m0 = (tile / p.tn) * (GT_BM * CTAS) + (int)rank * GT_BM;       // this CTA's rows of A
                const int sbeg = split * p.sper, send = min(p.KS, sbeg + p.sper);
                for (int j = sbeg; j < send; ++j, ++g) {
                    const uint32_t slot = g % GT_STAGES, ph = (g / GT_STAGES) & 1;
                    mbar_wait(empty + 8 * slot, ph ^ 1);
                    const uint32_t sa = smem_u32(smem + slot * GT_STAGE_BYTES), sb = sa + GT_A_BYTES;
                    const uint32_t bar = full + 8 * slot;
                    if (rank == 0) mbar_expect_tx(bar, GT_STAGE_BYTES * CTAS);
                    const int ka = j * GT_BK, kb = ka;
                    if (!p.a_mn) load(sa, &p.mapA, ka, m0, bar);
                    else {
#pragma unroll
                        for (int i = 0; i < GT_BM / 64; ++i) load(sa + i * 8192, &p.mapA, m0 + i * 64, ka, bar);
                    }
                    if (!p.b_mn) load(sb, &p.mapB, kb, n0, bar);
                    else {
#pragma unroll
                        for (int i = 0; i < Cfg::B_ROWS / 64; ++i) load(sb + i * 8192, &p.mapB, n0 + i * 64, kb, bar);
                    }
                }
            }
        }
    } else if (warp == 1) {
      if (rank == 0) {
        // ---------------- MMA issuer (converged loop, elected lane; the leader CTA of a pair) ----------------
        const uint32_t idesc = idesc_bf16(GT_BM * CTAS, p.bn, p.a_mn, p.b_mn);
        // K-major: 8-row groups 1024 B apart, K step of 16 = +32 B inside the 128-B swizzle row.
        // MN-major: 8-k groups 1024 B apart (SBO), 64-wide mn atoms 8192 B apart (LBO), K step of 16 = +2048 B.
        const uint32_t a_step = p.a_mn ? 2048u : 32u, a_lbo = p.a_mn ? 8192u : 16u;
        const uint32_t b_step = p.b_mn ? 2048u : 32u, b_lbo = p.b_mn ? 8192u : 16u;
        const bool leader = elect_one();
        uint32_t g = 0, ti = 0;
        for (int it = first; it < items; it += stride, ++ti) {
            const int split = it % p.ksplit;
            const int sbeg = split * p.sper, send = min(p.KS, sbeg + p.sper);
            const uint32_t buf = ti & 1, tph = (ti >> 1) & 1;
            mbar_wait(tempty + 8 * buf, tph ^ 1);
            tc_fence_after();
            const uint32_t dcol = tmem + buf * GT_BN;
            for (int j = sbeg; j < send; ++j, ++g) {
                const uint32_t slot = g % GT_STAGES, ph = (g / GT_STAGES) & 1;
                mbar_wait(full + 8 * slot, ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + slot * GT_STAGE_BYTES), sb = sa + GT_A_BYTES;
#pragma unroll
                for (int kk = 0; kk < GT_BK / 16; ++kk) {
                    const uint64_t ad = smem_desc_sw128(sa + kk * a_step, a_lbo, 1024);
                    const uint64_t bd = smem_desc_sw128(sb + kk * b_step, b_lbo, 1024);
                    if (leader) { if (CTAS == 2) mma_bf16_pair(dcol, ad, bd, idesc, !(j == sbeg && kk == 0)); else mma_bf16(dcol, ad, bd, idesc, !(j == sbeg && kk == 0)); }
                }
                if (leader) { if (CTAS == 2) mma_commit_pair(empty + 8 * slot); else mma_commit(empty + 8 * slot); }
            }
            if (leader) { if (CTAS == 2) mma_commit_pair(tfull + 8 * buf); else mma_commit(tfull + 8 * buf); }
        }
      }
    } else {
        // ---------------- epilogue: warps 2..9; TMEM lane quadrant = warp % 4, two warps per quadrant ----------------
        // A thread owns one accumulator row (TMEM lane) and stores 128..192 contiguous bytes of it per chunk.  The
        // epilogue is latency bound (tcgen05.ld -> bias -> stores), not bandwidth bound: two warps per quadrant take
        // alternate column chunks and each prefetches its next chunk from TMEM while it stores the current one.
        const int q = warp & 3, half = (warp - 2) >> 2;
        const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                            ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
        const bool atomic = p.ksplit > 1;
        const int E = p.norm_E;
        // Output through TMA: the warp re-tiles its 32 rows x (32 | E) columns in a private shared-memory block and one
        // lane issues a tile store (or fp32 add-reduce for accumulate / split-K).  Stores straight from the
        // row-per-thread registers touch 32 half-used sectors per instruction and saturate the L1TEX / L2 store path.
        const bool ctma = p.c_tma != 0;
        uint8_t* stg = smem + GT_STAGES * GT_STAGE_BYTES + (warp - 2) * GT_STG_BYTES;
        const uint32_t stg_u32 = smem_u32(stg);
        const bool c_add = atomic || p.accumulate;
        uint32_t ti = 0;
        auto release = [&](uint32_t bar) {                  // one arrival per warp on the (leader's) tempty barrier
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CTAS == 2) mbar_arrive_leader(bar); else mbar_arrive(bar); }
        };
        for (int it = first; it < items; it += stride, ++ti) {
            const int split = it % p.ksplit, tile = it / p.ksplit;
            const int n0 = (tile % p.tn) * p.bn, m0 = (tile / p.tn) * (GT_BM * CTAS) + (int)rank * GT_BM;
            const uint32_t buf = ti & 1, tph = (ti >> 1) & 1;
            mbar_wait(tfull + 8 * buf, tph);
            tc_fence_after();
            const int m = m0 + q * 32 + lane;
            size_t row = (size_t)m;
            if (p.swapB > 0 && m < p.M) row = (size_t)(m % p.swapB) * p.swapT + (size_t)(m / p.swapB);
            float* crow = p.C + row * p.ldc;
            const bool add_bias = p.bias != nullptr && split == 0;
            const uint32_t tacc = tmem + ((uint32_t)(q * 32) << 16) + buf * GT_BN;
            if (E > 0) {
                // fused tf.nn.l2_normalize over groups of E output columns (utils/ops.py:323-324): C = normalised
                // rows, inv[row][group] = 1/norm (negative on the clamped branch, as amss_l2norm_fwd writes it)
                const int groups = min(p.bn, p.N - n0) / E, e4 = E >> 2, ngr = p.N / E;
#pragma unroll 1
                for (int g = half; g < groups; g += 2) {
                    uint32_t v[GT_EMAX];
                    const int nb = n0 + g * E;
                    // the group's bias first: all loads in flight together and under the TMEM load (a load next to each use
                    // serialised ten L1 / L2 round trips per group -- 35 % of the kernel's stall samples, ncu source page)
                    float4 bb[GT_EMAX / 4];
#pragma unroll
                    for (int c = 0; c < GT_EMAX / 4; ++c)
                        bb[c] = (add_bias && c < e4) ? __ldg(reinterpret_cast<const float4*>(p.bias + nb) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    load_group(tacc + g * E, E, v);
                    tmem_ld_wait();
                    float ss = 0.f;
#pragma unroll
                    for (int c = 0; c < GT_EMAX / 4; ++c)
                        if (c < e4) {
                            if (add_bias) {
                                v[4 * c] = __float_as_uint(__uint_as_float(v[4 * c]) + bb[c].x);
                                v[4 * c + 1] = __float_as_uint(__uint_as_float(v[4 * c + 1]) + bb[c].y);
                                v[4 * c + 2] = __float_as_uint(__uint_as_float(v[4 * c + 2]) + bb[c].z);
                                v[4 * c + 3] = __float_as_uint(__uint_as_float(v[4 * c + 3]) + bb[c].w);
                            }
#pragma unroll
                            for (int e = 0; e < 4; ++e) ss = fmaf(__uint_as_float(v[4 * c + e]), __uint_as_float(v[4 * c + e]), ss);
                        }
                    const float inv = rsqrtf(fmaxf(ss, 1e-12f));
                    if (ctma) {
                        if (lane == 0) tma_store_wait_read();       // the previous store has finished reading the block
                        __syncwarp();
                        float4* dst = reinterpret_cast<float4*>(stg + lane * E * 4);
#pragma unroll
                        for (int c = 0; c < GT_EMAX / 4; ++c)
                            if (c < e4)
                                dst[c] = make_float4(__uint_as_float(v[4 * c]) * inv, __uint_as_float(v[4 * c + 1]) * inv,
                                                     __uint_as_float(v[4 * c + 2]) * inv, __uint_as_float(v[4 * c + 3]) * inv);
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) tma_store_2d(&p.mapC, nb, m0 + q * 32, stg_u32, false);
                        if (m < p.M) p.inv[row * ngr + nb / E] = (ss >= 1e-12f) ? inv : -inv;
                    } else if (m < p.M) {
                        float4* dst = reinterpret_cast<float4*>(crow + nb);
#pragma unroll
                        for (int c = 0; c < GT_EMAX / 4; ++c)
                            if (c < e4)
                                dst[c] = make_float4(__uint_as_float(v[4 * c]) * inv, __uint_as_float(v[4 * c + 1]) * inv,
                                                     __uint_as_float(v[4 * c + 2]) * inv, __uint_as_float(v[4 * c + 3]) * inv);
                        p.inv[row * ngr + nb / E] = (ss >= 1e-12f) ? inv : -inv;
                    }
                }
                release(tempty + 8 * buf);                  // this warp's share of the accumulator is drained
                continue;
            }
            const int chunks = (min(p.bn, p.N - n0) + 31) >> 5;
            uint32_t v[32], vn[32];
            if (half < chunks) { tmem_ld32(tacc + half * 32, v); tmem_ld_wait(); }
#pragma unroll 1
            for (int ci = half; ci < chunks; ci += 2) {
                const bool more = ci + 2 < chunks;
                if (more) tmem_ld32(tacc + (ci + 2) * 32, vn);
                const int nb = n0 + ci * 32;
                float4 bb[8];                                       // the chunk's bias, all loads in flight together
#pragma unroll
                for (int gq = 0; gq < 8; ++gq)
                    bb[gq] = (add_bias && vec_ok && nb + 4 * gq + 4 <= p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + nb) + gq)
                                                                           : make_float4(0.f, 0.f, 0.f, 0.f);
                if (ctma && ci * 32 + 32 <= p.bn) {                 // (a chunk cut by the tile edge takes the direct path)
                    if (lane == 0) tma_store_wait_read();           // the previous store has finished reading the block
                    __syncwarp();
#pragma unroll
                    for (int gq = 0; gq < 8; ++gq) {
                        float4 o = make_float4(__uint_as_float(v[4 * gq]), __uint_as_float(v[4 * gq + 1]),
                                               __uint_as_float(v[4 * gq + 2]), __uint_as_float(v[4 * gq + 3]));
                        if (add_bias && nb + 4 * gq + 4 <= p.N) { o.x += bb[gq].x; o.y += bb[gq].y; o.z += bb[gq].z; o.w += bb[gq].w; }
                        // 128-byte rows, SWIZZLE_128B: 16-byte chunk gq of row r sits at chunk gq ^ (r & 7)
                        *reinterpret_cast<float4*>(stg + lane * 128 + ((gq ^ (lane & 7)) << 4)) = o;
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) tma_store_2d(&p.mapC, nb, m0 + q * 32, stg_u32, c_add);
                } else if (m < p.M) {
                    const int nend = min(p.N, n0 + p.bn);
                    if (!atomic && vec_ok && nb + 32 <= nend) {
                        float4* dst = reinterpret_cast<float4*>(crow + nb);
#pragma unroll
                        for (int gq = 0; gq < 8; ++gq) {
                            float4 o = make_float4(__uint_as_float(v[4 * gq]), __uint_as_float(v[4 * gq + 1]),
                                                   __uint_as_float(v[4 * gq + 2]), __uint_as_float(v[4 * gq + 3]));
                            if (add_bias) { o.x += bb[gq].x; o.y += bb[gq].y; o.z += bb[gq].z; o.w += bb[gq].w; }
                            if (p.accumulate) { const float4 old = dst[gq]; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                            dst[gq] = o;
                        }
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) {
                            const int n = nb + jj;
                            if (n < nend) {
                                float o = __uint_as_float(v[jj]);
                                if (add_bias) o += __ldg(p.bias + n);
                                if (atomic) atomicAdd(crow + n, o);
                                else crow[n] = p.accumulate ? crow[n] + o : o;
                            }
                        }
                    }
                }
                if (more) {
                    tmem_ld_wait();
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) v[jj] = vn[jj];
                }
            }
            release(tempty + 8 * buf);                      // this warp's share of the accumulator is drained
        }
        if (ctma && lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    if (CTAS == 2) pair_sync(); else __syncthreads();              // both CTAs done with each other's barriers / TMEM
    if (warp == 1) { if (CTAS == 2) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512); }
}

__global__ void __launch_bounds__(GT_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GtParams p) { gemm_tc_body<1>(p); }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GT_THREADS, 1) gemm_tc_pair_kernel(const __grid_constant__ GtParams p) {
    gemm_tc_body<2>(p);
}

__global__ void zero_rows_kernel(float* C, int M, int N, int ldc) {
    const int64_t n = (int64_t)M * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        C[(i / N) * ldc + (i % N)] = 0.f;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) { cudaGetLastError(); f = nullptr; }
        return (EncodeTiledFn)f;
    }();
    return fn;
}

// Operand X[r][k], r < R, k < K.  K-major: src[r*ld + k], box {64 k, rows_box}; MN-major: src[k*ld + r], box {64 r, 64 k}.
int make_operand_map(CUtensorMap* m, const uint16_t* src, int R, int K, int ld, int mn, int rows_box) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) { set_error("gemm_tc: cuTensorMapEncodeTiled is not available from this driver"); return AMSS_ERR_CUDA; }
    cuuint64_t dims[2], strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2], es[2] = {1, 1};
    if (!mn) { dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)R; box[0] = GT_BK; box[1] = (cuuint32_t)rows_box; }
    else { dims[0] = (cuuint64_t)R; dims[1] = (cuuint64_t)K; box[0] = 64; box[1] = GT_BK; }
    const CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t*>(src), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d) for R=%d K=%d ld=%d mn=%d", (int)rc, R, K, ld, mn);
        return AMSS_ERR_CUDA;
    }
    return AMSS_OK;
}

// fp32 output C[M][N] (ldc): box {bw columns, 32 rows}; 128-byte rows are swizzled (conflict-free staging writes)
int make_c_map(CUtensorMap* m, float* C, int M, int N, int ldc, int bw) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) { set_error("gemm_tc: cuTensorMapEncodeTiled is not available from this driver"); return AMSS_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M}, strides[1] = {(cuuint64_t)ldc * 4};
    cuuint32_t box[2] = {(cuuint32_t)bw, 32}, es[2] = {1, 1};
    const CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, C, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            bw == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d) for C M=%d N=%d ldc=%d", (int)rc, M, N, ldc); return AMSS_ERR_CUDA; }
    return AMSS_OK;
}

inline int pad8(int x) { return (x + 7) & ~7; }

}  // namespace

bool gemm_tc_supported(int M, int N, int K, int lda, int ldb, int ldc, int transa, int transb) {
    (void)lda; (void)ldb; (void)ldc; (void)transa; (void)transb;
    return M >= 1 && N >= 1 && K >= 1;
}

// bf16 copies of the two fp32 operands (leading dimensions padded to 8 elements = 16 bytes, a TMA requirement)
size_t gemm_tc_workspace(int M, int N, int K, int transa, int transb, int precision) {
    (void)precision;
    const size_t a = transa ? (size_t)K * pad8(M) : (size_t)M * pad8(K);
    const size_t b = transb ? (size_t)N * pad8(K) : (size_t)K * pad8(N);
    return 256 + align_up(a * 2, 256) + align_up(b * 2, 256);
}

int convert_bf16(const float* src, int rows, int cols, int ld, uint16_t* dst, int ldd, cudaStream_t st) {
    const int64_t units = (int64_t)rows * (ldd / 8);
    AMSS_LAUNCH(convert_bf16_kernel, (int)std::max<int64_t>(1, std::min<int64_t>((units + 255) / 256, 16 * kNumSMs)), 256, 0, st, src,
                rows, cols, ld, ldd, (uint4*)dst);
    return AMSS_OK;
}

// C[M,N] (+)= A B^T-style product of two bf16 row-major operands.  a_mn = 0: A[m][k] = A[m*lda + k]; 1: A[k*lda + m].
// b_mn = 0: B[n][k] = B[n*ldb + k]; 1: B[k*ldb + n].  lda, ldb multiples of 8 and base pointers 16-byte aligned (TMA;
// a box may not start at an inner coordinate that is not a multiple of 16 bytes either -- the unit raises an
// illegal-instruction fault -- so sub-matrix views are passed as offset POINTERS that keep this alignment).
// norm_E > 0 (multiple of 8, <= 48, divides N; no accumulate): the epilogue l2-normalises every group of norm_E
// consecutive output columns and writes the reciprocal norms to inv[M][N / norm_E] (amss_l2norm_fwd semantics).
int gemm_bf16(const uint16_t* A, int lda, int a_mn, const uint16_t* B, int ldb, int b_mn, const float* bias, int M, int N,
              int K, int accumulate, float* C, int ldc, int swapB, int swapT, int norm_E, float* inv, cudaStream_t st) {
    if ((lda & 7) || (ldb & 7) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) {
        set_error("gemm_bf16: operands must be 16-byte aligned with leading dimensions that are multiples of 8 (lda=%d ldb=%d)", lda, ldb);
        return AMSS_ERR_INVALID_ARG;
    }
    GtParams p;
    p.bn = GT_BN; p.norm_E = 0; p.inv = nullptr;
    if (norm_E > 0) {
        if ((norm_E & 7) || norm_E > GT_EMAX || N % norm_E || accumulate || !inv || (ldc & 3) ||
            (reinterpret_cast<uintptr_t>(C) & 15)) {
            set_error("gemm_bf16: fused l2-normalise needs norm_E %% 8 == 0, norm_E <= %d, N %% norm_E == 0, no accumulate (E=%d N=%d)",
                      GT_EMAX, norm_E, N);
            return AMSS_ERR_UNSUPPORTED;
        }
        p.bn = GT_BN / norm_E * norm_E; p.norm_E = norm_E; p.inv = inv;
    }
    // Tile shape.  Columns: the fewest tiles of at most 256 columns, each a multiple of 32 wide (whole 32-column TMA store
    // boxes; N = 600 -> 3 x 224 instead of 3 x 256: the MMA work follows the tile, not N).  Rows: CTA pairs (256-row
    // tiles, a third less operand traffic per flop) whenever there is more than one 128-row tile; measured faster than
    // single CTAs even at M = 600 (768 padded rows).  AMSS_GEMM_CTAS=1|2 forces the choice (debug).
    if (!norm_E) {
        const int nt = (N + GT_BN - 1) / GT_BN;
        p.bn = std::min(GT_BN, ((N + nt - 1) / nt + 31) & ~31);
    }
    static const int forced = [] { const char* e = getenv("AMSS_GEMM_CTAS"); return e ? atoi(e) : 0; }();
    const int ctas = forced == 1 ? 1 : (forced == 2 ? 2 : (M > GT_BM ? 2 : 1));
    p.tm = (M + GT_BM * ctas - 1) / (GT_BM * ctas); p.tn = (N + p.bn - 1) / p.bn;
    p.KS = (K + GT_BK - 1) / GT_BK;
    int rc = make_operand_map(&p.mapA, A, M, K, lda, a_mn, GT_BM);
    if (rc != AMSS_OK) return rc;
    rc = make_operand_map(&p.mapB, B, N, K, ldb, b_mn, GT_BN / ctas);
    if (rc != AMSS_OK) return rc;
    p.c_tma = swapB == 0 && (ldc & 3) == 0 && (N & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(bias) & 15) == 0;
    if (p.c_tma) {
        rc = make_c_map(&p.mapC, C, M, N, ldc, norm_E > 0 ? norm_E : 32);
        if (rc != AMSS_OK) return rc;
    }
    p.a_mn = a_mn; p.b_mn = b_mn;
    p.bias = bias; p.C = C; p.ldc = ldc;
    p.M = M; p.N = N; p.K = K; p.accumulate = accumulate; p.swapB = swapB; p.swapT = swapT;
    const int tiles = p.tm * p.tn, slots = kNumSMs / ctas;      // concurrently running tiles
    // split-K against wave quantisation: items = tiles * ksplit run `slots` at a time; take the split with the best
    // occupancy of the last wave (each extra split costs one more add-reduce pass over C, hence the small penalty)
    int ksplit = 1;
    if (p.KS >= 8 && !norm_E && tiles < 4 * slots) {
        double best = 0.0;
        for (int ks = 1; ks <= std::min(p.KS / 4, 32); ++ks) {
            const int items = tiles * ks, waves = (items + slots - 1) / slots;
            const double eff = (double)items / ((double)waves * slots) - 0.004 * (ks - 1);
            if (eff > best + 1e-9) { best = eff; ksplit = ks; }
        }
    }
    const int sper = (p.KS + ksplit - 1) / ksplit;
    ksplit = (p.KS + sper - 1) / sper;
    p.ksplit = ksplit;
    p.sper = sper;
    if (ksplit > 1 && !accumulate) {
        const int64_t n = (int64_t)M * N;
        AMSS_LAUNCH(zero_rows_kernel, (int)std::min<int64_t>((n + 255) / 256, 8 * kNumSMs), 256, 0, st, C, M, N, ldc);
    }
    const int grid = std::min(tiles * ksplit, slots) * ctas;
    if (ctas == 2) {
        AMSS_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GtCfg<2>::SMEM));
        AMSS_LAUNCH(gemm_tc_pair_kernel, grid, GT_THREADS, GtCfg<2>::SMEM, st, p);
    } else {
        AMSS_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GtCfg<1>::SMEM));
        AMSS_LAUNCH(gemm_tc_kernel, grid, GT_THREADS, GtCfg<1>::SMEM, st, p);
    }
    return AMSS_OK;
}

int gemm_tc(const float* A, int lda, const float* B, int ldb, const float* bias, int M, int N, int K, int transa,
            int transb, int accumulate, int precision, float* C, int ldc, int swapB, int swapT, void* workspace,
            size_t workspace_bytes, cudaStream_t st) {
    if (!workspace || workspace_bytes < gemm_tc_workspace(M, N, K, transa, transb, precision)) {
        set_error("gemm_tc: workspace too small (%zu < %zu)", workspace_bytes, gemm_tc_workspace(M, N, K, transa, transb, precision));
        return AMSS_ERR_WORKSPACE;
    }
    // A: not transposed -> [M,K] K-major; transposed -> source [K,M], MN-major.  B: not transposed -> source [K,N],
    // MN-major; transposed -> source [N,K], K-major.
    const int ar = transa ? K : M, ac = transa ? M : K, br = transb ? N : K, bc = transb ? K : N;
    uint16_t* Ab = (uint16_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    uint16_t* Bb = Ab + align_up((size_t)ar * pad8(ac) * 2, 256) / 2;
    int rc = convert_bf16(A, ar, ac, lda, Ab, pad8(ac), st);
    if (rc != AMSS_OK) return rc;
    rc = convert_bf16(B, br, bc, ldb, Bb, pad8(bc), st);
    if (rc != AMSS_OK) return rc;
    return gemm_bf16(Ab, pad8(ac), transa ? 1 : 0, Bb, pad8(bc), transb ? 0 : 1, bias, M, N, K, accumulate, C, ldc, swapB, swapT, 0,
                     nullptr, st);
}

}  // namespace amss
