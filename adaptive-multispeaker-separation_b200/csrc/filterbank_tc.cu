// Adaptive analysis filterbank on the 5th-gen tensor cores (models/adapt.py:115-117):
//   y[r,tp,n], argmax[r,tp,n] = max / arg-max over t in [tp*pool, (tp+1)*pool) of
//   X[r,t,n] = sum_k x[r, t + k - pad_left] * filt[k,n]          (conv2d SAME stride 1, then
//   max_pool_with_argmax VALID with pool == hop)
// as a Toeplitz implicit GEMM, D[n, t] = sum_k Wt[n,k] * H[t,k], H[t,k] = xp[t+k] (Hankel):
//   * A operand (M = 128 filters per CTA) = bf16 filters, pre-packed per 64-tap stage in the
//     canonical no-swizzle K-major core-matrix layout, streamed by the TMA engine
//     (cp.async.bulk, 16 KB per stage) through a 6-deep mbarrier ring;
//   * B operand (N = 256 time positions per tile) = the Hankel matrix, NEVER materialised: shared
//     memory holds G[u] = bf16(xp[t0+u .. t0+u+8)) (16 B per sample, 20 KB per tile) and because
//     H is constant along anti-diagonals the core matrix (t/8, k/8) is the 128-byte block
//     G[8*(t/8 + k/8) ..], i.e. one descriptor with LBO = SBO = 128 B serves every tap offset;
//   * accumulators: 128 lanes (filters) x 256 columns (time) fp32 in TMEM, double buffered, so the
//     max / arg-max over time is a per-thread scan over columns (no shuffles) that overlaps the
//     next tile's MMAs.  The [Bt, L, N] tensor never exists anywhere.
// Warp roles (384 threads): w0 filter-stage loader, w1 MMA issuer, w2 TMEM allocator, w4-7 epilogue
// (TMEM lane quadrant = warp % 4), w8-11 Hankel builders.  Persistent: one CTA per SM.
#include "common.cuh"
#include "tc.cuh"
#include <cstdlib>

namespace amss {
namespace {

using namespace tc;

constexpr int FT_THREADS = 384;
constexpr int FT_NT = 256;            // time positions per tile (MMA N)
constexpr int FT_KS = 64;             // taps per filter stage
constexpr int FT_STAGE_BYTES = 128 * FT_KS * 2;   // 16 KB
constexpr int FT_STAGES = 6;

struct FtParams {
    const float* x;                   // [Bt][L]
    const uint8_t* packed;            // [MT][KS][16 KB]
    float* y;                         // [Bt][Tp][N]
    int64_t* argmax;                  // [Bt][Tp][N] or null
    int Bt, L, N, Tp, pool, pl, KS, MT, QB, tiles_per_unit, units_per_signal;
    const int* gate;                  // device flag written by mix_is_sum_kernel (null = always run)
    int gate_zero;                    // run iff (*gate == 0) == (gate_zero != 0)
    int B;                            // pair kernel: mixtures (rows [0,B) = mixtures, [B,3B) = their two sources)
};

__device__ __forceinline__ bool gated_off(const FtParams& p) {
    if (!p.gate) return false;
    const int g = *reinterpret_cast<const volatile int*>(p.gate);
    return (g == 0) != (p.gate_zero != 0);
}

// filt[W][N] fp32 -> bf16 stages: element (filter f, tap k) of m-tile m, stage j at
//   ((m*KS + j) * 16 KB) + ((k%64)/8 * 16 + (f%128)/8) * 128 + (f%8)*16 + (k%8)*2
__global__ void pack_filter_kernel(const float* __restrict__ filt, int W, int N, int KS, int MT,
                                   uint4* __restrict__ packed) {
    const int64_t units = (int64_t)MT * KS * (FT_STAGE_BYTES / 16);
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < units; u += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(u & 7);
        const int fgrp = (int)((u >> 3) & 15);
        const int kch = (int)((u >> 7) & 7);
        const int64_t st = u >> 10;
        const int j = (int)(st % KS), m = (int)(st / KS);
        const int f = m * 128 + fgrp * 8 + r;
        const int k0 = j * FT_KS + kch * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (f < N && k0 + e < W) ? filt[(size_t)(k0 + e) * N + f] : 0.f;
        packed[u] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(FT_THREADS, 1) analysis_tc_kernel(FtParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // carve-up: [A stages][G0][G1][xs]
    uint8_t* a_stage = smem;
    const uint32_t g_bytes = (uint32_t)p.QB * 128;
    uint8_t* g_buf = a_stage + FT_STAGES * FT_STAGE_BYTES;
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(g_buf + 2 * g_bytes);
    __shared__ __align__(8) uint64_t bars[2 * FT_STAGES + 8];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mtile = blockIdx.x % p.MT, rank = blockIdx.x / p.MT, nranks = gridDim.x / p.MT;
    if (rank >= nranks) return;   // surplus CTAs when gridDim is not a multiple of MT
    if (gated_off(p)) return;     // the linear-mixture kernel handles this batch

    const uint32_t a_full = smem_u32(&bars[0]), a_empty = smem_u32(&bars[FT_STAGES]);
    const uint32_t g_full = smem_u32(&bars[2 * FT_STAGES]), g_empty = g_full + 16;
    const uint32_t t_full = g_full + 32, t_empty = g_full + 48;
    if (tid == 0) {
        for (int s = 0; s < FT_STAGES; ++s) { mbar_init(a_full + 8 * s, 1); mbar_init(a_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(g_full + 8 * b, 128); mbar_init(g_empty + 8 * b, 1);
            mbar_init(t_full + 8 * b, 1);   mbar_init(t_empty + 8 * b, 128);
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(&tmem_base_s), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    const int64_t units = (int64_t)p.Bt * p.units_per_signal;
    const int tpu = p.tiles_per_unit;

    if (warp == 0) {
        // ---------------- filter-stage loader (one lane) ----------------
        if (lane == 0) {
            const uint8_t* src = p.packed + (size_t)mtile * p.KS * FT_STAGE_BYTES;
            uint32_t ga = 0;
            for (int64_t u = rank; u < units; u += nranks)
                for (int tt = 0; tt < tpu; ++tt)
                    for (int j = 0; j < p.KS; ++j, ++ga) {
                        const uint32_t slot = ga % FT_STAGES, ph = (ga / FT_STAGES) & 1;
                        mbar_wait(a_empty + 8 * slot, ph ^ 1);
                        mbar_expect_tx(a_full + 8 * slot, FT_STAGE_BYTES);
                        bulk_g2s(smem_u32(a_stage + slot * FT_STAGE_BYTES), src + (size_t)j * FT_STAGE_BYTES,
                                 FT_STAGE_BYTES, a_full + 8 * slot);
                    }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (one lane) ----------------
        if (lane == 0) {
            const uint32_t idesc = idesc_bf16(128, FT_NT, 0, 0);
            uint32_t ga = 0, it = 0;
            for (int64_t u = rank; u < units; u += nranks)
                for (int tt = 0; tt < tpu; ++tt, ++it) {
                    const uint32_t buf = it & 1, ph = (it >> 1) & 1;
                    mbar_wait(t_empty + 8 * buf, ph ^ 1);
                    mbar_wait(g_full + 8 * buf, ph);
                    tc_fence_after();
                    const uint32_t gaddr = smem_u32(g_buf + buf * g_bytes);
                    const uint32_t dcol = tmem + buf * FT_NT;
                    for (int j = 0; j < p.KS; ++j, ++ga) {
                        const uint32_t slot = ga % FT_STAGES, aph = (ga / FT_STAGES) & 1;
                        mbar_wait(a_full + 8 * slot, aph);
                        tc_fence_after();
                        const uint32_t aaddr = smem_u32(a_stage + slot * FT_STAGE_BYTES);
#pragma unroll
                        for (int kk = 0; kk < FT_KS / 16; ++kk) {
                            const uint64_t ad = smem_desc(aaddr + kk * 4096, 2048, 128);
                            const uint64_t bd = smem_desc(gaddr + (j * 8 + kk * 2) * 128, 128, 128);
                            mma_bf16(dcol, ad, bd, idesc, (j | kk) != 0);
                        }
                        mma_commit(a_empty + 8 * slot);
                    }
                    mma_commit(t_full + 8 * buf);
                    mma_commit(g_empty + 8 * buf);
                }
        }
    } else if (warp >= 4 && warp < 8) {
        // ---------------- epilogue: max / arg-max over the time columns ----------------
        const int q = warp & 3;
        const int f = mtile * 128 + q * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t it = 0;
        for (int64_t u = rank; u < units; u += nranks) {
            const int r = (int)(u / p.units_per_signal), ug = (int)(u % p.units_per_signal);
            float best = -INFINITY;
            int bestt = 0;
            for (int tt = 0; tt < tpu; ++tt, ++it) {
                const uint32_t buf = it & 1, ph = (it >> 1) & 1;
                const int t0 = (ug * tpu + tt) * FT_NT;
                mbar_wait(t_full + 8 * buf, ph);
                tc_fence_after();
#pragma unroll 1
                for (int c0 = 0; c0 < FT_NT; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(lane_base + buf * FT_NT + c0, v);
                    tmem_ld_wait();
                    if (c0 + 32 == FT_NT) {   // accumulator drained: hand the buffer back before the stores
                        tc_fence_before();
                        mbar_arrive(t_empty + 8 * buf);
                    }
                    const int tb = t0 + c0;
                    if (p.pool < 32) {
                        // not reachable (supported() requires pool >= 32)
                    }
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) {
                        const float val = __uint_as_float(v[jj]);
                        if (val > best) { best = val; bestt = tb + jj; }
                    }
                    if (((tb + 32) % p.pool) == 0) {
                        const int tp = (tb + 32) / p.pool - 1;
                        if (tp < p.Tp && f < p.N) {
                            const size_t o = ((size_t)r * p.Tp + tp) * p.N + f;
                            p.y[o] = best;
                            if (p.argmax) p.argmax[o] = (int64_t)bestt * p.N + f;
                        }
                        best = -INFINITY;
                        bestt = tb + 32;
                    }
                }
            }
        }
    } else if (warp >= 8) {
        // ---------------- Hankel builders: G[u] = bf16(xp[t0+u .. t0+u+8)) ----------------
        const int bt = tid - 256;   // 0..127
        const int nunits = p.QB * 8;
        const int ns = nunits + 8;
        uint32_t it = 0;
        for (int64_t u = rank; u < units; u += nranks) {
            const int r = (int)(u / p.units_per_signal), ug = (int)(u % p.units_per_signal);
            const float* xr = p.x + (size_t)r * p.L;
            for (int tt = 0; tt < tpu; ++tt, ++it) {
                const uint32_t buf = it & 1, ph = (it >> 1) & 1;
                const int t0 = (ug * tpu + tt) * FT_NT;
                for (int i = bt; i < ns; i += 128) {
                    const int s = t0 + i - p.pl;
                    xs[i] = __float2bfloat16_rn((s >= 0 && s < p.L) ? __ldg(xr + s) : 0.f);
                }
                named_bar_sync(1, 128);
                mbar_wait(g_empty + 8 * buf, ph ^ 1);
                uint8_t* g = g_buf + buf * g_bytes;
                const unsigned short* xsu = reinterpret_cast<const unsigned short*>(xs);
                for (int i = bt; i < nunits; i += 128) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        w[e] = (uint32_t)xsu[i + 2 * e] | ((uint32_t)xsu[i + 2 * e + 1] << 16);
                    *reinterpret_cast<uint4*>(g + (size_t)i * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                fence_async_smem();
                mbar_arrive(g_full + 8 * buf);
                named_bar_sync(1, 128);   // xs is rewritten by the next tile
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 512);
}

// =================================================================================================
// Mixture rows by linearity.  The reference runs the front end on [B mixtures ; B*S sources] (adapt.py:41-48) and its
// data pipeline builds every mixture as the fp32 sum of its sources (data/dataset.py:462-468).  The convolution is
// linear, so when x_mix == x_0 + x_1 holds bit for bit (checked on the device for every batch, mix_is_sum_kernel) the
// mixture's pre-pool response is the sum of the two source responses: only the SOURCE rows are multiplied and the
// epilogue pools a, b and a + b.  One tile = 128 time positions of BOTH sources: ONE N = 256 MMA per K step over the two
// sources' Hankel blocks interleaved in shared memory (accumulator column 16 jb + 8 s + r = source s, time 8 jb + r), so the
// pipeline, the filter ring and the TMEM double buffering are the ones of analysis_tc_kernel; a third of the tensor work of
// the batch disappears.  (Two N = 128 MMAs per K step, one per source, read the filter operand twice: 16 KB per 128 clk = the
// whole shared-memory pipe; 4.95 -> 4.52 ms for 128 mixtures with the single MMA.)  If the check fails the kernel exits
// at once and analysis_tc_kernel (gated the other way) runs the stock three-signal path.
// =================================================================================================
constexpr int FP_NT = 128;            // time positions per tile and per source

__global__ void mix_is_sum_kernel(const float* __restrict__ x, int B, int64_t L, int* __restrict__ mismatch) {
    const int64_t n = (int64_t)B * L;
    bool bad = false;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / L, o = i - b * L;
        const float s = __fadd_rn(x[(B + 2 * b) * L + o], x[(B + 2 * b + 1) * L + o]);
        bad |= !(x[i] == s);
    }
    if (bad) *mismatch = 1;
}

// the same in 16-byte units (L % 4 == 0, aligned rows)
__global__ void mix_is_sum_vec_kernel(const float4* __restrict__ x, int B, int64_t L4, int* __restrict__ mismatch) {
    const int64_t n = (int64_t)B * L4;
    bool bad = false;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / L4, o = i - b * L4;
        const float4 m = __ldg(x + i), a = __ldg(x + (B + 2 * b) * L4 + o), c = __ldg(x + (B + 2 * b + 1) * L4 + o);
        bad |= !(m.x == __fadd_rn(a.x, c.x)) | !(m.y == __fadd_rn(a.y, c.y)) | !(m.z == __fadd_rn(a.z, c.z)) |
               !(m.w == __fadd_rn(a.w, c.w));
    }
    if (bad) *mismatch = 1;
}

__global__ void __launch_bounds__(FT_THREADS, 1) analysis_pair_tc_kernel(FtParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // carve-up: [A stages][G(buf 0: src 0, src 1)][G(buf 1: src 0, src 1)][xs src 0][xs src 1]
    uint8_t* a_stage = smem;
    const uint32_t g1 = (uint32_t)p.QB * 128, g_bytes = 2 * g1;          // QB = blocks per source here
    uint8_t* g_buf = a_stage + FT_STAGES * FT_STAGE_BYTES;
    const int ns = p.QB * 8 + 8;
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(g_buf + 2 * g_bytes);   // [2][ns]
    __shared__ __align__(8) uint64_t bars[2 * FT_STAGES + 8];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mtile = blockIdx.x % p.MT, rank = blockIdx.x / p.MT, nranks = gridDim.x / p.MT;
    if (rank >= nranks) return;
    if (gated_off(p)) return;

    const uint32_t a_full = smem_u32(&bars[0]), a_empty = smem_u32(&bars[FT_STAGES]);
    const uint32_t g_full = smem_u32(&bars[2 * FT_STAGES]), g_empty = g_full + 16;
    const uint32_t t_full = g_full + 32, t_empty = g_full + 48;
    if (tid == 0) {
        for (int s = 0; s < FT_STAGES; ++s) { mbar_init(a_full + 8 * s, 1); mbar_init(a_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(g_full + 8 * b, 128); mbar_init(g_empty + 8 * b, 1);
            mbar_init(t_full + 8 * b, 1);   mbar_init(t_empty + 8 * b, 128);
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(&tmem_base_s), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    const int64_t units = (int64_t)p.B * p.Tp;          // (mixture, pooling window)
    const int tpu = p.tiles_per_unit;                   // pool / 128

    if (warp == 0) {
        if (lane == 0) {
            const uint8_t* src = p.packed + (size_t)mtile * p.KS * FT_STAGE_BYTES;
            uint32_t ga = 0;
            for (int64_t u = rank; u < units; u += nranks)
                for (int tt = 0; tt < tpu; ++tt)
                    for (int j = 0; j < p.KS; ++j, ++ga) {
                        const uint32_t slot = ga % FT_STAGES, ph = (ga / FT_STAGES) & 1;
                        mbar_wait(a_empty + 8 * slot, ph ^ 1);
                        mbar_expect_tx(a_full + 8 * slot, FT_STAGE_BYTES);
                        bulk_g2s(smem_u32(a_stage + slot * FT_STAGE_BYTES), src + (size_t)j * FT_STAGE_BYTES,
                                 FT_STAGE_BYTES, a_full + 8 * slot);
                    }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: converged loop, elected lane ----------------
        // ONE N = 256 MMA per K step covers both sources: their Hankel blocks are INTERLEAVED in G (source s, block m at
        // 256 m + 128 s), so the operand's N index n/8 = 2 j + s walks 128-byte units (SBO = 128) and a k-block is 256 bytes
        // further (LBO = 256).  Two N = 128 MMAs read the 4 KB filter operand twice per K step: 16 KB per 128 clk is the whole
        // shared-memory pipe (ncu: tensor-core wavefronts 79.5 % + LSU 13 %, profiles/r02q_ncu_full_analysis_pair_B128.csv);
        // this form reads 12 KB.
        const uint32_t idesc = idesc_bf16(128, FT_NT, 0, 0);
        const bool leader = elect_one();
        uint32_t ga = 0, it = 0;
        for (int64_t u = rank; u < units; u += nranks)
            for (int tt = 0; tt < tpu; ++tt, ++it) {
                const uint32_t buf = it & 1, ph = (it >> 1) & 1;
                mbar_wait(t_empty + 8 * buf, ph ^ 1);
                mbar_wait(g_full + 8 * buf, ph);
                tc_fence_after();
                const uint32_t gaddr = smem_u32(g_buf + buf * g_bytes);
                const uint32_t dcol = tmem + buf * FT_NT;
                for (int j = 0; j < p.KS; ++j, ++ga) {
                    const uint32_t slot = ga % FT_STAGES, aph = (ga / FT_STAGES) & 1;
                    mbar_wait(a_full + 8 * slot, aph);
                    tc_fence_after();
                    const uint32_t aaddr = smem_u32(a_stage + slot * FT_STAGE_BYTES);
#pragma unroll
                    for (int kk = 0; kk < FT_KS / 16; ++kk) {
                        const uint64_t ad = smem_desc(aaddr + kk * 4096, 2048, 128);
                        const uint64_t bd = smem_desc(gaddr + (j * 8 + kk * 2) * 256, 256, 128);
                        if (leader) mma_bf16(dcol, ad, bd, idesc, (j | kk) != 0);
                    }
                    if (leader) mma_commit(a_empty + 8 * slot);
                }
                if (leader) { mma_commit(t_full + 8 * buf); mma_commit(g_empty + 8 * buf); }
                __syncwarp();
            }
    } else if (warp >= 4 && warp < 8) {
        // ---------------- epilogue: max / arg-max of source 0, source 1 and their sum ----------------
        const int q = warp & 3;
        const int f = mtile * 128 + q * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t it = 0;
        for (int64_t u = rank; u < units; u += nranks) {
            const int b = (int)(u / p.Tp), tp = (int)(u % p.Tp);
            float best[3] = {-INFINITY, -INFINITY, -INFINITY};
            int bestt[3] = {0, 0, 0};
            for (int tt = 0; tt < tpu; ++tt, ++it) {
                const uint32_t buf = it & 1, ph = (it >> 1) & 1;
                const int t0 = (tp * tpu + tt) * FP_NT;
                mbar_wait(t_full + 8 * buf, ph);
                tc_fence_after();
                // accumulator columns: 16 jb + 8 s + r = source s at time t0 + 8 jb + r (time ascending within the scan)
#pragma unroll 1
                for (int c0 = 0; c0 < FT_NT; c0 += 64) {
                    uint32_t v0[32], v1[32];
                    tmem_ld32(lane_base + buf * FT_NT + c0, v0);
                    tmem_ld32(lane_base + buf * FT_NT + c0 + 32, v1);
                    tmem_ld_wait();
                    if (c0 + 64 == FT_NT) {   // accumulator drained: hand the buffer back
                        tc_fence_before();
                        mbar_arrive(t_empty + 8 * buf);
                    }
                    const int tb = t0 + c0 / 2;
#pragma unroll
                    for (int hb = 0; hb < 4; ++hb)
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            const uint32_t* v = hb < 2 ? v0 : v1;
                            const float a = __uint_as_float(v[16 * (hb & 1) + r]), c = __uint_as_float(v[16 * (hb & 1) + 8 + r]), m = a + c;
                            const int t = tb + 8 * hb + r;
                            if (a > best[1]) { best[1] = a; bestt[1] = t; }
                            if (c > best[2]) { best[2] = c; bestt[2] = t; }
                            if (m > best[0]) { best[0] = m; bestt[0] = t; }
                        }
                }
            }
            if (f < p.N) {
                const int rows[3] = {b, p.B + 2 * b, p.B + 2 * b + 1};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const size_t o = ((size_t)rows[k] * p.Tp + tp) * p.N + f;
                    p.y[o] = best[k];
                    if (p.argmax) p.argmax[o] = (int64_t)bestt[k] * p.N + f;
                }
            }
        }
    } else if (warp >= 8) {
        // ---------------- Hankel builders, both sources of the mixture ----------------
        const int bt = tid - 256;   // 0..127
        const int nunits = p.QB * 8;
        uint32_t it = 0;
        for (int64_t u = rank; u < units; u += nranks) {
            const int b = (int)(u / p.Tp), tp = (int)(u % p.Tp);
            const float* xr0 = p.x + (size_t)(p.B + 2 * b) * p.L;
            const float* xr1 = xr0 + p.L;
            for (int tt = 0; tt < tpu; ++tt, ++it) {
                const uint32_t buf = it & 1, ph = (it >> 1) & 1;
                const int t0 = (tp * tpu + tt) * FP_NT;
                for (int i = bt; i < ns; i += 128) {
                    const int s = t0 + i - p.pl;
                    const bool in = s >= 0 && s < p.L;
                    xs[i] = __float2bfloat16_rn(in ? __ldg(xr0 + s) : 0.f);
                    xs[ns + i] = __float2bfloat16_rn(in ? __ldg(xr1 + s) : 0.f);
                }
                named_bar_sync(1, 128);
                mbar_wait(g_empty + 8 * buf, ph ^ 1);
                const unsigned short* xsu = reinterpret_cast<const unsigned short*>(xs);
#pragma unroll
                for (int src = 0; src < 2; ++src) {
                    uint8_t* g = g_buf + buf * g_bytes + src * 128;          // interleaved: block m of source s at 256 m + 128 s
                    for (int i = bt; i < nunits; i += 128) {
                        uint32_t w[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            w[e] = (uint32_t)xsu[src * ns + i + 2 * e] | ((uint32_t)xsu[src * ns + i + 2 * e + 1] << 16);
                        *reinterpret_cast<uint4*>(g + (size_t)(i >> 3) * 256 + (i & 7) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
                fence_async_smem();
                mbar_arrive(g_full + 8 * buf);
                named_bar_sync(1, 128);   // xs is rewritten by the next tile
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 512);
}

}  // namespace

bool filterbank_analysis_tc_supported(int L, int W, int N, int pool, int hop, int mode) {
    (void)L;
    if (mode != AMSS_POOL_MAX) return false;
    if (pool != hop) return false;
    if (pool < 32) return false;
    if (pool >= FT_NT ? (pool % FT_NT != 0) : (FT_NT % pool != 0)) return false;
    if (W < 16 || W > 4096 || N < 8) return false;
    return true;
}

size_t filterbank_analysis_tc_workspace(int Bt, int L, int W, int N, int pool, int hop, int precision) {
    (void)Bt; (void)L; (void)pool; (void)hop; (void)precision;
    const int KS = (W + FT_KS - 1) / FT_KS, MT = (N + 127) / 128;
    return 512 + (size_t)MT * KS * FT_STAGE_BYTES;      // alignment slack + the mixture-check flag + packed filter stages
}

namespace {

int launch_stock(FtParams p, int Bt, int L, int pool, int hop, cudaStream_t st) {
    p.Bt = Bt;
    p.QB = p.KS * 8 + 31;
    const int64_t positions = (int64_t)p.Tp * pool;
    if (pool >= FT_NT) {
        p.tiles_per_unit = pool / FT_NT;
        p.units_per_signal = p.Tp;
    } else {
        p.tiles_per_unit = 1;
        p.units_per_signal = (int)((positions + FT_NT - 1) / FT_NT);
    }
    (void)L; (void)hop;
    const size_t smem = (size_t)FT_STAGES * FT_STAGE_BYTES + 2 * (size_t)p.QB * 128 + ((size_t)p.QB * 8 + 16) * 2;
    if (smem > 220 * 1024) { set_error("filterbank_analysis_tc: needs %zu B of shared memory", smem); return AMSS_ERR_UNSUPPORTED; }
    AMSS_CUDA(cudaFuncSetAttribute(analysis_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (kNumSMs / p.MT) * p.MT;
    AMSS_LAUNCH(analysis_tc_kernel, grid, FT_THREADS, smem, st, p);
    return AMSS_OK;
}

// common set-up: packs the filter, fills the shape fields
int prepare(FtParams& p, const float* x, const float* filt, int L, int W, int N, int pool, int hop, float* y, int64_t* argmax,
            void* workspace, size_t workspace_bytes, int Bt, int precision, cudaStream_t st) {
    if (!filterbank_analysis_tc_supported(L, W, N, pool, hop, AMSS_POOL_MAX)) {
        set_error("filterbank_analysis_tc: unsupported shape W=%d N=%d pool=%d hop=%d", W, N, pool, hop);
        return AMSS_ERR_UNSUPPORTED;
    }
    if (!workspace || workspace_bytes < filterbank_analysis_tc_workspace(Bt, L, W, N, pool, hop, precision)) {
        set_error("filterbank_analysis_tc: workspace too small");
        return AMSS_ERR_WORKSPACE;
    }
    p.KS = (W + FT_KS - 1) / FT_KS;
    p.MT = (N + 127) / 128;
    if (p.MT > kNumSMs) { set_error("filterbank_analysis_tc: too many filters"); return AMSS_ERR_UNSUPPORTED; }
    uint8_t* base = (uint8_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    uint8_t* packed = base + 256;                       // base[0..3] = the mixture-check flag
    {
        const int64_t units = (int64_t)p.MT * p.KS * (FT_STAGE_BYTES / 16);
        const int blocks = (int)std::min<int64_t>((units + 255) / 256, 4 * kNumSMs);
        AMSS_LAUNCH(pack_filter_kernel, blocks, 256, 0, st, filt, W, N, p.KS, p.MT, (uint4*)packed);
    }
    p.x = x; p.packed = packed; p.y = y; p.argmax = argmax;
    p.L = L; p.N = N; p.pool = pool;
    p.Tp = (L - pool) / hop + 1;
    p.pl = (W - 1) / 2;
    p.gate = nullptr; p.gate_zero = 0; p.B = 0;
    return AMSS_OK;
}

}  // namespace

int filterbank_analysis_tc(const float* x, const float* filt, int Bt, int L, int W, int N, int pool, int hop,
                           int precision, float* y, int64_t* argmax, void* workspace, size_t workspace_bytes,
                           cudaStream_t st) {
    FtParams p;
    const int rc = prepare(p, x, filt, L, W, N, pool, hop, y, argmax, workspace, workspace_bytes, Bt, precision, st);
    if (rc != AMSS_OK) return rc;
    return launch_stock(p, Bt, L, pool, hop, st);
}

// Front end of a training batch, x = [B mixtures ; B*S sources] (adapt.py:41-48).  With two sources per mixture the
// mixture rows are obtained by linearity whenever x_mix == x_0 + x_1 bit for bit (device-side check per batch);
// otherwise -- and for any other S -- all B*(S+1) rows go through analysis_tc_kernel.
bool filterbank_analysis_mix_tc_supported(int S, int L, int W, int N, int pool, int hop) {
    return S == 2 && filterbank_analysis_tc_supported(L, W, N, pool, hop, AMSS_POOL_MAX) && pool % FP_NT == 0;
}

int filterbank_analysis_mix_tc(const float* x, const float* filt, int B, int S, int L, int W, int N, int pool, int hop,
                               int precision, float* y, int64_t* argmax, void* workspace, size_t workspace_bytes,
                               cudaStream_t st) {
    const int Bt = B * (S + 1);
    FtParams p;
    int rc = prepare(p, x, filt, L, W, N, pool, hop, y, argmax, workspace, workspace_bytes, Bt, precision, st);
    if (rc != AMSS_OK) return rc;
    // AMSS_NO_LINEAR_MIX=1: always the stock three-signal kernel (A/B measurements; results agree within bf16 rounding)
    static const bool off = [] { const char* e = getenv("AMSS_NO_LINEAR_MIX"); return e && e[0] == '1'; }();
    if (off || !filterbank_analysis_mix_tc_supported(S, L, W, N, pool, hop)) return launch_stock(p, Bt, L, pool, hop, st);
    int* flag = reinterpret_cast<int*>(const_cast<uint8_t*>(p.packed) - 256);
    AMSS_CUDA(cudaMemsetAsync(flag, 0, 4, st));
    if ((L & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0)
        AMSS_LAUNCH(mix_is_sum_vec_kernel, 8 * kNumSMs, 256, 0, st, reinterpret_cast<const float4*>(x), B, (int64_t)(L / 4), flag);
    else
        AMSS_LAUNCH(mix_is_sum_kernel, 4 * kNumSMs, 256, 0, st, x, B, (int64_t)L, flag);
    FtParams q = p;                                   // linear-mixture kernel: runs iff no mismatch was found
    q.gate = flag; q.gate_zero = 1; q.B = B; q.Bt = Bt;
    q.QB = q.KS * 8 + FP_NT / 8 - 1;
    q.tiles_per_unit = pool / FP_NT;
    q.units_per_signal = q.Tp;
    const size_t smem = (size_t)FT_STAGES * FT_STAGE_BYTES + 4 * (size_t)q.QB * 128 + 2 * ((size_t)q.QB * 8 + 8) * 2;
    if (smem > 220 * 1024) return launch_stock(p, Bt, L, pool, hop, st);
    AMSS_CUDA(cudaFuncSetAttribute(analysis_pair_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (kNumSMs / q.MT) * q.MT;
    AMSS_LAUNCH(analysis_pair_tc_kernel, grid, FT_THREADS, smem, st, q);
    p.gate = flag; p.gate_zero = 0;                   // stock kernel: runs iff a mismatch was found
    return launch_stock(p, Bt, L, pool, hop, st);
}

}  // namespace amss
