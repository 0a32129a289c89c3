// BLSTM recurrence on the 5th-gen tensor cores (utils/ops.py:358-383, BasicLSTMCell i,j,f,o).
//
// One thread-block CLUSTER per (direction, sub-batch of NB mixtures); CTA c of the cluster owns
// hidden units [32c, 32c+32) for all T steps (persistent).  The recurrent weights never move after
// the prologue: each CTA's slice is converted to bf16 once and parked in TENSOR MEMORY as the A
// operand (tcgen05.st), so a time step's MMAs read only the tiny activation operand from shared
// memory (A-in-shared-memory MMAs at N = 16..64 cost ~100 clk each in operand fetch, measured).
//
// Forward step:  D[gate col, mixture] = W_h^T[own 128 gate cols, :] * h_{t-1}^T
//   * A (TMEM)  = the CTA's 128 gate columns of W_h (lane = gate*32 + unit), K = H;
//   * B (smem)  = h_{t-1} of the whole direction, [NB x H] bf16 K-major core matrices, double buffered
//     in every CTA.  After its cell update a CTA stages its 32 fresh h columns (a contiguous block of
//     the operand layout) and ONE thread pushes the block to every CTA of the cluster with
//     cp.async.bulk shared::cta -> shared::cluster; the copies complete_tx on the DESTINATION's
//     mbarrier, so the data path has no barrier.cluster at all -- the MMA warp of each CTA just waits
//     for NC blocks on its local mbarrier;
//   * D (TMEM)  = 128 lanes (gate columns) x NB columns (mixtures), fp32.
//   Fused epilogue: tcgen05.ld -> + hoisted input projection (loaded while the MMAs run) ->
//   sigmoid/tanh (MUFU.TANH) -> cell/hidden update with the cell state in registers for all T steps.
//
// Backward step (reverse time):  dh_t = dy_t + dz_{t+1} W_h^T.  CTA c multiplies ITS OWN 128 dz
//   columns (B operand, produced locally, never gathered) with the matching W_h columns
//   (A = W_h[all units, own cols] in TMEM, M = H in 128-row tiles, K = 128) and reduce-scatters the
//   bf16 partial sums to the unit owners with the same bulk-copy / remote-mbarrier scheme.
//
// Global stores (saved gates, c, y, dZ) are issued by dedicated WRITER warps from shared-memory
// staging, one step behind, so the recurrence never waits for global-memory traffic.
#include "common.cuh"
#include "tc.cuh"
#include <algorithm>
#include <cstdlib>

namespace amss {
namespace {

using namespace tc;

constexpr int BT_THREADS = 288;   // forward: warps 0-3 compute (TMEM lane quadrant = warp), 4 MMA / control (TMEM owner), 5-8 writers
constexpr int BW_THREADS = 256;   // backward: warps 0-3 compute (they issue their own MMAs), 4-7 writers (warp 4 owns the TMEM allocation);
                                  // 2 x 256 threads leave 128 registers per thread

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void bar_sync_named(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_named(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(tanh_fast(0.5f * x), 0.5f, 0.5f); }
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__host__ __device__ inline uint32_t pow2_cols(uint32_t c) { uint32_t r = 32; while (r < c) r <<= 1; return r; }

// Optional step profile (diagnostics): clock64() stamps of CTA 0, steps [PROF_S0, PROF_S0+4).
constexpr int PROF_S0 = 100, PROF_N = 4, PROF_K = 12;
long long* g_prof = nullptr;
long long* g_prof_bwd = nullptr;
// Optional schedule trace (diagnostics): per CTA {SM id, globaltimer at start, at end, at the end of the prologue}; forward at [0, 4*4096),
// backward at [4*4096, 8*4096).
long long* g_sched = nullptr;
constexpr int SCHED_MAX = 4096;
__device__ __forceinline__ long long globaltimer_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t sm_id() { uint32_t r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#define PROFB(k) do { if (prof0 && n >= PROF_S0 && n < PROF_S0 + PROF_N) prof0[(n - PROF_S0) * PROF_K + (k)] = clock64(); } while (0)
#define PROF(k) do { if (prof && s >= PROF_S0 && s < PROF_S0 + PROF_N) prof[(s - PROF_S0) * PROF_K + (k)] = clock64(); } while (0)

// =================================================================================================
// forward
// =================================================================================================
struct RecTcFwd {
    long long* prof;
    long long* sched;
    int dbg;              // experiments (AMSS_BLSTM_DBG): 1 = no global stores, 2 = no input-projection loads, 4 = TMA input loads
    const float* Wh[2];   // [H][ldw]
    int ldw;
    float* gates;         // [2][T][B][4H] in: hoisted input projection (+bias); out: activated gates
    float* cst;           // [2][T][B][H]
    float* y;             // [T][B][2H]
    int B, T, H, nsub;
    float forget_bias;
};

// TMEM: D at columns [0, NB), A at [FW_ACOL, FW_ACOL + 8*ksteps).  H = 300: 32 + 152 columns -> a 256-column allocation, so
// TWO CTAs fit the 512 columns of an SM (the NB = 16 variant runs two clusters per SM set: one CTA's exchange / MMA latency
// is covered by the other CTA's gate math).
constexpr uint32_t FW_ACOL = 32;
constexpr int FW_ZP = 36;   // row pitch (floats) of the [gate][mixture][32 units] staging tiles: 16-byte rows for the writers'
                            // vector accesses, and bank = 4*mixture + unit for the compute threads' scalar ones (conflict free)

// Lane assignment of the accumulator (= row order of the parked W_h slice): lane 32q + 8g + j holds gate g of unit 8q + j.
// tcgen05.ld.16x256b puts the TMEM lanes {l, l+8} x two loads (lanes 32q.. and 32q+16..) into ONE thread, so thread t of warp q
// receives all four gates of unit 8q + t/4 for the mixtures 8*rep + 2*(t%4) + {0,1}: the whole cell update runs in registers
// (no shared-memory transpose, no barrier between the activations and the cell update).
template <int NB>
__global__ void __launch_bounds__(BT_THREADS, NB == 16 ? 2 : 1) blstm_rec_fwd_tc_kernel(RecTcFwd p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[6];          // [0] MMA done, [1..2] h_full[buf], [3..5] zx_full[slot]
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = cluster_ctarank(), NC = cluster_nctarank();
    if (p.sched && tid == 0 && blockIdx.x < SCHED_MAX) {
        p.sched[4 * blockIdx.x] = sm_id(); p.sched[4 * blockIdx.x + 1] = globaltimer_ns();
    }
    const int cid = blockIdx.x / NC, d = cid / p.nsub, sub = cid % p.nsub;
    const int H = p.H, T = p.T, B = p.B, H4 = 4 * H;
    const int b0 = sub * NB, nvalid = min(NB, B - b0);
    const int u0 = crank * 32;
    const int KST = (H + 15) / 16;                // K steps of 16
    const int KCHB = NC * 4;                      // k-chunks (of 8) present in the h buffers
    constexpr int BG = NB / 8;                    // mixture groups of 8 (= repetitions of the TMEM load)
    constexpr int ZP = FW_ZP;
    constexpr uint32_t SLICE = 4 * BG * 128;      // bytes of one CTA's h block (32 units x NB mixtures, bf16)
    const uint32_t h_bytes = (uint32_t)KCHB * BG * 128;
    uint8_t* h_s = smem;                                              // [2][h_bytes]   B operand
    uint8_t* hst = h_s + 2 * h_bytes;                                 // [2][SLICE]     own block, source of the bulk copies
    float* og = reinterpret_cast<float*>(hst + 2 * SLICE);            // [6][NB][ZP]    activated gates i j f o, c, h -> writers
    float* zxs = og + 6 * NB * ZP;                                    // [3][4][NB][ZP] hoisted input projection, 3 steps ahead
    const uint32_t bar_mma = smem_u32(&bars[0]), h_full = smem_u32(&bars[1]), zx_full = smem_u32(&bars[3]);
    const uint32_t tcols = pow2_cols(FW_ACOL + 8 * KST);
    const bool tma_zx = (H & 3) == 0 && (p.dbg & 4);   // (experiment) TMA bulk copies for the input projection rows

    if (tid == 0) {
        mbar_init(bar_mma, 1); mbar_init(h_full, 1); mbar_init(h_full + 8, 1);
        const uint32_t zc = tma_zx ? 1 : 128;
        mbar_init(zx_full, zc); mbar_init(zx_full + 8, zc); mbar_init(zx_full + 16, zc);
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(smem_u32(&tmem_base_s), tcols);
    {   // zero the h operand buffers (h_{-1} = 0, padded units) and the staging tiles (rows / units the copies never write)
        const uint32_t total = (uint32_t)(reinterpret_cast<uint8_t*>(zxs + 3 * 4 * NB * ZP) - smem);
        for (uint32_t i = tid * 16; i < total; i += BT_THREADS * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (warp < 4) {   // park this CTA's W_h slice in TMEM:  A[32q + 8g + j][k] = Wh[k][g*H + u0 + 8q + j]
        const float* Wh = p.Wh[d];
        const int g = lane >> 3, u = u0 + 8 * warp + (lane & 7);
        for (int kk = 0; kk < KST; ++kk) {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = kk * 16 + 2 * j;
                const float e0 = (u < H && k < H) ? __ldg(Wh + (size_t)k * p.ldw + g * H + u) : 0.f;
                const float e1 = (u < H && k + 1 < H) ? __ldg(Wh + (size_t)(k + 1) * p.ldw + g * H + u) : 0.f;
                w[j] = pack_bf16(e0, e1);
            }
            tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + FW_ACOL + kk * 8, w);
        }
        tmem_st_wait();
    }
    if (tid == 0) {   // arm the first phase of both h buffers (h_0 -> buf 1, h_1 -> buf 0)
        if (T > 1) mbar_expect_tx(h_full + 8, NC * SLICE);
        if (T > 2) mbar_expect_tx(h_full, NC * SLICE);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                               // every CTA's barriers / buffers are ready for remote traffic
    tc_fence_after();
    const uint32_t idesc = idesc_bf16(128, NB, 0, 0);
    if (p.sched && tid == 0 && blockIdx.x < SCHED_MAX) p.sched[4 * blockIdx.x + 3] = globaltimer_ns();

    if (warp < 4) {
        // =========================== compute warps ===========================
        const int q = warp, j8 = lane >> 2, m4 = lane & 3;
        const int ul = 8 * q + j8;                    // this thread's unit; its mixtures are 8*rep + 2*m4 + e
        float creg[2 * BG];
#pragma unroll
        for (int i = 0; i < 2 * BG; ++i) creg[i] = 0.f;
        const float fb = p.forget_bias;
        long long* prof = (blockIdx.x == 0 && tid == 0) ? p.prof : nullptr;
        for (int s = 0; s < T; ++s) {
            PROF(0);
            // input projection of this step (staged three steps ago): fetched before the wait for the MMAs
            const int slot = s % 3;
            mbar_wait(zx_full + 8 * slot, (s / 3) & 1);
            float z[4][2 * BG];
            {
                const float* zb = zxs + slot * (4 * NB * ZP) + ul;
#pragma unroll
                for (int g = 0; g < 4; ++g)
#pragma unroll
                    for (int i = 0; i < 2 * BG; ++i) z[g][i] = zb[(g * NB + 8 * (i >> 1) + 2 * m4 + (i & 1)) * ZP];
            }
            uint32_t v0[4 * BG], v1[4 * BG];          // v0: gates i (r 0,1) / j (r 2,3);  v1: gates f / o
            if (s > 0) {
                mbar_wait(bar_mma, (s - 1) & 1);
                PROF(1);
                tc_fence_after();
                if (NB == 16) {
                    tmem_ld_16x256b_x2(tmem + ((uint32_t)(q * 32) << 16), v0);
                    tmem_ld_16x256b_x2(tmem + ((uint32_t)(q * 32 + 16) << 16), v1);
                } else {
                    tmem_ld_16x256b_x4(tmem + ((uint32_t)(q * 32) << 16), v0);
                    tmem_ld_16x256b_x4(tmem + ((uint32_t)(q * 32 + 16) << 16), v1);
                }
                tmem_ld_wait();
                tc_fence_before();
            } else {
#pragma unroll
                for (int i = 0; i < 4 * BG; ++i) v0[i] = v1[i] = 0u;
            }
            PROF(2);
            float hv[2 * BG];
#pragma unroll
            for (int i = 0; i < 2 * BG; ++i) {
                const int r = 4 * (i >> 1) + (i & 1);
                const float gi = sigmoid_fast(__uint_as_float(v0[r]) + z[0][i]);
                const float gj = tanh_fast(__uint_as_float(v0[r + 2]) + z[1][i]);
                const float gf = sigmoid_fast(__uint_as_float(v1[r]) + z[2][i] + fb);
                const float go = sigmoid_fast(__uint_as_float(v1[r + 2]) + z[3][i]);
                const float c = fmaf(creg[i], gf, gi * gj);
                creg[i] = c;
                hv[i] = (8 * (i >> 1) + 2 * m4 + (i & 1)) < nvalid ? tanh_fast(c) * go : 0.f;
                z[0][i] = gi; z[1][i] = gj; z[2][i] = gf; z[3][i] = go;
            }
            PROF(3);
            {   // h block in the operand layout; units are paired across lanes t, t^4 so every store is one 32-bit word
                uint8_t* hsl = hst + (s & 1) * SLICE + (size_t)q * BG * 128 + (2 * m4 + (j8 & 1)) * 16 + (j8 >> 1) * 4;
#pragma unroll
                for (int rep = 0; rep < BG; ++rep) {
                    const float mine = (j8 & 1) ? hv[2 * rep + 1] : hv[2 * rep];
                    const float give = (j8 & 1) ? hv[2 * rep] : hv[2 * rep + 1];
                    const float got = __shfl_xor_sync(0xffffffffu, give, 4);
                    *reinterpret_cast<uint32_t*>(hsl + rep * 128) = (j8 & 1) ? pack_bf16(got, mine) : pack_bf16(mine, got);
                }
            }
            PROF(5);
            fence_async_smem();                   // staged block -> visible to the bulk-copy engine
            bar_arrive_named(2, 160);             // hand the block to the MMA/control warp (no wait here)
            PROF(6);
            // ---- off the critical path: activated gates, c, h of this step -> staging for the writer warps ----
            if (s > 0) bar_sync_named(3, 256);    // the writers have read the previous step's tiles
            {
                float* ob = og + ul;
#pragma unroll
                for (int i = 0; i < 2 * BG; ++i) {
                    const int b = 8 * (i >> 1) + 2 * m4 + (i & 1);
#pragma unroll
                    for (int g = 0; g < 4; ++g) ob[(g * NB + b) * ZP] = z[g][i];
                    ob[(4 * NB + b) * ZP] = creg[i];
                    ob[(5 * NB + b) * ZP] = hv[i];
                }
            }
            bar_arrive_named(1, 256);             // tiles complete, zx slot consumed
            PROF(4);
        }
    } else if (warp == 4) {
        // =========================== MMA issuer (converged loop, elected lane) ===========================
        long long* prof = (blockIdx.x == 0 && lane == 0) ? p.prof : nullptr;
        const bool leader = elect_one();
        auto push_h = [&](int s) {                  // after the cell phase of step s: push the staged block to every CTA
            bar_sync_named(2, 160);
            if ((uint32_t)lane < NC && s + 1 < T) { // one bulk copy per lane: all NC pushes issue in parallel
                const uint32_t dst = smem_u32(h_s + ((s + 1) & 1) * h_bytes) + crank * SLICE;
                const uint32_t bar = h_full + 8 * ((s + 1) & 1);
                bulk_s2c(mapa(dst, lane), smem_u32(hst + (s & 1) * SLICE), SLICE, mapa(bar, lane));
            }
            __syncwarp();
        };
        push_h(0);
        for (int s = 1; s < T; ++s) {
            const uint32_t buf = s & 1;
            mbar_wait(h_full + 8 * buf, ((s - 1) >> 1) & 1);        // all NC blocks of h_{s-1} have landed
            PROF(9);
            if (leader && s + 2 < T) mbar_expect_tx(h_full + 8 * buf, NC * SLICE);   // re-arm for h_{s+1}
            tc_fence_after();
            const uint32_t haddr = smem_u32(h_s + buf * h_bytes);
            for (int kk = 0; kk < KST; ++kk) {
                const uint64_t bd = smem_desc(haddr + kk * 2 * BG * 128, BG * 128, 128);
                if (leader) mma_bf16_ts(tmem, tmem + FW_ACOL + kk * 8, bd, idesc, kk > 0);
            }
            if (leader) mma_commit(bar_mma);
            PROF(10);
            push_h(s);
        }
    } else {
        // =========================== writer warps: input projection in, saved gates / c / y out ===========================
        const int wt = tid - 160;                     // 0..127
        const int u4 = (wt & 7) * 4, rb = wt >> 3;    // unit quad, mixture row within a group of 16
        const bool vec = (H & 3) == 0 && u0 + u4 + 4 <= H;
        long long* prof = (blockIdx.x == 0 && wt == 0) ? p.prof : nullptr;
        // hoisted input projection (+bias) of step sn -> zxs[sn % 3], three steps ahead.  H % 4 == 0: one TMA bulk copy per
        // (gate, mixture) row of this CTA's units (<= 128 B), issued by one warp, completing on zx_full[slot] (the LSU never sees
        // them: per-thread cp.async copies stalled the writers ~1500 clk per step on their own outstanding-request limit).
        // Rows / units that do not exist are never written and stay zero from the prologue.  Other H: 4-byte cp.async copies.
        const int nu = min(32, H - u0);
        auto stage_zx = [&](int sn) {
            const int t = d == 0 ? sn : T - 1 - sn;
            const uint32_t bar = zx_full + 8 * (sn % 3);
            if (tma_zx) {
                if (wt < 32) {
                    if (wt == 0) mbar_expect_tx(bar, (uint32_t)(4 * nvalid * nu * 4));
                    __syncwarp();
                    const uint32_t sb = smem_u32(zxs + (sn % 3) * (4 * NB * ZP));
                    for (int r = wt; r < 4 * NB; r += 32) {
                        const int g4 = r / NB, b = r % NB;
                        if (b < nvalid)
                            bulk_g2s(sb + (g4 * NB + b) * ZP * 4, p.gates + (((size_t)d * T + t) * B + b0 + b) * H4 + g4 * H + u0, nu * 4, bar);
                    }
                }
                return;
            }
            const uint32_t sb = smem_u32(zxs + (sn % 3) * (4 * NB * ZP));
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4)
#pragma unroll
                for (int rr = 0; rr < NB / 16; ++rr) {
                    const int b = rr * 16 + rb;
                    const bool ok = b < nvalid;
                    const float* gi = ok ? p.gates + (((size_t)d * T + t) * B + b0 + b) * H4 + g4 * H + u0 + u4 : p.gates;
                    const uint32_t dst = sb + ((g4 * NB + b) * ZP + u4) * 4;
                    if (p.dbg & 2) continue;
                    if (vec) {
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gi), "r"(ok ? 16u : 0u) : "memory");
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const bool okj = ok && u0 + u4 + j < H;
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 4 * j), "l"(okj ? gi + j : p.gates), "r"(okj ? 4u : 0u) : "memory");
                        }
                    }
                }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
        };
        for (int n = 0; n < 3 && n < T; ++n) stage_zx(n);
        for (int s = 0; s < T; ++s) {
            const int t = d == 0 ? s : T - 1 - s;
            bar_sync_named(1, 256);                       // this step's tiles are complete; zx slot s%3 has been consumed
            long long tw0 = 0, tw1 = 0, tw2 = 0;          // profile stamps stay in registers until the stores are out
            if (prof) tw0 = clock64();
            if (!(p.dbg & 8) && s + 3 < T) stage_zx(s + 3);   // first: refill the slot just freed, three steps ahead (A/B: 428 -> 399 us)
            float4 gv[4 * (NB / 16)], cv[NB / 16], yv[NB / 16];
#pragma unroll
            for (int rr = 0; rr < NB / 16; ++rr) {
                const int b = rr * 16 + rb;
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) gv[g4 * (NB / 16) + rr] = *reinterpret_cast<const float4*>(og + (g4 * NB + b) * ZP + u4);
                cv[rr] = *reinterpret_cast<const float4*>(og + (4 * NB + b) * ZP + u4);
                yv[rr] = *reinterpret_cast<const float4*>(og + (5 * NB + b) * ZP + u4);
            }
            if (s + 1 < T) bar_arrive_named(3, 256);      // tiles are in registers: the compute warps may overwrite them
            if (prof) tw1 = clock64();
            if ((p.dbg & 8) && s + 3 < T) stage_zx(s + 3);
            if (prof) tw2 = clock64();
            auto put4 = [&](float* dst, const float4& v) {
                if (vec) __stcg(reinterpret_cast<float4*>(dst), v);
                else {
                    if (u0 + u4 < H) __stcg(dst, v.x);
                    if (u0 + u4 + 1 < H) __stcg(dst + 1, v.y);
                    if (u0 + u4 + 2 < H) __stcg(dst + 2, v.z);
                    if (u0 + u4 + 3 < H) __stcg(dst + 3, v.w);
                }
            };
#pragma unroll
            for (int rr = 0; rr < NB / 16; ++rr) {
                const int b = rr * 16 + rb;
                if (b < nvalid && !(p.dbg & 1)) {
                    const size_t row = ((size_t)d * T + t) * B + b0 + b;
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4) put4(p.gates + row * H4 + g4 * H + u0 + u4, gv[g4 * (NB / 16) + rr]);
                    put4(p.cst + row * H + u0 + u4, cv[rr]);
                    put4(p.y + ((size_t)t * B + b0 + b) * 2 * H + d * H + u0 + u4, yv[rr]);
                }
            }
            if (prof && s >= PROF_S0 && s < PROF_S0 + PROF_N) {
                long long* pr = prof + (s - PROF_S0) * PROF_K;
                pr[8] = clock64(); pr[7] = tw0; pr[11] = tw2; prof[48 + (s - PROF_S0)] = tw1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                               // no CTA exits while peers may still push into it
    if (p.sched && tid == 0 && blockIdx.x < SCHED_MAX) p.sched[4 * blockIdx.x + 2] = globaltimer_ns();
    if (warp == 4) tmem_dealloc(tmem, tcols);
}

size_t fwd_smem(int NC, int NB);
size_t bwd_smem(int NC, int NB);

template <int NB>
int launch_fwd(const RecTcFwd& p, int NC, cudaStream_t st) {
    const size_t smem = fwd_smem(NC, NB);
    if (smem > 226 * 1024) { set_error("blstm_rec_fwd_tc: H=%d NB=%d needs %zu B of shared memory", p.H, NB, smem); return AMSS_ERR_UNSUPPORTED; }
    AMSS_CUDA(cudaFuncSetAttribute(blstm_rec_fwd_tc_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AMSS_CUDA(cudaFuncSetAttribute(blstm_rec_fwd_tc_kernel<NB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (NC > 8) AMSS_CUDA(cudaFuncSetAttribute(blstm_rec_fwd_tc_kernel<NB>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(NC * 2 * p.nsub);
    cfg.blockDim = dim3(BT_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    AMSS_CUDA(cudaLaunchKernelEx(&cfg, blstm_rec_fwd_tc_kernel<NB>, p));
    count_launch();
    return AMSS_OK;
}

// =================================================================================================
// backward
// =================================================================================================
struct RecTcBwd {
    long long* prof;
    long long* sched;
    int dbg;              // experiments (AMSS_BLSTM_DBG): 1 = no global stores, 2 = no saved-activation loads
    const float* Wh[2];   // [H][ldw]
    int ldw;
    const float* gates;   // [2][T][B][4H] activated gates (saved by the forward pass)
    const float* cst;     // [2][T][B][H]
    const float* dy;      // [T][B][2H]
    float* dZ;            // [2][T][B][4H] fp32 (may be null when the bf16 copy + bias partials are requested instead)
    uint16_t* dZb;        // [2][T][B][ldzb] bf16 copy for the weight / input-gradient GEMMs (optional)
    int ldzb;
    float* dbpart;        // [2][nsub][4H] per-cluster column sums of dZ = bias-gradient partials (optional)
    int B, T, H, nsub, MT;
};

// TMEM: D tile m at [m*NB, +NB), A tile m at [acol + 64m, +64) with acol = MT*NB rounded up to 32 columns.  H = 300, NB = 16:
// 64 + 192 = 256 columns, two CTAs per SM.
__host__ __device__ inline uint32_t bw_acol(int MT, int NB) { return (uint32_t)((MT * NB + 31) & ~31); }

template <int NB>
__global__ void __launch_bounds__(BW_THREADS, NB == 16 ? 2 : 1) blstm_rec_bwd_tc_kernel(RecTcBwd p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[9];          // [0] MMAs done (all tiles), [4..5] r_full[buf], [6..8] sv_full[slot]
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = cluster_ctarank(), NC = cluster_nctarank();
    if (p.sched && tid == 0 && blockIdx.x < SCHED_MAX) {
        p.sched[4 * blockIdx.x] = sm_id(); p.sched[4 * blockIdx.x + 1] = globaltimer_ns();
    }
    const int cid = blockIdx.x / NC, d = cid / p.nsub, sub = cid % p.nsub;
    const int H = p.H, T = p.T, B = p.B, H4 = 4 * H, MT = p.MT;
    const int b0 = sub * NB, nvalid = min(NB, B - b0);
    const int u0 = crank * 32;
    constexpr int BG = NB / 8;
    constexpr int DZP = 132;                                             // row pitch (floats) of the dz staging tile [NB][128 gate cols]:
                                                                         // bank = 4*mixture + column for the compute warps' scalar stores,
                                                                         // 16-byte rows for the writers' vector loads
    constexpr uint32_t BLK = 32 * NB * 2;                                // one (src CTA -> dest CTA) block of bf16 partial sums:
                                                                         //   [NB/4 mixture quads][32 units][4 mixtures] (8-byte items)
    constexpr uint32_t ZLBO = BG * 128 + 16;                             // k-chunk stride of the dz operand, padded by 16 B: the four
                                                                         // 8-unit groups of a warp's 2-byte stores fall on different banks
    const uint32_t r_bytes = (uint32_t)NC * BLK;
    uint8_t* z_s = smem;                                                 // dz operand [16 k-chunks][NB][8] bf16, K-major
    uint8_t* r_s = z_s + 16 * ZLBO;                                      // [2][NC][BLK]   received partials (by source)
    uint8_t* p_s = r_s + 2 * r_bytes;                                    // [2][NC][BLK]   partials to send (by dest)
    float* dzs = reinterpret_cast<float*>(p_s + 2 * r_bytes);            // [2][NB][DZP] fp32 dz staging for the writers
    float* svs = dzs + 2 * NB * DZP;                                     // [3][7][NB][32] saved gates / c / c_prev / dy, 3 steps ahead (cp.async ring)
    const uint32_t bar_mma = smem_u32(&bars[0]), r_full = smem_u32(&bars[4]), sv_full = smem_u32(&bars[6]);
    const uint32_t BW_ACOL = bw_acol(MT, NB);
    const uint32_t tcols = pow2_cols(BW_ACOL + 64 * MT);

    if (tid == 0) {
        mbar_init(bar_mma, MT);                  // one tcgen05.commit per issuing warp (tile)
        mbar_init(r_full, 1); mbar_init(r_full + 8, 1);
        mbar_init(sv_full, 128); mbar_init(sv_full + 8, 128); mbar_init(sv_full + 16, 128);
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(smem_u32(&tmem_base_s), tcols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (warp < 4) {   // park A[u][g] = Wh[u][gate(g)*H + u0 + g%32] in TMEM (lane = unit of the M tile)
        const float* Wh = p.Wh[d];
        const bool vec = (H & 3) == 0 && (p.ldw & 3) == 0;
        for (int m = 0; m < MT; ++m) {
            const int u = m * 128 + warp * 32 + lane;
            for (int kk = 0; kk < 8; ++kk) {
                const int gate = kk >> 1, c0 = u0 + (kk & 1) * 16;
                const float* src = Wh + (size_t)u * p.ldw + gate * H + c0;
                float e[16];
                if (vec && u < H && c0 + 16 <= H) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + j);
                        e[4 * j] = v.x; e[4 * j + 1] = v.y; e[4 * j + 2] = v.z; e[4 * j + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) e[j] = (u < H && c0 + j < H) ? __ldg(src + j) : 0.f;
                }
                uint32_t w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] = pack_bf16(e[2 * j], e[2 * j + 1]);
                tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + BW_ACOL + m * 64 + kk * 8, w);
            }
        }
        tmem_st_wait();
    }
    if (tid == 0) {   // arm the first phase of both receive buffers (step 1 -> buf 1, step 2 -> buf 0)
        if (T > 1) mbar_expect_tx(r_full + 8, NC * BLK);
        if (T > 2) mbar_expect_tx(r_full, NC * BLK);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t idesc = idesc_bf16(128, NB, 0, 0);
    if (p.sched && tid == 0 && blockIdx.x < SCHED_MAX) p.sched[4 * blockIdx.x + 3] = globaltimer_ns();

    if (warp < 4) {
        // =========================== compute warps ===========================
        const int q = warp;
        constexpr int IT = NB / 4;                    // mixtures per thread: b = q*IT + i, unit = lane
        float dcc[IT];
#pragma unroll
        for (int i = 0; i < IT; ++i) dcc[i] = 0.f;
        long long* prof0 = (blockIdx.x == 0 && tid == 0) ? p.prof : nullptr;
        const bool leader = elect_one();
        for (int s = T - 1; s >= 0; --s) {
            const int n = T - 1 - s;                  // step counter
            PROFB(0);
            if (n > 0 && q < MT) {
                // MMAs of tile q, issued by THIS warp: chains issued by several warps in parallel run at the tensor pipe's rate
                // (~36 clk per N = 16 MMA) instead of one thread's issue rate (~55 clk, tools/ld_probe.cu)
                PROFB(9);
                tc_fence_after();
                const uint32_t zaddr = smem_u32(z_s);
                for (int kk = 0; kk < 8; ++kk) {
                    const uint64_t bd = smem_desc(zaddr + kk * 2 * ZLBO, ZLBO, 128);
                    if (leader) mma_bf16_ts(tmem + q * NB, tmem + BW_ACOL + q * 64 + kk * 8, bd, idesc, kk > 0);
                }
                if (leader) mma_commit(bar_mma);
                __syncwarp();
                PROFB(10);
            }
            if (n > 0) {
                // partial sums P[u, b] = sum_{own cols} Wh[u, g] dz_{next}[b, g]: tile m (units m*128 + q*32 + lane) belongs to the
                // owner CTA m*4 + q and is sent by THIS warp (no other warp is involved).  All tiles are drained together: one
                // barrier wait, back-to-back TMEM loads, one proxy fence (the chains of the MT issuing warps finish together anyway).
                // Sending with st.async straight from registers (no staging, no fence) leaves the sender 400 clk earlier but the
                // partial sums land at the same time: the exchange is bound by the SM's DSMEM port (~20 KB in + out per step), A/B
                // measured 515 (bulk copies) vs 544 us (st.async) per launch at 128 mixtures.
                uint8_t* ps = p_s + (n & 1) * r_bytes;
                mbar_wait(bar_mma, (n - 1) & 1);      // ONE barrier: every issuing warp's commit arrives on it
                PROFB(2);
                tc_fence_after();
                uint32_t acc[3][NB];                  // (a fourth tile, H > 384, is drained by the block below)
#pragma unroll
                for (int m = 0; m < 3; ++m)
                    if (m < MT && (uint32_t)(m * 4 + q) < NC) {           // warp-uniform
                        if (NB == 16) tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + m * NB, acc[m]);
                        else tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + m * NB, acc[m]);
                    }
                tmem_ld_wait();
                tc_fence_before();
                PROFB(11);
                auto stage_tile = [&](const uint32_t* av, uint32_t dest) {
                    uint8_t* pd = ps + (size_t)dest * BLK + lane * 8;
#pragma unroll
                    for (int g = 0; g < NB / 4; ++g)
                        *reinterpret_cast<uint2*>(pd + g * 256) =
                            make_uint2(pack_bf16(__uint_as_float(av[4 * g]), __uint_as_float(av[4 * g + 1])),
                                       pack_bf16(__uint_as_float(av[4 * g + 2]), __uint_as_float(av[4 * g + 3])));
                };
#pragma unroll
                for (int m = 0; m < 3; ++m)
                    if (m < MT && (uint32_t)(m * 4 + q) < NC) stage_tile(acc[m], m * 4 + q);
                if (MT > 3 && (uint32_t)(12 + q) < NC) {
                    tc_fence_after();
                    if (NB == 16) tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + 3 * NB, acc[0]);
                    else tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + 3 * NB, acc[0]);
                    tmem_ld_wait();
                    tc_fence_before();
                    stage_tile(acc[0], 12 + q);
                }
                fence_async_smem();
                __syncwarp();
                PROFB(8);
                for (int m = 0; m < MT; ++m) {        // uniform control flow, elected lane (a divergent push serialises per lane)
                    const uint32_t dest = m * 4 + q;
                    if (dest < NC && leader)
                        bulk_s2c(mapa(smem_u32(r_s + (n & 1) * r_bytes) + crank * BLK, dest), smem_u32(ps) + dest * BLK, BLK,
                                 mapa(r_full + 8 * (n & 1), dest));
                }
                __syncwarp();
                PROFB(3);
            }
            // While the partial sums travel: everything of this step that does not depend on dh.  With A = go (1 - tanh(c)^2) the update is
            //   dc = dcc + dh A,  dz_i = dc F0,  dz_j = dc F1,  dz_f = dc F2,  dz_o = dh F3,  dcc' = dc gf
            // so only ~6 flops per (unit, mixture) remain between the arrival of the partial sums and the next step's MMAs.
            float fA[IT], f0[IT], f1[IT], f2[IT], f3[IT], fgf[IT], dyv[IT];
            {
                mbar_wait(sv_full + 8 * (n % 3), (n / 3) & 1);    // staged three steps ago by the writer warps (cp.async)
                const float* svb = svs + (n % 3) * (7 * NB * 32) + (q * IT) * 32 + lane;
#pragma unroll
                for (int i = 0; i < IT; ++i) {
                    const float gi = svb[(0 * NB + i) * 32], gj = svb[(1 * NB + i) * 32], gf = svb[(2 * NB + i) * 32], go = svb[(3 * NB + i) * 32];
                    const float c = svb[(4 * NB + i) * 32], cprev = svb[(5 * NB + i) * 32];
                    dyv[i] = svb[(6 * NB + i) * 32];
                    const float tc_ = tanh_fast(c);
                    fA[i] = go * (1.f - tc_ * tc_);
                    f0[i] = gj * gi * (1.f - gi);
                    f1[i] = gi * (1.f - gj * gj);
                    f2[i] = cprev * gf * (1.f - gf);
                    f3[i] = tc_ * go * (1.f - go);
                    fgf[i] = gf;
                    // pin the factors HERE (the compiler otherwise sinks this arithmetic below the wait for the partial sums,
                    // back onto the critical path)
                    asm volatile("" : "+f"(fA[i]), "+f"(f0[i]), "+f"(f1[i]), "+f"(f2[i]), "+f"(f3[i]), "+f"(dyv[i]));
                }
            }
            PROFB(1);
            float dh[IT];
#pragma unroll
            for (int i = 0; i < IT; ++i) dh[i] = dyv[i];
            if (n > 0) {
                mbar_wait(r_full + 8 * (n & 1), ((n - 1) >> 1) & 1);       // every CTA's partials for my units have landed
                PROFB(4);
                if (tid == 0 && n + 2 < T) mbar_expect_tx(r_full + 8 * (n & 1), NC * BLK);   // re-arm for step n+2
                // this thread's unit = lane, mixtures q*IT .. q*IT+IT-1: one 8-byte item per mixture quad, lanes consecutive.
                // All loads of a source batch are issued (volatile asm: not re-ordered, not serialised behind the adds) before the
                // first add; sources are summed in index order (deterministic).
                const uint32_t rb = smem_u32(r_s + (n & 1) * r_bytes) + (uint32_t)(q * (IT / 4) * 32 + lane) * 8;
                for (uint32_t c0 = 0; c0 < NC; c0 += 5) {
                    uint32_t wx[5][IT / 4], wy[5][IT / 4];
#pragma unroll
                    for (uint32_t c = 0; c < 5; ++c) {
                        const uint32_t cc = c0 + c < NC ? c0 + c : c0;        // (a batch past the end re-reads its first block, masked below)
#pragma unroll
                        for (int g = 0; g < IT / 4; ++g)
                            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(wx[c][g]), "=r"(wy[c][g]) : "r"(rb + cc * BLK + g * 256));
                    }
#pragma unroll
                    for (int g = 0; g < IT / 4; ++g)     // every load of the batch is in flight before the first add may issue
                        asm volatile("" : "+r"(wx[0][g]), "+r"(wy[0][g]), "+r"(wx[1][g]), "+r"(wy[1][g]), "+r"(wx[2][g]), "+r"(wy[2][g]),
                                          "+r"(wx[3][g]), "+r"(wy[3][g]), "+r"(wx[4][g]), "+r"(wy[4][g]));
#pragma unroll
                    for (uint32_t c = 0; c < 5; ++c) {
                        if (c0 + c < NC) {
#pragma unroll
                            for (int g = 0; g < IT / 4; ++g) {
                                dh[4 * g] += bf16_lo(wx[c][g]); dh[4 * g + 1] += bf16_hi(wx[c][g]);
                                dh[4 * g + 2] += bf16_lo(wy[c][g]); dh[4 * g + 3] += bf16_hi(wy[c][g]);
                            }
                        }
                    }
                }
                PROFB(5);
            }
            // gate derivatives; dz -> fp32 staging (writers) and bf16 operand of the next step's MMA.  All 8 * IT store addresses are
            // one base register + immediates: k = g4*32 + lane, b = q*IT + i  ->  operand offset (k>>3)*ZLBO + (b>>3)*128 + (b&7)*16
            // + (k&7)*2 = zbase + g4*4*ZLBO + i*16 (the IT mixtures of a thread never straddle a group of 8)
            float* dzb = dzs + (n & 1) * (NB * DZP) + (q * IT) * DZP + lane;
            uint8_t* zbase = z_s + (size_t)(lane >> 3) * ZLBO + (lane & 7) * 2 + ((q * IT) >> 3) * 128 + ((q * IT) & 7) * 16;
#pragma unroll
            for (int i = 0; i < IT; ++i) {
                const float dc = fmaf(dh[i], fA[i], dcc[i]);
                float dz[4];
                dz[0] = dc * f0[i];
                dz[1] = dc * f1[i];
                dz[2] = dc * f2[i];
                dz[3] = dh[i] * f3[i];
                dcc[i] = dc * fgf[i];
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                    dzb[i * DZP + g4 * 32] = dz[g4];
                    *reinterpret_cast<__nv_bfloat16*>(zbase + g4 * 4 * ZLBO + i * 16) = __float2bfloat16_rn(dz[g4]);
                }
            }
            PROFB(6);
            fence_async_smem();
            bar_sync_named(1, 256);                   // dz staged by every warp: the next step's MMAs may issue, writers may store
            PROFB(7);
        }
    } else {
        // =========================== writer warps: dZ -> global ===========================
        const int wt = tid - 128, wq = wt >> 5, wu = u0 + (wt & 31);
        // saved activations of step n (time s = T-1-n): rows k = gi gj gf go c c_prev dy, 32 units each (128 B per mixture):
        // 16-byte loads, thread = (4 consecutive units, mixture rows rb, rb+16, ...), issued two steps ahead
        const int u4 = (wt & 7) * 4, rb = wt >> 3;
        const bool vec = (H & 3) == 0 && u0 + u4 + 4 <= H;
        auto row_ptr = [&](int k, int s, int b) -> const float* {
            const int t = d == 0 ? s : T - 1 - s;
            const int tprev = d == 0 ? t - 1 : t + 1;
            const size_t row = ((size_t)d * T + t) * B + b0 + b;
            if (k < 4) return p.gates + row * H4 + k * H + u0 + u4;
            if (k == 4) return p.cst + row * H + u0 + u4;
            if (k == 5) return p.cst + (((size_t)d * T + tprev) * B + b0 + b) * H + u0 + u4;
            return p.dy + ((size_t)t * B + b0 + b) * 2 * H + d * H + u0 + u4;
        };
        // asynchronous global -> shared copies (no registers, the writer never waits for them); completion arrives on
        // sv_full[slot].  Invalid rows / units are zero-filled (src-size 0).
        auto stage_sv = [&](int n) {
            const int s = T - 1 - n;
            const uint32_t sb = smem_u32(svs + (n % 3) * (7 * NB * 32));
#pragma unroll
            for (int k = 0; k < 7; ++k)
#pragma unroll
                for (int rr = 0; rr < NB / 16; ++rr) {
                    const int b = rr * 16 + rb;
                    const bool ok = b < nvalid && !(k == 5 && s == 0) && !(p.dbg & 2);
                    const float* gi = ok ? row_ptr(k, s, b) : p.cst;
                    const uint32_t dst = sb + ((k * NB + b) * 32 + u4) * 4;
                    if (vec) {
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gi), "r"(ok ? 16u : 0u) : "memory");
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const bool okj = ok && u0 + u4 + j < H;
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 4 * j), "l"(okj ? gi + j : p.cst), "r"(okj ? 4u : 0u) : "memory");
                        }
                    }
                }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(sv_full + 8 * (n % 3)) : "memory");
        };
        // dZ of a step leaves as 8-byte (4 x bf16) / 16-byte (4 x fp32) stores: thread = (mixture rb [+16], gate jj, units u4..u4+3),
        // the 8 threads of a mixture row cover one gate's 32 units (64 B / 128 B contiguous).  Per-thread column sums over t and
        // the thread's mixtures feed the bias gradient (combined across the mixture rows at the end, in a fixed order).
        float bsum[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) bsum[i] = 0.f;
        const bool vec8 = vec && (p.ldzb & 3) == 0;
        for (int n = 0; n < 3 && n < T; ++n) stage_sv(n);
        for (int s = T - 1; s >= 0; --s) {
            const int t = d == 0 ? s : T - 1 - s, n = T - 1 - s;
            bar_sync_named(1, 256);                               // dz(n) staged; svs[n%3] consumed by the compute warps
            if (n + 3 < T) stage_sv(n + 3);                       // refill the slot just freed, three steps ahead
            const float* dzb = dzs + (n & 1) * (NB * DZP);
            const size_t r0 = ((size_t)d * T + t) * B + b0;
#pragma unroll
            for (int rr = 0; rr < NB / 16; ++rr) {
                const int b = rr * 16 + rb;
                float4 v[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) v[jj] = *reinterpret_cast<const float4*>(dzb + b * DZP + jj * 32 + u4);
                if (b < nvalid) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        bsum[4 * jj] += v[jj].x; bsum[4 * jj + 1] += v[jj].y; bsum[4 * jj + 2] += v[jj].z; bsum[4 * jj + 3] += v[jj].w;
                        if (p.dbg & 1) continue;
                        if (p.dZ) {
                            float* zo = p.dZ + (r0 + b) * H4 + jj * H + u0 + u4;
                            if (vec) __stcg(reinterpret_cast<float4*>(zo), v[jj]);
                            else {
                                if (u0 + u4 < H) __stcg(zo, v[jj].x);
                                if (u0 + u4 + 1 < H) __stcg(zo + 1, v[jj].y);
                                if (u0 + u4 + 2 < H) __stcg(zo + 2, v[jj].z);
                                if (u0 + u4 + 3 < H) __stcg(zo + 3, v[jj].w);
                            }
                        }
                        if (p.dZb) {
                            __nv_bfloat16* zb = reinterpret_cast<__nv_bfloat16*>(p.dZb) + (r0 + b) * p.ldzb + jj * H + u0 + u4;
                            if (vec8) *reinterpret_cast<uint2*>(zb) = make_uint2(pack_bf16(v[jj].x, v[jj].y), pack_bf16(v[jj].z, v[jj].w));
                            else {
                                if (u0 + u4 < H) zb[0] = __float2bfloat16_rn(v[jj].x);
                                if (u0 + u4 + 1 < H) zb[1] = __float2bfloat16_rn(v[jj].y);
                                if (u0 + u4 + 2 < H) zb[2] = __float2bfloat16_rn(v[jj].z);
                                if (u0 + u4 + 3 < H) zb[3] = __float2bfloat16_rn(v[jj].w);
                            }
                        }
                    }
                }
            }
        }
        if (p.dbpart) {   // column sums: [16 mixture rows][128 gate cols] partials through the (now idle) dz staging tile
            float* red = dzs + (((T - 1) & 1) ^ 1) * (NB * DZP);          // the tile the last step did not use
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
                *reinterpret_cast<float4*>(red + rb * DZP + jj * 32 + u4) = make_float4(bsum[4 * jj], bsum[4 * jj + 1], bsum[4 * jj + 2], bsum[4 * jj + 3]);
            bar_sync_named(3, 128);
            float tot = 0.f;
#pragma unroll
            for (int r = 0; r < 16; ++r) tot += red[r * DZP + wt];
            if (wu < H) p.dbpart[((size_t)d * p.nsub + sub) * H4 + wq * H + wu] = tot;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (p.sched && tid == 0 && blockIdx.x < SCHED_MAX) p.sched[4 * blockIdx.x + 2] = globaltimer_ns();
    if (warp == 4) tmem_dealloc(tmem, tcols);
}

template <int NB>
int launch_bwd(const RecTcBwd& p, int NC, cudaStream_t st) {
    const size_t smem = bwd_smem(NC, NB);
    if (smem > 226 * 1024) { set_error("blstm_rec_bwd_tc: H=%d NB=%d needs %zu B of shared memory", p.H, NB, smem); return AMSS_ERR_UNSUPPORTED; }
    AMSS_CUDA(cudaFuncSetAttribute(blstm_rec_bwd_tc_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AMSS_CUDA(cudaFuncSetAttribute(blstm_rec_bwd_tc_kernel<NB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (NC > 8) AMSS_CUDA(cudaFuncSetAttribute(blstm_rec_bwd_tc_kernel<NB>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(NC * 2 * p.nsub);
    cfg.blockDim = dim3(BW_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    AMSS_CUDA(cudaLaunchKernelEx(&cfg, blstm_rec_bwd_tc_kernel<NB>, p));
    count_launch();
    return AMSS_OK;
}

// Co-resident clusters are limited by the GPC size (a cluster of 10 CTAs fits once in a 16-20 SM GPC), not by
// SMs / NC: ask the occupancy API, then take the smallest sub-batch (16 / 32 / 64 mixtures) that runs in one wave.
template <typename K>
int max_clusters(K kernel, int NC, size_t smem, int threads) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(NC * 16);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (NC > 8) cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n < 2) { cudaGetLastError(); n = std::max(2, kNumSMs / NC / 2); }
    return n;
}

size_t fwd_smem(int NC, int NB) {
    return 2 * (size_t)NC * 4 * (NB / 8) * 128 + 2 * (size_t)4 * (NB / 8) * 128 + (size_t)(6 + 12) * NB * FW_ZP * 4;
}
size_t bwd_smem(int NC, int NB) {
    return (size_t)16 * ((NB / 8) * 128 + 16) + 4 * (size_t)NC * 32 * NB * 2 + (size_t)2 * NB * 132 * 4 + (size_t)3 * 7 * NB * 32 * 4;
}

// Sub-batch size.  Measured on B200 (tools/blstm_bench.py, T = 250, I = 600, H = 300; fwd + bwd layer times in ms):
//   B = 32: 1.36 (NB 16) vs 2.00 (NB 32);  64: 1.43 vs 2.06;  128: 1.83 vs 2.21;  256: 3.27 vs 4.25
// -- the 16-mixture clusters win at every batch size, also where they need more than one wave of co-resident clusters
// (a step of an NB = 16 cluster is ~1.6x shorter and two of its CTAs share an SM), so NB = 16 is the default whenever
// the kernel fits; AMSS_BLSTM_NB=32 forces the one-CTA-per-SM variant (A/B runs).
int pick_nb(int B, int maxc16, int maxc32) {
    (void)B; (void)maxc16; (void)maxc32;
    if (const char* e = getenv("AMSS_BLSTM_NB")) { const int v = atoi(e); if (v == 16 || v == 32) return v; }
    return 16;
}

struct MaxC { int c16 = 0, c32 = 0; };

}  // namespace

void blstm_tc_set_sched(long long* dev_buf) { g_sched = dev_buf; }
void blstm_tc_set_profile(long long* dev_buf) { g_prof = dev_buf; g_prof_bwd = dev_buf ? dev_buf + 64 : nullptr; }

bool blstm_rec_tc_supported(int B, int T, int H) {
    (void)B; (void)T;
    const int NC = (H + 31) / 32;
    return H >= 8 && NC <= 16;
}

int blstm_rec_fwd_tc(const float* Wh_fw, const float* Wh_bw, int ldw, float* gates, float* cst, float* y, int B, int T,
                     int H, float forget_bias, cudaStream_t st) {
    const int NC = (H + 31) / 32;
    RecTcFwd p;
    p.prof = g_prof;
    p.sched = g_sched;
    { const char* e = getenv("AMSS_BLSTM_DBG"); p.dbg = e ? atoi(e) : 0; }
    p.Wh[0] = Wh_fw; p.Wh[1] = Wh_bw; p.ldw = ldw; p.gates = gates; p.cst = cst; p.y = y;
    p.B = B; p.T = T; p.H = H; p.forget_bias = forget_bias;
    static MaxC mc[17];
    if (!mc[NC].c16) {
        mc[NC].c16 = max_clusters(blstm_rec_fwd_tc_kernel<16>, NC, fwd_smem(NC, 16), BT_THREADS);
        mc[NC].c32 = max_clusters(blstm_rec_fwd_tc_kernel<32>, NC, fwd_smem(NC, 32), BT_THREADS);
    }
    const int nb = pick_nb(B, mc[NC].c16, mc[NC].c32);   // larger batches run as several waves of clusters
    p.nsub = (B + nb - 1) / nb;
    if (nb == 16) return launch_fwd<16>(p, NC, st);
    return launch_fwd<32>(p, NC, st);
}

// Sub-batches (clusters per direction) the backward recurrence will run with: sizes the bias-partial buffer.
int blstm_rec_bwd_tc_nsub(int B, int H);

static int bwd_nb(int B, int NC) {
    static MaxC mc[17];
    if (!mc[NC].c16) {
        mc[NC].c16 = max_clusters(blstm_rec_bwd_tc_kernel<16>, NC, bwd_smem(NC, 16), BW_THREADS);
        mc[NC].c32 = max_clusters(blstm_rec_bwd_tc_kernel<32>, NC, bwd_smem(NC, 32), BW_THREADS);
    }
    return pick_nb(B, mc[NC].c16, mc[NC].c32);
}

int blstm_rec_bwd_tc(const float* Wh_fw, const float* Wh_bw, int ldw, const float* gates, const float* cst,
                     const float* dy, float* dZ, uint16_t* dZb, int ldzb, float* dbpart, int B, int T, int H, cudaStream_t st) {
    const int NC = (H + 31) / 32;
    RecTcBwd p;
    p.prof = g_prof_bwd;
    p.sched = g_sched ? g_sched + 4 * SCHED_MAX : nullptr;
    { const char* e = getenv("AMSS_BLSTM_DBG"); p.dbg = e ? atoi(e) : 0; }
    p.Wh[0] = Wh_fw; p.Wh[1] = Wh_bw; p.ldw = ldw; p.gates = gates; p.cst = cst; p.dy = dy; p.dZ = dZ;
    p.dZb = dZb; p.ldzb = ldzb; p.dbpart = dbpart;
    p.B = B; p.T = T; p.H = H; p.MT = (NC * 32 + 127) / 128;
    const int nb = bwd_nb(B, NC);
    p.nsub = (B + nb - 1) / nb;
    if (nb == 16) return launch_bwd<16>(p, NC, st);
    return launch_bwd<32>(p, NC, st);
}

int blstm_rec_bwd_tc_nsub(int B, int H) {
    const int NC = (H + 31) / 32;
    const int nb = bwd_nb(B, NC);
    return (B + nb - 1) / nb;
}

// co-resident clusters of both variants (diagnostics: amss_debug_blstm_clusters)
void blstm_tc_max_clusters(int H, int* out4) {
    const int NC = (H + 31) / 32;
    out4[0] = max_clusters(blstm_rec_fwd_tc_kernel<16>, NC, fwd_smem(NC, 16), BT_THREADS);
    out4[1] = max_clusters(blstm_rec_fwd_tc_kernel<32>, NC, fwd_smem(NC, 32), BT_THREADS);
    out4[2] = max_clusters(blstm_rec_bwd_tc_kernel<16>, NC, bwd_smem(NC, 16), BW_THREADS);
    out4[3] = max_clusters(blstm_rec_bwd_tc_kernel<32>, NC, bwd_smem(NC, 32), BW_THREADS);
}

}  // namespace amss
