// BLSTM recurrence on the 5th-gen tensor cores (utils/ops.py:358-383, BasicLSTMCell i,j,f,o).
//
// One thread-block CLUSTER per (direction, sub-batch of NB mixtures); CTA c of the cluster owns
// hidden units [32c, 32c+32) for all T steps (persistent, weights resident in shared memory).
//
// Forward step:  D[gate col, mixture] = W_h^T[own 128 gate cols, :] * h_{t-1}^T
//   * A operand  = the CTA's 128 gate columns of W_h (lane = gate*32 + unit), bf16, K-major
//     core-matrix layout, packed once;
//   * B operand  = h_{t-1} of the whole direction, [NB x H] bf16 K-major, double buffered in every
//     CTA; each CTA scatters its 32 fresh h columns into ALL CTAs of the cluster with
//     st.shared::cluster (DSMEM) and ONE barrier.cluster per step publishes them;
//   * D in TMEM: 128 lanes (gate columns) x NB columns (mixtures), fp32.
//   Fused epilogue: tcgen05.ld -> + hoisted input projection (prefetched one step ahead) ->
//   sigmoid/tanh (MUFU.TANH) -> cell/hidden update with the cell state in registers for all T steps.
//
// Backward step (reverse time):  dh_t = dy_t + dz_{t+1} W_h^T.  CTA c multiplies ITS OWN 128 dz
//   columns (B operand, produced locally, never gathered) with the matching W_h columns
//   (A = W_h[all units, own cols], M = H in 128-row tiles, K = 128) and reduce-scatters the partial
//   sums (bf16) through DSMEM to the CTAs owning the units; again one barrier.cluster per step.
//
// Global stores (saved gates, c, y, dZ) are issued by dedicated WRITER warps that join the cluster
// barrier with .relaxed semantics, so the release fence of the compute warps never waits for
// outstanding global-memory traffic.
#include "common.cuh"
#include "tc.cuh"
#include <algorithm>

namespace amss {
namespace {

using namespace tc;

constexpr int BT_THREADS = 288;
constexpr int FW_NACC = 1;        // forward: independent TMEM accumulators (K split round-robin), summed in the epilogue   // warps 0-3 compute (TMEM lane quadrant = warp), 4 MMA/control, 5-8 writers

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, uint2 v) {
    asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_cluster() { asm volatile("fence.proxy.async.shared::cluster;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_named(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_named(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(tanh_fast(0.5f * x), 0.5f, 0.5f); }
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// =================================================================================================
// forward
// =================================================================================================
// Optional step profile (diagnostics): clock64() stamps of CTA 0, steps [PROF_S0, PROF_S0+4).
constexpr int PROF_S0 = 100, PROF_N = 4, PROF_K = 12;
long long* g_prof = nullptr;
#define PROF(k) do { if (prof && s >= PROF_S0 && s < PROF_S0 + PROF_N) prof[(s - PROF_S0) * PROF_K + (k)] = clock64(); } while (0)

struct RecTcFwd {
    long long* prof;
    const float* Wh[2];   // [H][ldw]
    int ldw;
    float* gates;         // [2][T][B][4H] in: hoisted input projection (+bias); out: activated gates
    float* cst;           // [2][T][B][H]
    float* y;             // [T][B][2H]
    int B, T, H, nsub;
    float forget_bias;
};

template <int NB>
__global__ void __launch_bounds__(BT_THREADS, 1) blstm_rec_fwd_tc_kernel(RecTcFwd p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = cluster_ctarank(), NC = cluster_nctarank();
    const int cid = blockIdx.x / NC, d = cid / p.nsub, sub = cid % p.nsub;
    const int H = p.H, T = p.T, B = p.B, H4 = 4 * H;
    const int b0 = sub * NB, nvalid = min(NB, B - b0);
    const int u0 = crank * 32;
    const int KCH = (H + 15) / 16 * 2;            // k-chunks (of 8) consumed by the MMAs
    const int KCHB = NC * 4;                      // k-chunks present in the h buffers
    constexpr int BG = NB / 8;                    // batch groups of 8
    constexpr int GXP = NB + 1;                   // gx row pitch
    const uint32_t a_bytes = (uint32_t)KCH * 2048, h_bytes = (uint32_t)KCHB * BG * 128;
    uint8_t* a_s = smem;
    uint8_t* h_s = smem + a_bytes;                                    // [2][h_bytes]
    float* gx = reinterpret_cast<float*>(h_s + 2 * h_bytes);          // [128][NB+1] activated gates
    float* cy = gx + 128 * GXP;                                       // [2][2][NB][33]  (c | h) staging
    const uint32_t bar = smem_u32(&bar_mma);

    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 4) tmem_alloc(smem_u32(&tmem_base_s), FW_NACC * NB < 32 ? 32 : FW_NACC * NB);
    if (warp < 4) {   // pack this CTA's slice of W_h:  A[g][k] = Wh[k][gate(g)*H + u0 + g%32]
        const float* Wh = p.Wh[d];
        const int u = u0 + lane, g = warp * 32 + lane;
        for (int kc = 0; kc < KCH; ++kc) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int k = kc * 8 + e;
                v[e] = (u < H && k < H) ? __ldg(Wh + (size_t)k * p.ldw + warp * H + u) : 0.f;
            }
            *reinterpret_cast<uint4*>(a_s + (size_t)(kc * 16 + (g >> 3)) * 128 + (g & 7) * 16) =
                make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        }
    }
    for (uint32_t i = tid * 16; i < 2 * h_bytes; i += BT_THREADS * 16) *reinterpret_cast<uint4*>(h_s + i) = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_arrive_release();
    cluster_wait();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = idesc_bf16(128, NB, 0, 0);

    if (warp < 4) {
        // =========================== compute warps ===========================
        const int q = warp, ug = u0 + lane;
        constexpr int ITEMS = (4 * NB + 127) / 128;   // (k-chunk, mixture) items per thread in the cell phase
        float creg[ITEMS][8];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
#pragma unroll
            for (int e = 0; e < 8; ++e) creg[i][e] = 0.f;
        float zx[NB];
        auto load_zx = [&](int t) {
            const float* gp = p.gates + (((size_t)d * T + t) * B + b0) * H4 + q * H + ug;
#pragma unroll
            for (int b = 0; b < NB; ++b) zx[b] = (b < nvalid && ug < H) ? __ldcg(gp + (size_t)b * H4) : 0.f;
        };
        const float fb = q == 2 ? p.forget_bias : 0.f;
        long long* prof = (blockIdx.x == 0 && tid == 0) ? p.prof : nullptr;
        for (int s = 0; s < T; ++s) {
            const int t = d == 0 ? s : T - 1 - s;
            PROF(0);
            load_zx(t);                           // in flight while the MMA of this step runs
            uint32_t acc[NB];
            if (s > 0) {
                mbar_wait(bar, (s - 1) & 1);
                PROF(1);
                tc_fence_after();
#pragma unroll
                for (int a = 0; a < FW_NACC; ++a) {
                    if (a >= KCH / 2) break;          // fewer k-steps than accumulators (tiny H)
                    uint32_t part[NB];
                    if (NB == 16) tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + a * NB, part);
                    else {
#pragma unroll
                        for (int c0 = 0; c0 < NB; c0 += 32) tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + a * NB + c0, part + c0);
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int b = 0; b < NB; ++b)
                        acc[b] = a == 0 ? part[b] : __float_as_uint(__uint_as_float(acc[b]) + __uint_as_float(part[b]));
                }
            } else {
#pragma unroll
                for (int b = 0; b < NB; ++b) acc[b] = 0u;
            }
            PROF(2);
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const float z = __uint_as_float(acc[b]) + zx[b] + fb;
                gx[(q * 32 + lane) * GXP + b] = q == 1 ? tanh_fast(z) : sigmoid_fast(z);
            }
            PROF(3);
            bar_sync_named(1, 256);               // gx complete (compute + writer warps)
            PROF(4);
            const uint32_t hdst = smem_u32(h_s + ((s + 1) & 1) * h_bytes);
            float* cys = cy + (s & 1) * (2 * NB * 33);
#pragma unroll
            for (int it = 0; it < ITEMS; ++it) {
                const int item = tid + it * 128;
                if (item < 4 * NB) {
                    const int kc = item / NB, b = item % NB;
                    float hv[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int ul = kc * 8 + e;
                        const float gi = gx[(0 * 32 + ul) * GXP + b], gj = gx[(1 * 32 + ul) * GXP + b];
                        const float gf = gx[(2 * 32 + ul) * GXP + b], go = gx[(3 * 32 + ul) * GXP + b];
                        const float c = fmaf(creg[it][e], gf, gi * gj);
                        creg[it][e] = c;
                        hv[e] = b < nvalid ? tanh_fast(c) * go : 0.f;
                        cys[b * 33 + ul] = c;
                        cys[NB * 33 + b * 33 + ul] = hv[e];
                    }
                    if (s + 1 < T) {
                        const uint4 pk = make_uint4(pack_bf16(hv[0], hv[1]), pack_bf16(hv[2], hv[3]), pack_bf16(hv[4], hv[5]),
                                                    pack_bf16(hv[6], hv[7]));
                        const uint32_t off = (uint32_t)(((crank * 4 + kc) * BG + (b >> 3)) * 128 + (b & 7) * 16);
                        for (uint32_t r = 0; r < NC; ++r) st_cluster_v4(mapa(hdst + off, r), pk);
                    }
                }
            }
            PROF(5);
            fence_proxy_async_cluster();
            tc_fence_before();
            PROF(6);
            cluster_arrive_release();
            PROF(7);
            cluster_wait();
            PROF(8);
        }
    } else if (warp == 4) {
        // =========================== MMA issuer ===========================
        long long* prof = (blockIdx.x == 0 && lane == 0) ? p.prof : nullptr;
        for (int s = 0; s < T; ++s) {
            PROF(9);
            if (s > 0) {                              // converged: every lane computes the (uniform) descriptors
                tc_fence_after();
                fence_async_smem();
                const uint32_t aaddr = smem_u32(a_s), haddr = smem_u32(h_s + (s & 1) * h_bytes);
                const bool leader = elect_one();
                for (int kk = 0; kk < KCH / 2; ++kk) {
                    const uint64_t ad = smem_desc(aaddr + kk * 4096, 2048, 128);
                    const uint64_t bd = smem_desc(haddr + kk * 2 * BG * 128, BG * 128, 128);
                    if (leader) mma_bf16(tmem + (kk % FW_NACC) * NB, ad, bd, idesc, kk >= FW_NACC);
                }
                if (leader) mma_commit(bar);
            }
            PROF(10);
            __syncwarp();
            cluster_arrive_relaxed();
            cluster_wait();
        }
    } else {
        // =========================== writer warps: saved gates, c, y -> global ===========================
        const int wt = tid - 160;                     // 0..127
        const int wq = wt >> 5, wu = u0 + (wt & 31);  // gate / unit for the gate stores
        auto flush_cy = [&](int sprev) {
            const int tp = d == 0 ? sprev : T - 1 - sprev;
            const float* cys = cy + (sprev & 1) * (2 * NB * 33);
            const int ul = wt & 31;
            if (u0 + ul < H)
                for (int b = wt >> 5; b < nvalid; b += 4) {
                    __stcg(p.cst + (((size_t)d * T + tp) * B + b0 + b) * H + u0 + ul, cys[b * 33 + ul]);
                    __stcg(p.y + ((size_t)tp * B + b0 + b) * 2 * H + d * H + u0 + ul, cys[NB * 33 + b * 33 + ul]);
                }
        };
        for (int s = 0; s < T; ++s) {
            const int t = d == 0 ? s : T - 1 - s;
            bar_sync_named(1, 256);
            if (wu < H) {
                float* gout = p.gates + (((size_t)d * T + t) * B + b0) * H4 + wq * H + wu;
                for (int b = 0; b < nvalid; ++b) __stcg(gout + (size_t)b * H4, gx[wt * GXP + b]);
            }
            if (s > 0) flush_cy(s - 1);
            cluster_arrive_relaxed();
            cluster_wait();
        }
        flush_cy(T - 1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, FW_NACC * NB < 32 ? 32 : FW_NACC * NB);
}

template <int NB>
int launch_fwd(const RecTcFwd& p, int NC, cudaStream_t st) {
    const int KCH = (p.H + 15) / 16 * 2, KCHB = NC * 4;
    const size_t smem = (size_t)KCH * 2048 + 2 * (size_t)KCHB * (NB / 8) * 128 + (size_t)128 * (NB + 1) * 4 +
                        (size_t)2 * 2 * NB * 33 * 4;
    if (smem > 226 * 1024) { set_error("blstm_rec_fwd_tc: H=%d NB=%d needs %zu B of shared memory", p.H, NB, smem); return AMSS_ERR_UNSUPPORTED; }
    AMSS_CUDA(cudaFuncSetAttribute(blstm_rec_fwd_tc_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    const float* Wh[2];   // [H][ldw]
    int ldw;
    const float* gates;   // [2][T][B][4H] activated gates (saved by the forward pass)
    const float* cst;     // [2][T][B][H]
    const float* dy;      // [T][B][2H]
    float* dZ;            // [2][T][B][4H]
    int B, T, H, nsub, MT;
};

template <int NB>
__global__ void __launch_bounds__(BT_THREADS, 1) blstm_rec_bwd_tc_kernel(RecTcBwd p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t crank = cluster_ctarank(), NC = cluster_nctarank();
    const int cid = blockIdx.x / NC, d = cid / p.nsub, sub = cid % p.nsub;
    const int H = p.H, T = p.T, B = p.B, H4 = 4 * H, MT = p.MT;
    const int b0 = sub * NB, nvalid = min(NB, B - b0);
    const int u0 = crank * 32;
    constexpr int BG = NB / 8;
    constexpr int ZP = NB + 1;
    const uint32_t a_bytes = (uint32_t)MT * 128 * 128 * 2;               // [MT*128 units][128 own cols] bf16
    const uint32_t r_bytes = (uint32_t)NC * 32 * NB * 2;                 // partial sums from every CTA, bf16
    uint8_t* a_s = smem;
    uint8_t* z_s = a_s + a_bytes;                                        // dz operand [NB][128] bf16, K-major
    uint8_t* r_s = z_s + NB * 128 * 2;                                   // [2][NC][32][NB] bf16
    float* dzs = reinterpret_cast<float*>(r_s + 2 * r_bytes);            // [128][NB+1] fp32 dz staging for the writers
    const uint32_t bar = smem_u32(&bar_mma);
    const uint32_t tcols = MT * NB <= 32 ? 32 : (MT * NB <= 64 ? 64 : (MT * NB <= 128 ? 128 : 256));

    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 4) tmem_alloc(smem_u32(&tmem_base_s), tcols);
    {   // pack A[u][g] = Wh[u][gate(g)*H + u0 + g%32]; unit id = (u, k-chunk of 8 own cols)
        const float* Wh = p.Wh[d];
        const int nunits = MT * 128 * 16;
        for (int id = tid; id < nunits; id += BT_THREADS) {
            const int u = id % (MT * 128), kc = id / (MT * 128);
            const int gate = kc >> 2, ul = (kc & 3) * 8;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
                v[e] = (u < H && u0 + ul + e < H) ? __ldg(Wh + (size_t)u * p.ldw + gate * H + u0 + ul + e) : 0.f;
            *reinterpret_cast<uint4*>(a_s + (size_t)(kc * (MT * 16) + (u >> 3)) * 128 + (u & 7) * 16) =
                make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_arrive_release();
    cluster_wait();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = idesc_bf16(128, NB, 0, 0);

    if (warp < 4) {
        // =========================== compute warps ===========================
        const int q = warp;
        constexpr int IT = NB / 4;                    // mixtures per thread: b = q*IT + i, unit = lane
        const int ug = u0 + lane;
        float dcc[IT];
#pragma unroll
        for (int i = 0; i < IT; ++i) dcc[i] = 0.f;
        float sv[IT][7];                              // gi gj gf go c cprev dy  (prefetched one step ahead)
        auto prefetch = [&](int s) {
            const int t = d == 0 ? s : T - 1 - s;
            const int tprev = d == 0 ? t - 1 : t + 1;
#pragma unroll
            for (int i = 0; i < IT; ++i) {
                const int b = q * IT + i;
                const bool ok = b < nvalid && ug < H;
                const size_t row = ((size_t)d * T + t) * B + b0 + b;
                const float* g = p.gates + row * H4 + ug;
                sv[i][0] = ok ? __ldcg(g) : 0.f;
                sv[i][1] = ok ? __ldcg(g + H) : 0.f;
                sv[i][2] = ok ? __ldcg(g + 2 * H) : 0.f;
                sv[i][3] = ok ? __ldcg(g + 3 * H) : 0.f;
                sv[i][4] = ok ? __ldcg(p.cst + row * H + ug) : 0.f;
                sv[i][5] = (ok && s > 0) ? __ldcg(p.cst + (((size_t)d * T + tprev) * B + b0 + b) * H + ug) : 0.f;
                sv[i][6] = ok ? __ldcg(p.dy + ((size_t)t * B + b0 + b) * 2 * H + d * H + ug) : 0.f;
            }
        };
        for (int s = T - 1; s >= 0; --s) {
            const int n = T - 1 - s;                  // step counter
            prefetch(s);                              // in flight while the MMA of this step runs
            const uint32_t rbuf = smem_u32(r_s + (n & 1) * r_bytes);
            if (n > 0) {
                // partial sums P[u, b] = sum_{own cols} Wh[u, g] dz_{next}[b, g]  ->  owner CTA of unit u
                mbar_wait(bar, (n - 1) & 1);
                tc_fence_after();
                for (int m = 0; m < MT; ++m) {
                    const uint32_t dest = m * 4 + q;  // units m*128 + q*32 + lane  ->  CTA dest, local unit = lane
                    uint32_t acc[NB];
                    if (NB == 16) tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + m * NB, acc);
                    else tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + m * NB, acc);
                    tmem_ld_wait();
                    if (dest < NC) {
                        const uint32_t base = mapa(rbuf + (crank * 32 + lane) * (NB * 2), dest);
#pragma unroll
                        for (int b = 0; b < NB; b += 8)
                            st_cluster_v4(base + b * 2,
                                          make_uint4(pack_bf16(__uint_as_float(acc[b]), __uint_as_float(acc[b + 1])),
                                                     pack_bf16(__uint_as_float(acc[b + 2]), __uint_as_float(acc[b + 3])),
                                                     pack_bf16(__uint_as_float(acc[b + 4]), __uint_as_float(acc[b + 5])),
                                                     pack_bf16(__uint_as_float(acc[b + 6]), __uint_as_float(acc[b + 7]))));
                    }
                }
                tc_fence_before();
            }
            cluster_arrive_release();
            cluster_wait();
            // reduce the partials of every CTA for (unit = lane, mixtures q*IT .. +IT)
            float dh[IT];
#pragma unroll
            for (int i = 0; i < IT; ++i) dh[i] = sv[i][6];
            if (n > 0) {
                const uint8_t* rb = r_s + (n & 1) * r_bytes + (size_t)lane * (NB * 2) + q * IT * 2;
                for (uint32_t c = 0; c < NC; ++c) {
                    const uint8_t* rp = rb + (size_t)c * 32 * NB * 2;
                    if (IT == 4) {
                        const uint2 w = *reinterpret_cast<const uint2*>(rp);
                        dh[0] += bf16_lo(w.x); dh[1] += bf16_hi(w.x); dh[2] += bf16_lo(w.y); dh[3] += bf16_hi(w.y);
                    } else {
                        const uint4 w = *reinterpret_cast<const uint4*>(rp);
                        dh[0] += bf16_lo(w.x); dh[1] += bf16_hi(w.x); dh[2] += bf16_lo(w.y); dh[3] += bf16_hi(w.y);
                        dh[4 % IT] += bf16_lo(w.z); dh[5 % IT] += bf16_hi(w.z); dh[6 % IT] += bf16_lo(w.w); dh[7 % IT] += bf16_hi(w.w);
                    }
                }
            }
            // gate derivatives; dz -> fp32 staging (writers) and bf16 operand of the next step's MMA
#pragma unroll
            for (int i = 0; i < IT; ++i) {
                const int b = q * IT + i;
                const float gi = sv[i][0], gj = sv[i][1], gf = sv[i][2], go = sv[i][3], c = sv[i][4], cprev = sv[i][5];
                const float tc_ = tanh_fast(c);
                const float dc = dcc[i] + dh[i] * go * (1.f - tc_ * tc_);
                float dz[4];
                dz[0] = dc * gj * gi * (1.f - gi);
                dz[1] = dc * gi * (1.f - gj * gj);
                dz[2] = dc * cprev * gf * (1.f - gf);
                dz[3] = dh[i] * tc_ * go * (1.f - go);
                dcc[i] = dc * gf;
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                    dzs[(g4 * 32 + lane) * ZP + b] = dz[g4];
                    const int k = g4 * 32 + lane;
                    *reinterpret_cast<__nv_bfloat16*>(z_s + (size_t)((k >> 3) * BG + (b >> 3)) * 128 + (b & 7) * 16 + (k & 7) * 2) =
                        __float2bfloat16_rn(dz[g4]);
                }
            }
            fence_async_smem();
            bar_sync_named(1, 288);                   // dz staged: MMA warp may issue, writers may store
        }
    } else if (warp == 4) {
        // =========================== MMA issuer ===========================
        for (int s = T - 1; s >= 0; --s) {
            const int n = T - 1 - s;
            if (n > 0) {                              // converged issue loop, elected lane executes the MMAs
                tc_fence_after();
                const uint32_t aaddr = smem_u32(a_s), zaddr = smem_u32(z_s);
                const bool leader = elect_one();
                for (int kk = 0; kk < 8; ++kk)
                    for (int m = 0; m < MT; ++m) {
                        const uint64_t ad = smem_desc(aaddr + m * 2048 + kk * 2 * (MT * 16) * 128, (MT * 16) * 128, 128);
                        const uint64_t bd = smem_desc(zaddr + kk * 2 * BG * 128, BG * 128, 128);
                        if (leader) mma_bf16(tmem + m * NB, ad, bd, idesc, kk > 0);
                    }
                if (leader) mma_commit(bar);
            }
            __syncwarp();
            cluster_arrive_relaxed();
            cluster_wait();
            bar_sync_named(1, 288);
        }
    } else {
        // =========================== writer warps: dZ -> global ===========================
        const int wt = tid - 160, wq = wt >> 5, wu = u0 + (wt & 31);
        for (int s = T - 1; s >= 0; --s) {
            const int t = d == 0 ? s : T - 1 - s;
            cluster_arrive_relaxed();
            cluster_wait();
            bar_sync_named(1, 288);
            if (wu < H) {
                float* zo = p.dZ + (((size_t)d * T + t) * B + b0) * H4 + wq * H + wu;
                for (int b = 0; b < nvalid; ++b) __stcg(zo + (size_t)b * H4, dzs[wt * ZP + b]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, tcols);
}

template <int NB>
int launch_bwd(const RecTcBwd& p, int NC, cudaStream_t st) {
    const size_t smem = (size_t)p.MT * 128 * 128 * 2 + (size_t)NB * 128 * 2 + 2 * (size_t)NC * 32 * NB * 2 +
                        (size_t)128 * (NB + 1) * 4;
    if (smem > 226 * 1024) { set_error("blstm_rec_bwd_tc: H=%d NB=%d needs %zu B of shared memory", p.H, NB, smem); return AMSS_ERR_UNSUPPORTED; }
    AMSS_CUDA(cudaFuncSetAttribute(blstm_rec_bwd_tc_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (NC > 8) AMSS_CUDA(cudaFuncSetAttribute(blstm_rec_bwd_tc_kernel<NB>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(NC * 2 * p.nsub);
    cfg.blockDim = dim3(BT_THREADS);
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

// as many clusters as the chip holds (one CTA per SM); sub-batches of 16 / 32 / 64 mixtures
int pick_nb(int B, int NC, int nb_max) {
    const int max_clusters = std::max(2, kNumSMs / NC);
    for (int nb : {16, 32, 64}) {
        if (nb > nb_max) break;
        const int nsub = (B + nb - 1) / nb;
        if (2 * nsub <= max_clusters) return nb;
    }
    return nb_max;
}

}  // namespace

void blstm_tc_set_profile(long long* dev_buf) { g_prof = dev_buf; }

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
    p.Wh[0] = Wh_fw; p.Wh[1] = Wh_bw; p.ldw = ldw; p.gates = gates; p.cst = cst; p.y = y;
    p.B = B; p.T = T; p.H = H; p.forget_bias = forget_bias;
    const int nb = pick_nb(B, NC, 64);
    p.nsub = (B + nb - 1) / nb;
    if (nb == 16) return launch_fwd<16>(p, NC, st);
    if (nb == 32) return launch_fwd<32>(p, NC, st);
    return launch_fwd<64>(p, NC, st);
}

int blstm_rec_bwd_tc(const float* Wh_fw, const float* Wh_bw, int ldw, const float* gates, const float* cst,
                     const float* dy, float* dZ, int B, int T, int H, cudaStream_t st) {
    const int NC = (H + 31) / 32;
    RecTcBwd p;
    p.Wh[0] = Wh_fw; p.Wh[1] = Wh_bw; p.ldw = ldw; p.gates = gates; p.cst = cst; p.dy = dy; p.dZ = dZ;
    p.B = B; p.T = T; p.H = H; p.MT = (NC * 32 + 127) / 128;
    const int nb = pick_nb(B, NC, 32);
    p.nsub = (B + nb - 1) / nb;
    if (nb == 16) return launch_bwd<16>(p, NC, st);
    return launch_bwd<32>(p, NC, st);
}

}  // namespace amss
