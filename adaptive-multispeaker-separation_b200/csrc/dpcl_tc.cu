// Fused backward of the DPCL affinity loss (models/dpcl.py:41-86) and of the tf.nn.l2_normalize that
// produced the embeddings (utils/ops.py:323-324), tensor-core version (AMSS_PREC_BF16):
//     dV_i = g * D_i * (An v_i - Bn[:, l_i]),   dz_i = inv_i * (dV_i - v_i <v_i, dV_i>)
// The [points x E] x [E x E] product runs on tcgen05 (M = 128 points per tile, N = K = E padded to 16):
// loader warps stream 128-point tiles of V (coalesced float4), keep the fp32 rows in shared memory for
// the normalisation Jacobian and write the bf16 K-major A operand; the E x E matrix An of the current
// mixture is the resident B operand; accumulators are double buffered in TMEM; the epilogue thread of a
// point owns its whole row (TMEM lane), so the row dot product <v, dV> needs no shuffles.  One pass over
// V, one write of dz: the kernel is HBM-bound (2 x 4E bytes per point).
#include "common.cuh"
#include "tc.cuh"
#include <algorithm>

namespace amss {
namespace {

using namespace tc;

constexpr int DT_THREADS = 288;   // warps 0-3 loaders, 4 MMA (+TMEM alloc), 5-8 epilogue
constexpr int DT_MAXS = 4;

struct DtParams {
    const float* V;            // [B][TF][E] normalised embeddings
    const uint8_t* labels;     // [B][TF]
    const float* dloss;        // [1]
    const float* stats;        // [B][E*E + S*E + S + 1]  (An, Bn, dinv, loss) from the forward pass
    const float* inv_norm;     // [B*TF]
    float* dz;                 // [B][TF][E] fp32, or
    uint16_t* dzb;             // [B][TF][E] bf16 (E % 8 == 0) when non-null: the head GEMMs consume it directly
    int B, E, S, EK, pitch;
    int64_t TF, ntiles, per;   // tiles of 128 points per mixture; flat tiles per CTA
};

__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// RV: float4 per loader thread and tile held in registers (E/4 <= RV).  E4C: E/4 as a compile-time constant (0 = runtime):
// the unit -> (row, column) splits divide by it once per 16 bytes, and with 64-bit tile indices divided per tile the
// kernel was spending half of its issue slots on integer arithmetic (ncu: IMAD + ISETP + IADD3 = 40 % of instructions).
template <int RV, int E4C>
__global__ void __launch_bounds__(DT_THREADS, 2) dpcl_bwd_tc_kernel(DtParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[8];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int E = p.E, EK = p.EK, S = p.S, pitch = p.pitch;
    const uint32_t a_bytes = (uint32_t)(EK / 8) * 16 * 128;           // A operand: 128 rows x EK
    const uint32_t b_bytes = (uint32_t)(EK / 8) * (EK / 8) * 128;     // B operand: EK x EK
    float* vs = reinterpret_cast<float*>(smem);                       // [2][128][pitch] fp32 rows (v, then dz)
    uint8_t* a_s = smem + (size_t)2 * 128 * pitch * 4;                // [2][a_bytes]
    uint8_t* b_s = a_s + 2 * a_bytes;                                 // [2][b_bytes]   (by mixture generation)
    float* bn_s = reinterpret_cast<float*>(b_s + 2 * b_bytes);        // [2][S*E + S]   Bn, dinv
    const uint32_t a_full = smem_u32(&bars[0]), a_empty = a_full + 16, t_full = a_full + 32, t_empty = a_full + 48;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(a_full + 8 * b, 128); mbar_init(a_empty + 8 * b, 1);
            mbar_init(t_full + 8 * b, 1);   mbar_init(t_empty + 8 * b, 128);
        }
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(smem_u32(&tmem_base_s), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int64_t w0 = (int64_t)blockIdx.x * p.per, w1 = min(w0 + p.per, (int64_t)p.B * p.ntiles);
    const int sstride = E * E + S * E + S + 1;
    const int b_first = w0 < w1 ? (int)(w0 / p.ntiles) : 0;
    const int nt = (int)p.ntiles, t_first = w0 < w1 ? (int)(w0 - (int64_t)b_first * p.ntiles) : 0;   // (mixture, tile) walk
    const int ntile = w0 < w1 ? (int)(w1 - w0) : 0;                                                   // without divisions

    if (warp < 4) {
        // ---------------- loaders: V tile -> fp32 rows + bf16 A operand; An/Bn when the mixture changes ----------------
        int bcur = -1;
        const int e4 = E4C ? E4C : E / 4;
        // software pipeline: tile i+1 is fetched into registers (coalesced float4, unit u = tid + 128 j) before tile i is
        // handed to the MMA, so the HBM latency overlaps the previous tile's conversion / MMAs / epilogue
        float4 tilev[RV];
        auto fetch = [&](int b, int tix) {
            const int64_t p0 = (int64_t)tix * 128;
            const int np = (int)min((int64_t)128, p.TF - p0);
            const float4* src = reinterpret_cast<const float4*>(p.V + ((size_t)b * p.TF + p0) * E);
#pragma unroll
            for (int j = 0; j < RV; ++j) {
                const int u = tid + 128 * j;
                tilev[j] = (j < e4 && u < np * e4) ? __ldcs(src + u) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        if (ntile > 0) fetch(b_first, t_first);
        int b = b_first, tix = t_first;
        for (uint32_t i = 0; i < (uint32_t)ntile; ++i) {
            const uint32_t buf = i & 1, ph = (i >> 1) & 1;
            int bnext = b, tnext = tix + 1;
            if (tnext == nt) { tnext = 0; ++bnext; }
            mbar_wait(t_empty + 8 * buf, ph ^ 1);          // the epilogue of tile i-2 has finished with vs[buf] / TMEM[buf]
            mbar_wait(a_empty + 8 * buf, ph ^ 1);          // the MMAs of tile i-2 have finished with a_s[buf]
            if (b != bcur) {        // generation (b - b_first) & 1: An -> bf16 B operand [n][k] (An symmetric), Bn/dinv fp32
                const int gen = (b - b_first) & 1;
                const float* st = p.stats + (size_t)b * sstride;
                uint8_t* bd = b_s + gen * b_bytes;
                for (int u = tid; u < EK * (EK / 8); u += 128) {
                    const int n = u % EK, kc = u / EK;
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) { const int k = kc * 8 + e; v[e] = (n < E && k < E) ? st[n * E + k] : 0.f; }
                    *reinterpret_cast<uint4*>(bd + (size_t)(kc * (EK / 8) + (n >> 3)) * 128 + (n & 7) * 16) =
                        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                }
                float* bn = bn_s + gen * (S * E + S);
                for (int u = tid; u < S * E + S; u += 128) bn[u] = st[E * E + u];
                bcur = b;
            }
            float* vb = vs + (size_t)buf * 128 * pitch;
#pragma unroll
            for (int j = 0; j < RV; ++j) {
                if (j < e4) {
                    const int u = tid + 128 * j, r = u / e4, c = u - r * e4;
                    *reinterpret_cast<float4*>(vb + r * pitch + c * 4) = tilev[j];
                }
            }
            if (i + 1 < (uint32_t)ntile) fetch(bnext, tnext);
            named_sync(1, 128);
            uint8_t* ab = a_s + buf * a_bytes;
            const float* row = vb + tid * pitch;
            for (int kc = 0; kc < EK / 8; ++kc) {
                float v[8];
                if (kc * 8 + 8 <= E) {
                    const float4 x0 = *reinterpret_cast<const float4*>(row + kc * 8), x1 = *reinterpret_cast<const float4*>(row + kc * 8 + 4);
                    v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = kc * 8 + e < E ? row[kc * 8 + e] : 0.f;
                }
                *reinterpret_cast<uint4*>(ab + (size_t)(kc * 16 + (tid >> 3)) * 128 + (tid & 7) * 16) =
                    make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            }
            fence_async_smem();
            mbar_arrive(a_full + 8 * buf);
            b = bnext; tix = tnext;
        }
    } else if (warp == 4) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = idesc_bf16(128, EK, 0, 0);
        const bool leader = elect_one();
        int b = b_first, tix = t_first;
        for (uint32_t i = 0; i < (uint32_t)ntile; ++i) {
            const uint32_t buf = i & 1, ph = (i >> 1) & 1;
            mbar_wait(a_full + 8 * buf, ph);
            tc_fence_after();
            const uint32_t aaddr = smem_u32(a_s + buf * a_bytes), baddr = smem_u32(b_s + ((b - b_first) & 1) * b_bytes);
            for (int kk = 0; kk < EK / 16; ++kk) {
                const uint64_t ad = smem_desc(aaddr + kk * 2 * 16 * 128, 16 * 128, 128);
                const uint64_t bd = smem_desc(baddr + kk * 2 * (EK / 8) * 128, (EK / 8) * 128, 128);
                if (leader) mma_bf16(tmem + buf * 64, ad, bd, idesc, kk > 0);
            }
            if (leader) { mma_commit(a_empty + 8 * buf); mma_commit(t_full + 8 * buf); }
            if (++tix == nt) { tix = 0; ++b; }
        }
    } else {
        // ---------------- epilogue: thread = point (TMEM lane); dV row, <v,dV>, dz row ----------------
        const int q = warp & 3, r = q * 32 + lane, et = tid - 160;
        const float gscale = p.dloss[0] / (float)p.B;
        const int e4 = E4C ? E4C : E / 4;
        int b = b_first, tix = t_first;
        for (uint32_t i = 0; i < (uint32_t)ntile; ++i) {
            const int64_t p0 = (int64_t)tix * 128;
            const int np = (int)min((int64_t)128, p.TF - p0);
            const uint32_t buf = i & 1, ph = (i >> 1) & 1;
            const float* bn = bn_s + ((b - b_first) & 1) * (S * E + S);
            int l = 0;
            float inv = 0.f;
            if (r < np) { l = p.labels[(size_t)b * p.TF + p0 + r]; inv = p.inv_norm[(size_t)b * p.TF + p0 + r]; }
            mbar_wait(a_full + 8 * buf, ph);               // acquire the loaders' fp32 rows / Bn
            mbar_wait(t_full + 8 * buf, ph);
            tc_fence_after();
            float* vb = vs + (size_t)buf * 128 * pitch;
            float4* row4 = reinterpret_cast<float4*>(vb + r * pitch);
            const float d = gscale * bn[S * E + l];
            const float* bl = bn + l * E;
            float dv[64];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c * 16 < EK) {                         // warp-uniform
                    uint32_t v[16];
                    tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + buf * 64 + c * 16, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int e = c * 16 + j;
                        dv[e] = e < E ? d * (__uint_as_float(v[j]) - bl[e]) : 0.f;
                    }
                }
            }
            tc_fence_before();
            float dot = 0.f;
#pragma unroll
            for (int c = 0; c < 16; ++c)
                if (c < e4) {
                    const float4 x = row4[c];
                    dot = fmaf(x.x, dv[4 * c], dot); dot = fmaf(x.y, dv[4 * c + 1], dot);
                    dot = fmaf(x.z, dv[4 * c + 2], dot); dot = fmaf(x.w, dv[4 * c + 3], dot);
                }
            if (inv < 0.f) { inv = -inv; dot = 0.f; }      // clamped branch of l2_normalize: linear map
            if (p.dzb && (E & 7) == 0) {
                // bf16 dz straight from the registers of the thread that owns the point: E / 8 16-byte stores per row (the rows of
                // a warp are contiguous in memory, so the warp's stores fill whole lines between them); no staging pass, no
                // block barrier, no second trip through shared memory
                if (r < np) {
                    uint4* dst = reinterpret_cast<uint4*>(p.dzb + ((size_t)b * p.TF + p0 + r) * E);
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (2 * c < e4) {
                            const float4 x0 = row4[2 * c], x1 = row4[2 * c + 1];
                            __stcs(dst + c, make_uint4(pack_bf16(inv * (dv[8 * c] - x0.x * dot), inv * (dv[8 * c + 1] - x0.y * dot)),
                                                       pack_bf16(inv * (dv[8 * c + 2] - x0.z * dot), inv * (dv[8 * c + 3] - x0.w * dot)),
                                                       pack_bf16(inv * (dv[8 * c + 4] - x1.x * dot), inv * (dv[8 * c + 5] - x1.y * dot)),
                                                       pack_bf16(inv * (dv[8 * c + 6] - x1.z * dot), inv * (dv[8 * c + 7] - x1.w * dot))));
                        }
                }
            } else {
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    if (c < e4) {
                        const float4 x = row4[c];
                        row4[c] = make_float4(inv * (dv[4 * c] - x.x * dot), inv * (dv[4 * c + 1] - x.y * dot),
                                              inv * (dv[4 * c + 2] - x.z * dot), inv * (dv[4 * c + 3] - x.w * dot));
                    }
                named_sync(2, 128);                        // all dz rows of the tile are in vs[buf]
                if (p.dzb) {
                    const int e8 = E / 8;
                    uint4* dst = reinterpret_cast<uint4*>(p.dzb + ((size_t)b * p.TF + p0) * E);
                    for (int u = et; u < np * e8; u += 128) {
                        const int rr = u / e8, c = u - rr * e8;
                        const float4 f0 = *reinterpret_cast<const float4*>(vb + rr * pitch + c * 8);
                        const float4 f1 = *reinterpret_cast<const float4*>(vb + rr * pitch + c * 8 + 4);
                        __stcs(dst + u, make_uint4(pack_bf16(f0.x, f0.y), pack_bf16(f0.z, f0.w), pack_bf16(f1.x, f1.y), pack_bf16(f1.z, f1.w)));
                    }
                } else {
                    float4* dst = reinterpret_cast<float4*>(p.dz + ((size_t)b * p.TF + p0) * E);
                    for (int u = et; u < np * e4; u += 128) {
                        const int rr = u / e4, c = u - rr * e4;
                        __stcs(dst + u, *reinterpret_cast<const float4*>(vb + rr * pitch + c * 4));
                    }
                }
            }
            mbar_arrive(t_empty + 8 * buf);                // vs[buf] and TMEM[buf] may be reused
            if (++tix == nt) { tix = 0; ++b; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, 128);
}

}  // namespace

bool dpcl_bwd_tc_supported(int E, int S) { return E % 4 == 0 && E >= 8 && E <= 64 && S >= 1 && S <= DT_MAXS; }

int dpcl_bwd_tc(const float* V, const uint8_t* labels, const float* dloss, const float* stats, const float* inv_norm, int B,
                int64_t TF, int E, int S, float* dz, uint16_t* dz_bf16, cudaStream_t st) {
    DtParams p;
    p.V = V; p.labels = labels; p.dloss = dloss; p.stats = stats; p.inv_norm = inv_norm; p.dz = dz; p.dzb = dz_bf16;
    p.B = B; p.E = E; p.S = S; p.TF = TF;
    p.EK = (E + 15) / 16 * 16;
    p.pitch = 4 * (((E + 3) / 4) | 1);
    p.ntiles = (TF + 127) / 128;
    const int64_t total = (int64_t)B * p.ntiles;
    const int grid = (int)std::min<int64_t>(total, 2 * kNumSMs);
    p.per = (total + grid - 1) / grid;
    const size_t smem = (size_t)2 * 128 * p.pitch * 4 + 2 * (size_t)(p.EK / 8) * 16 * 128 + 2 * (size_t)(p.EK / 8) * (p.EK / 8) * 128 +
                        2 * (size_t)(S * E + S) * 4 + 64;
    if (E == 40) {          // the reference's embedding size (utils/trainer.py:74): E/4 folded into the index arithmetic
        AMSS_CUDA(cudaFuncSetAttribute((dpcl_bwd_tc_kernel<10, 10>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        AMSS_LAUNCH((dpcl_bwd_tc_kernel<10, 10>), grid, DT_THREADS, smem, st, p);
    } else if (E <= 40) {
        AMSS_CUDA(cudaFuncSetAttribute((dpcl_bwd_tc_kernel<10, 0>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        AMSS_LAUNCH((dpcl_bwd_tc_kernel<10, 0>), grid, DT_THREADS, smem, st, p);
    } else {
        AMSS_CUDA(cudaFuncSetAttribute((dpcl_bwd_tc_kernel<16, 0>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        AMSS_LAUNCH((dpcl_bwd_tc_kernel<16, 0>), grid, DT_THREADS, smem, st, p);
    }
    return AMSS_OK;
}

}  // namespace amss

// =================================================================================================
// DPCL forward statistics on the tensor cores (AMSS_PREC_BF16):
//   A_w = sum_p w_p v_p v_p^T  (E x E),   m_s = sum_{p in s} v_p  (E),   w_p = N_{l_p}^{-1/2}
// as ONE accumulating product  D = X^T [X | Y],  X[p][e] = sqrt(w_p) v[p][e],  Y[p][s] = 1[l_p = s]  (exact in bf16;
// the per-speaker factor 1/sqrt(w_s) is applied in fp32 by the epilogue, so it carries no systematic rounding):
// the operand tile (128 points, bf16, MN-major core matrices: a 16-byte unit = 8 consecutive e of one point)
// serves as BOTH the A operand (M = e, padded to 128 rows that stay zero) and the B operand (N = E + S padded
// to 16; the label columns live in the padding of the last unit).  A CTA owns one (mixture, chunk) and keeps
// the accumulator in TMEM across all of its tiles; the partials go to the same buffer the SIMT kernel fills.
// =================================================================================================
namespace amss {
namespace {

struct GtcParams {
    const float* V;            // [B][TF][E]
    const uint8_t* labels;     // [B][TF]
    const float* counts;       // [B][4]
    float* part;               // [B][chunks][E*E + S*E]
    int B, E, S, NN, chunks;
    int64_t TF, ntiles;
};

constexpr int GG_THREADS = 160;           // warps 0-3 loaders / final epilogue, warp 4 MMA (+TMEM alloc)
constexpr uint32_t GG_TILE = 16 * 2048;   // 16 point groups x (16 e-groups x 128 B)

// RV: float4 per embedding row held in registers (E <= 4*RV).  E4C > 0 (E == 4*E4C, E % 8 == 0): the tile is fetched as
// 128*E4C CONSECUTIVE float4 (unit u = tid + 128 j -> point u / E4C), i.e. fully coalesced, instead of one 4E-byte row
// per thread (32 half-used sectors per load instruction).
template <int RV, int E4C>
__global__ void __launch_bounds__(GG_THREADS, 3) dpcl_gram_tc_kernel(GtcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[5];          // full[2], empty[2], done
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int E = p.E, S = p.S, NN = p.NN;
    const int b = blockIdx.x / p.chunks, chunk = blockIdx.x % p.chunks;
    const uint32_t full = smem_u32(&bars[0]), empty = full + 16, done = full + 32;
    if (tid == 0) {
        mbar_init(full, 128); mbar_init(full + 8, 128); mbar_init(empty, 1); mbar_init(empty + 8, 1); mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(smem_u32(&tmem_base_s), 64);
    for (uint32_t i = tid * 16; i < 2 * GG_TILE; i += GG_THREADS * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int64_t t0 = p.ntiles * chunk / p.chunks, t1 = p.ntiles * (chunk + 1) / p.chunks;
    const int egn = (E + S + 7) / 8;                   // 16-byte units per point that carry data

    if (warp < 4) {
        float wS[4], iwS[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const float n = p.counts[b * 4 + s];
            wS[s] = n > 0.f ? rsqrtf(sqrtf(n)) : 0.f;       // sqrt(w) = N^{-1/4}
            iwS[s] = n > 0.f ? sqrtf(sqrtf(n)) : 0.f;       // 1 / sqrt(w)
        }
      if constexpr (E4C > 0) {
        // coalesced variant: registers hold E4C float4 of the tile + the labels of the points they belong to
        float4 rowv[E4C];
        int labv[E4C], lown = 255;
        auto fetch = [&](int64_t tile) {
            const int64_t p0 = tile * 128;
            const float4* src = reinterpret_cast<const float4*>(p.V + ((size_t)b * p.TF + p0) * E);
            const uint8_t* lb = p.labels + (size_t)b * p.TF + p0;
#pragma unroll
            for (int j = 0; j < E4C; ++j) {
                const int u = tid + 128 * j, pt = u / E4C;
                const bool ok = p0 + pt < p.TF;
                rowv[j] = ok ? __ldcs(src + u) : make_float4(0.f, 0.f, 0.f, 0.f);
                labv[j] = ok ? (int)__ldg(lb + pt) : 255;
            }
            lown = p0 + tid < p.TF ? (int)__ldg(lb + tid) : 255;
        };
        if (t0 < t1) fetch(t0);
        uint32_t i = 0;
        for (int64_t tile = t0; tile < t1; ++tile, ++i) {
            const uint32_t buf = i & 1, ph = (i >> 1) & 1;
            uint2 w[E4C];
#pragma unroll
            for (int j = 0; j < E4C; ++j) {
                const int l = labv[j];
                const float sw = l == 0 ? wS[0] : (l == 1 ? wS[1] : (l == 2 ? wS[2] : (l == 3 ? wS[3] : 0.f)));
                w[j] = make_uint2(pack_bf16(rowv[j].x * sw, rowv[j].y * sw), pack_bf16(rowv[j].z * sw, rowv[j].w * sw));
            }
            const int l0 = lown;
            if (tile + 1 < t1) fetch(tile + 1);
            mbar_wait(empty + 8 * buf, ph ^ 1);
            uint8_t* tbase = smem + buf * GG_TILE;
#pragma unroll
            for (int j = 0; j < E4C; ++j) {
                const int u = tid + 128 * j, pt = u / E4C, c = u - pt * E4C;
                *reinterpret_cast<uint2*>(tbase + (size_t)(pt >> 3) * 2048 + (c >> 1) * 128 + (pt & 7) * 16 + (c & 1) * 8) = w[j];
            }
            {   // the one-hot label columns e = E .. E+S-1 (exact in bf16): unit E/8 of this thread's point
                const uint32_t one = 0x3F80u;
                uint32_t q0 = (l0 == 0 && S > 0 ? one : 0u) | (l0 == 1 && S > 1 ? one << 16 : 0u);
                uint32_t q1 = (l0 == 2 && S > 2 ? one : 0u) | (l0 == 3 && S > 3 ? one << 16 : 0u);
                *reinterpret_cast<uint4*>(tbase + (size_t)(tid >> 3) * 2048 + (E4C / 2) * 128 + (tid & 7) * 16) = make_uint4(q0, q1, 0u, 0u);
            }
            fence_async_smem();
            mbar_arrive(full + 8 * buf);
        }
      } else {
        // software pipeline: the row (and label) of tile i+1 is loaded into registers before tile i is converted,
        // so the HBM latency overlaps the shared-memory stores / MMAs of the previous tile
        float4 rowv[RV];
        int lnext = 0;
        bool oknext = false;
        auto fetch = [&](int64_t tile) {
            const int64_t pt = tile * 128 + tid;
            oknext = pt < p.TF;
            lnext = oknext ? p.labels[(size_t)b * p.TF + pt] : 0;
            const float4* src = reinterpret_cast<const float4*>(p.V + ((size_t)b * p.TF + pt) * E);
#pragma unroll
            for (int c = 0; c < RV; ++c) rowv[c] = (oknext && c * 4 + 4 <= E) ? __ldcs(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        if (t0 < t1) fetch(t0);
        uint32_t i = 0;
        for (int64_t tile = t0; tile < t1; ++tile, ++i) {
            const uint32_t buf = i & 1, ph = (i >> 1) & 1;
            const bool ok = oknext;
            const int l = lnext;
            const float sw = ok ? (l == 0 ? wS[0] : (l == 1 ? wS[1] : (l == 2 ? wS[2] : wS[3]))) : 0.f;
            float4 cur[RV];
#pragma unroll
            for (int c = 0; c < RV; ++c) cur[c] = rowv[c];
            if (tile + 1 < t1) fetch(tile + 1);
            mbar_wait(empty + 8 * buf, ph ^ 1);
            uint8_t* tb = smem + buf * GG_TILE + (size_t)(tid >> 3) * 2048 + (tid & 7) * 16;
#pragma unroll
            for (int eg = 0; eg < (RV + 2) / 2; ++eg) {
                if (eg < egn) {
                    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 x0 = 2 * eg < RV ? cur[2 * eg < RV ? 2 * eg : 0] : z4, x1 = 2 * eg + 1 < RV ? cur[2 * eg + 1 < RV ? 2 * eg + 1 : 0] : z4;
                    float v[8] = {x0.x * sw, x0.y * sw, x0.z * sw, x0.w * sw, x1.x * sw, x1.y * sw, x1.z * sw, x1.w * sw};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int e = eg * 8 + j;
                        if (e >= E && e < E + S) v[j] = (ok && l == e - E) ? 1.f : 0.f;    // the one-hot label columns
                    }
                    *reinterpret_cast<uint4*>(tb + eg * 128) =
                        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                }
            }
            fence_async_smem();
            mbar_arrive(full + 8 * buf);
        }
      }
        // ---- final epilogue: D[e][e'] (lanes = e) -> partial Gram + per-speaker sums ----
        if (t1 > t0) {
            mbar_wait(done, 0);
            tc_fence_after();
        }
        float* dst = p.part + ((size_t)b * p.chunks + chunk) * ((size_t)E * E + (size_t)S * E);
        const int e = warp * 32 + lane;
        for (int c0 = 0; c0 < NN; c0 += 16) {
            uint32_t v[16];
            if (t1 > t0) { tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v); tmem_ld_wait(); }
            else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0u;
            }
            if (e < E) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int c = c0 + j;
                    if (c < E) dst[e * E + c] = __uint_as_float(v[j]);
                    else if (c < E + S) dst[E * E + (c - E) * E + e] = __uint_as_float(v[j]) * iwS[(c - E) & 3];
                }
            }
        }
    } else {
        const uint32_t idesc = idesc_bf16(128, NN, 1, 1);      // both operands MN-major
        const bool leader = elect_one();
        uint32_t i = 0;
        for (int64_t tile = t0; tile < t1; ++tile, ++i) {
            const uint32_t buf = i & 1, ph = (i >> 1) & 1;
            mbar_wait(full + 8 * buf, ph);
            tc_fence_after();
            const uint32_t ta = smem_u32(smem + buf * GG_TILE);
            for (int kk = 0; kk < 8; ++kk) {                   // 128 points = 8 K steps of 16
                const uint64_t d = smem_desc(ta + kk * 2 * 2048, 2048, 128);
                if (leader) mma_bf16(tmem, d, d, idesc, (i | kk) != 0);
            }
            if (leader) mma_commit(empty + 8 * buf);
        }
        if (leader && t1 > t0) mma_commit(done);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, 64);
}

}  // namespace

bool dpcl_gram_tc_supported(int E, int S) { return E % 4 == 0 && E >= 8 && E + S <= 64 && S >= 1 && S <= 4; }
int dpcl_gram_tc_chunks(int B) { return std::max(1, (3 * kNumSMs) / std::max(1, B)); }

int dpcl_gram_tc(const float* V, const uint8_t* labels, const float* counts, int B, int64_t TF, int E, int S, int chunks,
                 float* part, cudaStream_t st) {
    GtcParams p;
    p.V = V; p.labels = labels; p.counts = counts; p.part = part; p.B = B; p.E = E; p.S = S; p.TF = TF;
    p.NN = (E + S + 15) / 16 * 16;
    p.chunks = chunks;
    p.ntiles = (TF + 127) / 128;
    const size_t smem = 2 * (size_t)GG_TILE;
    if (E == 40 && (reinterpret_cast<uintptr_t>(V) & 15) == 0) {     // the reference's embedding size: coalesced tile fetch
        AMSS_CUDA(cudaFuncSetAttribute((dpcl_gram_tc_kernel<10, 10>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        AMSS_LAUNCH((dpcl_gram_tc_kernel<10, 10>), B * chunks, GG_THREADS, smem, st, p);
    } else if (E <= 40) {
        AMSS_CUDA(cudaFuncSetAttribute((dpcl_gram_tc_kernel<10, 0>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        AMSS_LAUNCH((dpcl_gram_tc_kernel<10, 0>), B * chunks, GG_THREADS, smem, st, p);
    } else {
        AMSS_CUDA(cudaFuncSetAttribute((dpcl_gram_tc_kernel<16, 0>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        AMSS_LAUNCH((dpcl_gram_tc_kernel<16, 0>), B * chunks, GG_THREADS, smem, st, p);
    }
    return AMSS_OK;
}

}  // namespace amss
