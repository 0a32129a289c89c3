// Dense GEMM on the 5th-gen tensor cores: C[M,N] (+)= op(A) op(B) (+ bias), fp32 in / fp32 out,
// bf16 operands, fp32 accumulation in TMEM.  Used (AMSS_PREC_BF16) for the hoisted BLSTM input
// projections, the embedding head (utils/ops.py:501-503) and every backward GEMM of those.
//
//   1. both operands are converted once to bf16 (same row-major shape, rows padded to 8 elements);
//   2. persistent CTAs (one per SM) walk (tile, k-split) work items.  Tile 128 (M, TMEM lanes) x 256
//      (N, TMEM columns) x 64 (K per stage); 4 loader warps stream 16-byte units with cp.async
//      straight into the canonical no-swizzle core-matrix layout -- K-major when the matrix is
//      K-contiguous in memory, MN-major when it is M/N-contiguous, so transposed operands
//      (dW = X^T dZ, dX = dZ W^T) need no transpose pass -- through a 4-stage mbarrier ring
//      (cp.async.mbarrier.arrive.noinc), i.e. ~190 KB of loads in flight per SM;
//   3. one warp issues tcgen05.mma (converged loop, elected lane);
//   4. TMEM accumulators are double buffered (2 x 256 columns): 4 epilogue warps drain tile i
//      (bias, row remap, accumulate / split-K red.global.add) while tile i+1 is being multiplied.
#include "common.cuh"
#include "tc.cuh"
#include <algorithm>

namespace amss {
namespace {

using namespace tc;

constexpr int GT_BM = 128, GT_BN = 256, GT_BK = 64, GT_STAGES = 4;
constexpr int GT_LOADERS = 128, GT_THREADS = 288;    // warps 0-3 loaders, 4 MMA (+TMEM alloc), 5-8 epilogue
// Padded core-matrix strides: consecutive core matrices along the GLOBAL-contiguous direction are shifted by
// 16 bytes, so a warp's cp.async covers whole 128-byte global lines AND lands in distinct shared-memory banks.
//   K-major tile  (R rows): core(rgrp, kc)  at kc*(R*16+16)   + rgrp*128      (LBO = R*16+16, SBO = 128)
//   MN-major tile (R rows): core(mc, kgrp)  at kgrp*(R/8)*144 + mc*144        (LBO = R*18,    SBO = 144)
__host__ __device__ constexpr int lbo_of(int R, bool kcontig) { return kcontig ? R * 16 + 16 : R * 18; }
__host__ __device__ constexpr int sbo_of(bool kcontig) { return kcontig ? 128 : 144; }
constexpr int GT_A_BYTES = GT_BM * 144, GT_B_BYTES = GT_BN * 144;      // the larger (MN-major) footprint: 8 * LBO
constexpr int GT_STAGE_BYTES = GT_A_BYTES + GT_B_BYTES;

struct GtParams {
    const __nv_bfloat16 *A, *B;       // bf16 copies: A [M][lda] or [K][lda] (ta), B [K][ldb] or [N][ldb] (tb)
    const float* bias;
    float* C;
    int lda, ldb, ldc, M, N, K, ta, tb, accumulate, swapB, swapT;
    int tm, tn, ksplit, kper;         // tiles, k-splits, K range per split (multiple of GT_BK)
};

__global__ void to_bf16_kernel(const float* __restrict__ src, int64_t rows, int cols, int ld, int ldp,
                               __nv_bfloat16* __restrict__ dst) {
    const int upr = ldp / 8;                                   // 16-byte units per row
    const int64_t units = rows * upr;
    const bool vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < units; u += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = u / upr;
        const int c = (int)(u - r * upr) * 8;
        const float* s = src + r * ld + c;
        float v[8];
        if (c + 8 <= cols && vec_ok) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(s)), b = __ldg(reinterpret_cast<const float4*>(s) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = c + e < cols ? __ldg(s + e) : 0.f;
        }
        *reinterpret_cast<uint4*>(dst + r * ldp + c) =
            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// One operand tile: R rows (M or N index) x 64 k of the bf16 matrix `src` (leading dimension ld, a multiple
// of 8).  kcontig: element (r,k) at src[r*ld + k]; otherwise at src[k*ld + r].  Every 16-byte unit goes to
// core matrix (kgrp*(R/8) + rgrp)*128; out-of-range units are zero-filled (src-size 0).
template <int R>
__device__ __forceinline__ void load_tile(const __nv_bfloat16* __restrict__ src, int ld, int r0, int rmax, int k0, int kmax,
                                          bool kcontig, uint32_t dst, int lt) {
    constexpr int PER = R * 8 / GT_LOADERS;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int u = lt + i * GT_LOADERS;
        const __nv_bfloat16* ptr;
        bool ok;
        uint32_t doff;
        if (kcontig) {
            const int kc = u & 7, rl = u >> 3;                                // 8 lanes = one row's 128 bytes
            const int r = r0 + rl, k = k0 + kc * 8;
            ok = r < rmax && k < kmax;
            ptr = src + (size_t)r * ld + k;
            doff = (uint32_t)(kc * lbo_of(R, true) + (rl >> 3) * 128 + (rl & 7) * 16);
        } else {
            const int mc = u % (R / 8), kl = u / (R / 8);                     // lanes = consecutive 16-byte units of a k-row
            const int r = r0 + mc * 8, k = k0 + kl;
            ok = r < rmax && k < kmax;
            ptr = src + (size_t)k * ld + r;
            doff = (uint32_t)((kl >> 3) * lbo_of(R, false) + mc * 144 + (kl & 7) * 16);
        }
        cp_async16(dst + doff, ok ? (const void*)ptr : (const void*)src, ok ? 16u : 0u);
    }
}

__global__ void __launch_bounds__(GT_THREADS, 1) gemm_tc_kernel(GtParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[2 * GT_STAGES + 4];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t full = smem_u32(&bars[0]), empty = smem_u32(&bars[GT_STAGES]);
    const uint32_t tfull = smem_u32(&bars[2 * GT_STAGES]), tempty = tfull + 16;
    if (tid == 0) {
        for (int s = 0; s < GT_STAGES; ++s) { mbar_init(full + 8 * s, GT_LOADERS); mbar_init(empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull + 8 * b, 1); mbar_init(tempty + 8 * b, 128); }
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(smem_u32(&tmem_base_s), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int items = p.tm * p.tn * p.ksplit;
    // valid extents of the (zero-padded) bf16 copies along their contiguous dimension
    const int a_rmax = p.ta ? ((p.M + 7) & ~7) : p.M, b_rmax = p.tb ? p.N : ((p.N + 7) & ~7);

    if (warp < 4) {
        // ---------------- loaders ----------------
        uint32_t g = 0;
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
            const int split = it % p.ksplit, tile = it / p.ksplit;
            const int n0 = (tile % p.tn) * GT_BN, m0 = (tile / p.tn) * GT_BM;
            const int kbeg = split * p.kper, kend = min(p.K, kbeg + p.kper);
            const int kpad = p.ta && p.tb ? kend : kend;   // K rows/cols beyond kend are never read (zero-filled)
            for (int k0 = kbeg; k0 < kend; k0 += GT_BK, ++g) {
                const uint32_t slot = g % GT_STAGES, ph = (g / GT_STAGES) & 1;
                mbar_wait(empty + 8 * slot, ph ^ 1);
                const uint32_t sa = smem_u32(smem + slot * GT_STAGE_BYTES);
                load_tile<GT_BM>(p.A, p.lda, m0, a_rmax, k0, kpad, !p.ta, sa, tid);
                load_tile<GT_BN>(p.B, p.ldb, n0, b_rmax, k0, kpad, p.tb != 0, sa + GT_A_BYTES, tid);
                cp_async_arrive_noinc(full + 8 * slot);
            }
        }
    } else if (warp == 4) {
        // ---------------- MMA issuer (converged loop, elected lane) ----------------
        const uint32_t idesc = idesc_bf16(GT_BM, GT_BN, p.ta ? 1 : 0, p.tb ? 0 : 1);
        const bool a_kc = !p.ta, b_kc = p.tb != 0;
        const uint32_t a_lbo = a_kc ? lbo_of(GT_BM, true) : lbo_of(GT_BM, false), a_sbo = a_kc ? 128 : 144;
        const uint32_t b_lbo = b_kc ? lbo_of(GT_BN, true) : lbo_of(GT_BN, false), b_sbo = b_kc ? 128 : 144;
        const bool leader = elect_one();
        uint32_t g = 0, ti = 0;
        for (int it = blockIdx.x; it < items; it += gridDim.x, ++ti) {
            const int split = it % p.ksplit;
            const int kbeg = split * p.kper, kend = min(p.K, kbeg + p.kper);
            const uint32_t buf = ti & 1, tph = (ti >> 1) & 1;
            mbar_wait(tempty + 8 * buf, tph ^ 1);
            tc_fence_after();
            const uint32_t dcol = tmem + buf * GT_BN;
            bool first = true;
            for (int k0 = kbeg; k0 < kend; k0 += GT_BK, ++g) {
                const uint32_t slot = g % GT_STAGES, ph = (g / GT_STAGES) & 1;
                mbar_wait(full + 8 * slot, ph);
                fence_async_smem();
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + slot * GT_STAGE_BYTES), sb = sa + GT_A_BYTES;
#pragma unroll
                for (int kk = 0; kk < GT_BK / 16; ++kk) {
                    const uint64_t ad = smem_desc(sa + kk * 2 * a_lbo, a_lbo, a_sbo);
                    const uint64_t bd = smem_desc(sb + kk * 2 * b_lbo, b_lbo, b_sbo);
                    if (leader) mma_bf16(dcol, ad, bd, idesc, !(first && kk == 0));
                }
                first = false;
                if (leader) mma_commit(empty + 8 * slot);
            }
            if (leader) mma_commit(tfull + 8 * buf);
        }
    } else {
        // ---------------- epilogue: warps 5..8, TMEM lane quadrant = warp % 4 ----------------
        const int q = warp & 3;
        const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                            ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
        const bool atomic = p.ksplit > 1;
        uint32_t ti = 0;
        for (int it = blockIdx.x; it < items; it += gridDim.x, ++ti) {
            const int split = it % p.ksplit, tile = it / p.ksplit;
            const int n0 = (tile % p.tn) * GT_BN, m0 = (tile / p.tn) * GT_BM;
            const uint32_t buf = ti & 1, tph = (ti >> 1) & 1;
            mbar_wait(tfull + 8 * buf, tph);
            tc_fence_after();
            const int m = m0 + q * 32 + lane;
            size_t row = (size_t)m;
            if (p.swapB > 0 && m < p.M) row = (size_t)(m % p.swapB) * p.swapT + (size_t)(m / p.swapB);
            float* crow = p.C + row * p.ldc;
            const bool add_bias = p.bias != nullptr && split == 0;
#pragma unroll 1
            for (int c0 = 0; c0 < GT_BN; c0 += 32) {
                uint32_t v[32];
                const bool live = n0 + c0 < p.N;           // warp-uniform
                if (live) {
                    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + buf * GT_BN + c0, v);
                    tmem_ld_wait();
                }
                if (c0 + 32 == GT_BN) {                     // accumulator drained: release it before the stores
                    tc_fence_before();
                    mbar_arrive(tempty + 8 * buf);
                }
                if (!live || m >= p.M) continue;
                const int nb = n0 + c0;
                if (!atomic && vec_ok && nb + 32 <= p.N) {
#pragma unroll
                    for (int gq = 0; gq < 8; ++gq) {
                        float4 o = make_float4(__uint_as_float(v[4 * gq]), __uint_as_float(v[4 * gq + 1]),
                                               __uint_as_float(v[4 * gq + 2]), __uint_as_float(v[4 * gq + 3]));
                        if (add_bias) {
                            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + nb) + gq);
                            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                        }
                        float4* dst = reinterpret_cast<float4*>(crow + nb) + gq;
                        if (p.accumulate) { const float4 old = *dst; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                        *dst = o;
                    }
                } else {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) {
                        const int n = nb + jj;
                        if (n < p.N) {
                            float o = __uint_as_float(v[jj]);
                            if (add_bias) o += __ldg(p.bias + n);
                            if (atomic) atomicAdd(crow + n, o);
                            else crow[n] = p.accumulate ? crow[n] + o : o;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, 512);
}

__global__ void zero_rows_kernel(float* C, int M, int N, int ldc) {
    const int64_t n = (int64_t)M * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        C[(i / N) * ldc + (i % N)] = 0.f;
}

inline int pad8(int x) { return (x + 7) & ~7; }
inline size_t bf16_bytes(int rows, int cols) { return align_up((size_t)rows * pad8(cols) * 2, 256); }

}  // namespace

bool gemm_tc_supported(int M, int N, int K, int lda, int ldb, int ldc, int transa, int transb) {
    (void)lda; (void)ldb; (void)ldc; (void)transa; (void)transb;
    return M >= 1 && N >= 1 && K >= 1;
}

size_t gemm_tc_workspace(int M, int N, int K, int transa, int transb, int precision) {
    (void)precision;
    return 512 + (transa ? bf16_bytes(K, M) : bf16_bytes(M, K)) + (transb ? bf16_bytes(N, K) : bf16_bytes(K, N));
}

int gemm_tc(const float* A, int lda, const float* B, int ldb, const float* bias, int M, int N, int K, int transa,
            int transb, int accumulate, int precision, float* C, int ldc, int swapB, int swapT, void* workspace,
            size_t workspace_bytes, cudaStream_t st) {
    if (!workspace || workspace_bytes < gemm_tc_workspace(M, N, K, transa, transb, precision)) {
        set_error("gemm_tc: workspace too small (%zu < %zu)", workspace_bytes, gemm_tc_workspace(M, N, K, transa, transb, precision));
        return AMSS_ERR_WORKSPACE;
    }
    const int a_rows = transa ? K : M, a_cols = transa ? M : K;
    const int b_rows = transb ? N : K, b_cols = transb ? K : N;
    __nv_bfloat16* Ab = (__nv_bfloat16*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    __nv_bfloat16* Bb = (__nv_bfloat16*)((char*)Ab + bf16_bytes(a_rows, a_cols));
    {
        const int64_t ua = (int64_t)a_rows * (pad8(a_cols) / 8), ub = (int64_t)b_rows * (pad8(b_cols) / 8);
        AMSS_LAUNCH(to_bf16_kernel, (int)std::min<int64_t>((ua + 255) / 256, 16 * kNumSMs), 256, 0, st, A, (int64_t)a_rows,
                    a_cols, lda, pad8(a_cols), Ab);
        AMSS_LAUNCH(to_bf16_kernel, (int)std::min<int64_t>((ub + 255) / 256, 16 * kNumSMs), 256, 0, st, B, (int64_t)b_rows,
                    b_cols, ldb, pad8(b_cols), Bb);
    }
    GtParams p;
    p.A = Ab; p.B = Bb; p.bias = bias; p.C = C; p.lda = pad8(a_cols); p.ldb = pad8(b_cols); p.ldc = ldc;
    p.M = M; p.N = N; p.K = K; p.ta = transa; p.tb = transb; p.accumulate = accumulate; p.swapB = swapB; p.swapT = swapT;
    p.tm = (M + GT_BM - 1) / GT_BM; p.tn = (N + GT_BN - 1) / GT_BN;
    const int tiles = p.tm * p.tn;
    const int kstages = (K + GT_BK - 1) / GT_BK;
    int ksplit = 1;
    if (tiles < kNumSMs / 2 && kstages >= 8) ksplit = std::max(1, std::min(kNumSMs / tiles, kstages / 4));
    const int sper = (kstages + ksplit - 1) / ksplit;
    ksplit = (kstages + sper - 1) / sper;
    p.ksplit = ksplit;
    p.kper = sper * GT_BK;
    if (ksplit > 1 && !accumulate) {
        const int64_t n = (int64_t)M * N;
        AMSS_LAUNCH(zero_rows_kernel, (int)std::min<int64_t>((n + 255) / 256, 8 * kNumSMs), 256, 0, st, C, M, N, ldc);
    }
    const size_t smem = (size_t)GT_STAGES * GT_STAGE_BYTES;
    AMSS_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = std::min(tiles * ksplit, kNumSMs);
    AMSS_LAUNCH(gemm_tc_kernel, grid, GT_THREADS, smem, st, p);
    return AMSS_OK;
}

}  // namespace amss
