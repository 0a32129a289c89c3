// Dense GEMM on the 5th-gen tensor cores: C[M,N] (+)= op(A) op(B) (+ bias), fp32 in / fp32 out,
// bf16 operands, fp32 accumulation in TMEM.  Used (AMSS_PREC_BF16) for the hoisted BLSTM input
// projections, the embedding head (utils/ops.py:501-503) and every backward GEMM of those.
//
//   1. pack: each operand is converted once to bf16 AND laid out as the exact shared-memory image the
//      MMA wants -- K-major core matrices, tiles of RT rows x 64 k stored contiguously (A: RT = 128,
//      16 KB; B: RT = 256, 32 KB), zero padded.  Transposed operands (dW = X^T dZ, dX = dZ W^T) are
//      transposed by the pack kernel's addressing (coalesced reads either way), so the GEMM kernel
//      only ever sees K-major x K-major;
//   2. persistent CTAs (one per SM) walk (tile, k-split) work items; ONE thread feeds a 4-stage
//      mbarrier ring with two cp.async.bulk (TMA engine) copies per stage (48 KB), i.e. ~190 KB of
//      loads in flight per SM and no per-element load instructions at all;
//   3. one warp issues tcgen05.mma 128x256x16 (converged loop, elected lane);
//   4. TMEM accumulators are double buffered (2 x 256 columns): 4 epilogue warps drain tile i
//      (bias, row remap, accumulate / split-K red.global.add) while tile i+1 is being multiplied.
#include "common.cuh"
#include "tc.cuh"
#include <algorithm>

namespace amss {
namespace {

using namespace tc;

constexpr int GT_BM = 128, GT_BN = 256, GT_BK = 64, GT_STAGES = 4;
constexpr int GT_THREADS = 192;                      // warp 0 loader, 1 MMA (+TMEM alloc), 2-5 epilogue
constexpr int GT_A_BYTES = GT_BM * GT_BK * 2, GT_B_BYTES = GT_BN * GT_BK * 2;
constexpr int GT_STAGE_BYTES = GT_A_BYTES + GT_B_BYTES;

struct GtParams {
    const uint8_t *A, *B;             // packed bf16 tiles: A [tm][KS][16 KB], B [tn][KS][32 KB]
    const float* bias;
    float* C;
    int ldc, M, N, K, KS, accumulate, swapB, swapT;
    int tm, tn, ksplit, sper;         // tiles, k-splits, stages per split
};

// src fp32 -> packed bf16 tiles.  Logical operand X[r][k], r < R, k < K: kcontig: X[r][k] = src[r*ld + k],
// otherwise X[r][k] = src[k*ld + r].  16-byte unit (r, k8 = k/8) of tile (r/RT, k/64) goes to
//   ((r/RT)*KS + k/64) * RT*128  +  ((k8 % 8) * (RT/8) + (r % RT)/8) * 128  +  (r % 8) * 16.
// Consecutive threads take consecutive r of one k8: 16-byte writes are contiguous in groups of 8 rows,
// reads are 32-byte runs (kcontig) or 4-byte elements coalesced across the warp (transposed source).
template <int RT>
__global__ void pack_bf16_kernel(const float* __restrict__ src, int R, int K, int ld, int kcontig, int KS, int rtiles,
                                 uint4* __restrict__ dst) {
    const int64_t rp = (int64_t)rtiles * RT;                      // padded rows
    const int64_t units = rp * KS * 8;
    const bool vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < units; u += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = u % rp;
        const int k8 = (int)(u / rp);                              // 0 .. KS*8-1
        const int k0 = k8 * 8;
        float v[8];
        if (r < R && k0 < K) {
            if (kcontig) {
                const float* s = src + r * ld + k0;
                if (k0 + 8 <= K && vec_ok) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(s)), b = __ldg(reinterpret_cast<const float4*>(s) + 1);
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = k0 + e < K ? __ldg(s + e) : 0.f;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = k0 + e < K ? __ldg(src + (size_t)(k0 + e) * ld + r) : 0.f;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = 0.f;
        }
        const int rt = (int)(r / RT), rl = (int)(r % RT), ks = k8 >> 3, kc = k8 & 7;
        const size_t off = ((size_t)rt * KS + ks) * (RT * 8) + (size_t)(kc * (RT / 8) + (rl >> 3)) * 8 + (rl & 7);
        dst[off] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
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
        for (int s = 0; s < GT_STAGES; ++s) { mbar_init(full + 8 * s, 1); mbar_init(empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull + 8 * b, 1); mbar_init(tempty + 8 * b, 128); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int items = p.tm * p.tn * p.ksplit;

    if (warp == 0) {
        // ---------------- loader: two bulk copies per stage (one lane) ----------------
        if (lane == 0) {
            uint32_t g = 0;
            for (int it = blockIdx.x; it < items; it += gridDim.x) {
                const int split = it % p.ksplit, tile = it / p.ksplit;
                const int nt = tile % p.tn, mt = tile / p.tn;
                const int sbeg = split * p.sper, send = min(p.KS, sbeg + p.sper);
                const uint8_t* pa = p.A + ((size_t)mt * p.KS + sbeg) * GT_A_BYTES;
                const uint8_t* pb = p.B + ((size_t)nt * p.KS + sbeg) * GT_B_BYTES;
                for (int j = sbeg; j < send; ++j, ++g, pa += GT_A_BYTES, pb += GT_B_BYTES) {
                    const uint32_t slot = g % GT_STAGES, ph = (g / GT_STAGES) & 1;
                    mbar_wait(empty + 8 * slot, ph ^ 1);
                    const uint32_t sa = smem_u32(smem + slot * GT_STAGE_BYTES);
                    mbar_expect_tx(full + 8 * slot, GT_STAGE_BYTES);
                    bulk_g2s(sa, pa, GT_A_BYTES, full + 8 * slot);
                    bulk_g2s(sa + GT_A_BYTES, pb, GT_B_BYTES, full + 8 * slot);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (converged loop, elected lane) ----------------
        const uint32_t idesc = idesc_bf16(GT_BM, GT_BN, 0, 0);
        const bool leader = elect_one();
        uint32_t g = 0, ti = 0;
        for (int it = blockIdx.x; it < items; it += gridDim.x, ++ti) {
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
                    const uint64_t ad = smem_desc(sa + kk * 2 * (GT_BM / 8) * 128, (GT_BM / 8) * 128, 128);
                    const uint64_t bd = smem_desc(sb + kk * 2 * (GT_BN / 8) * 128, (GT_BN / 8) * 128, 128);
                    if (leader) mma_bf16(dcol, ad, bd, idesc, !(j == sbeg && kk == 0));
                }
                if (leader) mma_commit(empty + 8 * slot);
            }
            if (leader) mma_commit(tfull + 8 * buf);
        }
    } else {
        // ---------------- epilogue: warps 2..5, TMEM lane quadrant = warp % 4 ----------------
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
    if (warp == 1) tmem_dealloc(tmem, 512);
}

__global__ void zero_rows_kernel(float* C, int M, int N, int ldc) {
    const int64_t n = (int64_t)M * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        C[(i / N) * ldc + (i % N)] = 0.f;
}

inline size_t packed_bytes(int R, int K, int RT) {
    return (size_t)((R + RT - 1) / RT) * ((K + GT_BK - 1) / GT_BK) * RT * 128;
}

}  // namespace

bool gemm_tc_supported(int M, int N, int K, int lda, int ldb, int ldc, int transa, int transb) {
    (void)lda; (void)ldb; (void)ldc; (void)transa; (void)transb;
    return M >= 1 && N >= 1 && K >= 1;
}

size_t gemm_tc_workspace(int M, int N, int K, int transa, int transb, int precision) {
    (void)precision; (void)transa; (void)transb;
    return 1024 + align_up(packed_bytes(M, K, GT_BM), 256) + align_up(packed_bytes(N, K, GT_BN), 256);
}

int gemm_tc(const float* A, int lda, const float* B, int ldb, const float* bias, int M, int N, int K, int transa,
            int transb, int accumulate, int precision, float* C, int ldc, int swapB, int swapT, void* workspace,
            size_t workspace_bytes, cudaStream_t st) {
    if (!workspace || workspace_bytes < gemm_tc_workspace(M, N, K, transa, transb, precision)) {
        set_error("gemm_tc: workspace too small (%zu < %zu)", workspace_bytes, gemm_tc_workspace(M, N, K, transa, transb, precision));
        return AMSS_ERR_WORKSPACE;
    }
    GtParams p;
    p.tm = (M + GT_BM - 1) / GT_BM; p.tn = (N + GT_BN - 1) / GT_BN;
    p.KS = (K + GT_BK - 1) / GT_BK;
    uint8_t* Ap = (uint8_t*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
    uint8_t* Bp = Ap + align_up(packed_bytes(M, K, GT_BM), 256);
    {
        // A[m][k]: not transposed -> src[m*lda + k] (K-contiguous); transposed -> src[k*lda + m]
        const int64_t ua = (int64_t)p.tm * GT_BM * p.KS * 8, ub = (int64_t)p.tn * GT_BN * p.KS * 8;
        AMSS_LAUNCH((pack_bf16_kernel<GT_BM>), (int)std::min<int64_t>((ua + 255) / 256, 16 * kNumSMs), 256, 0, st, A, M, K, lda,
                    transa ? 0 : 1, p.KS, p.tm, (uint4*)Ap);
        // B^T[n][k]: not transposed -> src[k*ldb + n]; transposed -> src[n*ldb + k] (K-contiguous)
        AMSS_LAUNCH((pack_bf16_kernel<GT_BN>), (int)std::min<int64_t>((ub + 255) / 256, 16 * kNumSMs), 256, 0, st, B, N, K, ldb,
                    transb ? 1 : 0, p.KS, p.tn, (uint4*)Bp);
    }
    p.A = Ap; p.B = Bp; p.bias = bias; p.C = C; p.ldc = ldc;
    p.M = M; p.N = N; p.K = K; p.accumulate = accumulate; p.swapB = swapB; p.swapT = swapT;
    const int tiles = p.tm * p.tn;
    int ksplit = 1;
    if (tiles < kNumSMs / 2 && p.KS >= 8) ksplit = std::max(1, std::min(kNumSMs / tiles, p.KS / 4));
    const int sper = (p.KS + ksplit - 1) / ksplit;
    ksplit = (p.KS + sper - 1) / sper;
    p.ksplit = ksplit;
    p.sper = sper;
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
