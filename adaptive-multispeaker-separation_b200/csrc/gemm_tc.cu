// Dense GEMM on the 5th-gen tensor cores: C[M,N] (+)= op(A) op(B) (+ bias), fp32 in / fp32 out,
// bf16 operands, fp32 accumulation in TMEM.  Used (AMSS_PREC_BF16) for the hoisted BLSTM input
// projections, the embedding head (utils/ops.py:501-503) and every backward GEMM of those.
//
// Tile 128 (M, TMEM lanes) x 256 (N, TMEM columns) x 64 (K per stage), 4-stage mbarrier ring.
// The fp32 -> bf16 conversion is fused into the operand load: 8 loader warps read 32-byte runs of
// the fp32 matrices and store 16-byte bf16 units straight into the canonical no-swizzle
// core-matrix layout -- K-major when the matrix is K-contiguous in memory, MN-major when it is
// M/N-contiguous -- so transposed operands (dW = X^T dZ, dX = dZ W^T) need no transpose pass.
// Small-MN / large-K products (weight gradients) are split along K over CTAs and reduced with
// red.global.add.f32.
#include "common.cuh"
#include "tc.cuh"
#include <algorithm>

namespace amss {
namespace {

using namespace tc;

constexpr int GT_BM = 128, GT_BN = 256, GT_BK = 64, GT_STAGES = 4;
constexpr int GT_LOADERS = 256, GT_THREADS = GT_LOADERS + 32;
constexpr int GT_A_BYTES = GT_BM * GT_BK * 2, GT_B_BYTES = GT_BN * GT_BK * 2;
constexpr int GT_STAGE_BYTES = GT_A_BYTES + GT_B_BYTES;

struct GtParams {
    const float *A, *B, *bias;
    float* C;
    int lda, ldb, ldc, M, N, K, ta, tb, accumulate, swapB, swapT;
    int ksplit, kper;     // K range per split (multiple of GT_BK)
};

// One operand tile: R rows (M or N index) x 64 k.  `kcontig`: the matrix is K-contiguous in memory
// (element (r,k) at src[r*ld + k]); otherwise MN-contiguous (element (r,k) at src[k*ld + r]).
// Both write 16-byte units into core matrices placed at (kgrp*(R/8) + rgrp)*128.
template <int R>
__device__ __forceinline__ void load_tile(const float* __restrict__ src, int ld, int r0, int rmax, int k0, int kmax,
                                          bool kcontig, uint8_t* dst, int lt) {
    constexpr int UNITS = R * 8;
    constexpr int PER = UNITS / GT_LOADERS;
    constexpr int BATCH = PER <= 8 ? PER : 8;
    const bool vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
#pragma unroll 1
    for (int base = 0; base < PER; base += BATCH) {
        float4 v[BATCH][2];
        int doff[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
            const int u = lt + (base + i) * GT_LOADERS;
            int r, k;
            if (kcontig) {
                const int rl = u % R, kc = u / R;                    // lanes -> consecutive rows
                r = r0 + rl; k = k0 + kc * 8;
                doff[i] = (kc * (R / 8) + (rl >> 3)) * 128 + (rl & 7) * 16;
            } else {
                const int kl = (u & 7) | ((u / R) << 3), mc = (u >> 3) % (R / 8);   // lanes -> consecutive k
                r = r0 + mc * 8; k = k0 + kl;
                doff[i] = ((kl >> 3) * (R / 8) + mc) * 128 + (kl & 7) * 16;
            }
            const float* ptr = kcontig ? src + (size_t)r * ld + k : src + (size_t)k * ld + r;
            const int run_pos = kcontig ? k : r, run_max = kcontig ? kmax : rmax;
            const bool other_ok = kcontig ? (r < rmax) : (k < kmax);
            if (other_ok && run_pos + 8 <= run_max && vec_ok && ((run_pos & 3) == 0)) {
                v[i][0] = __ldg(reinterpret_cast<const float4*>(ptr));
                v[i][1] = __ldg(reinterpret_cast<const float4*>(ptr) + 1);
            } else {
                float t[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) t[e] = (other_ok && run_pos + e < run_max) ? __ldg(ptr + e) : 0.f;
                v[i][0] = make_float4(t[0], t[1], t[2], t[3]);
                v[i][1] = make_float4(t[4], t[5], t[6], t[7]);
            }
        }
#pragma unroll
        for (int i = 0; i < BATCH; ++i)
            *reinterpret_cast<uint4*>(dst + doff[i]) =
                make_uint4(pack_bf16(v[i][0].x, v[i][0].y), pack_bf16(v[i][0].z, v[i][0].w),
                           pack_bf16(v[i][1].x, v[i][1].y), pack_bf16(v[i][1].z, v[i][1].w));
    }
}

__global__ void __launch_bounds__(GT_THREADS, 1) gemm_tc_kernel(GtParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[2 * GT_STAGES + 1];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * GT_BN, m0 = blockIdx.y * GT_BM, split = blockIdx.z;
    const int kbeg = split * p.kper, kend = min(p.K, kbeg + p.kper);
    const int nstage = (kend - kbeg + GT_BK - 1) / GT_BK;
    const uint32_t full = smem_u32(&bars[0]), empty = smem_u32(&bars[GT_STAGES]), done = smem_u32(&bars[2 * GT_STAGES]);
    if (tid == 0) {
        for (int s = 0; s < GT_STAGES; ++s) { mbar_init(full + 8 * s, GT_LOADERS); mbar_init(empty + 8 * s, 1); }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 8) tmem_alloc(smem_u32(&tmem_base_s), GT_BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp < 8) {
        // ---------------- loaders: fp32 global -> bf16 core matrices ----------------
        for (int j = 0; j < nstage; ++j) {
            const int slot = j % GT_STAGES, ph = (j / GT_STAGES) & 1;
            mbar_wait(empty + 8 * slot, ph ^ 1);
            uint8_t* sa = smem + slot * GT_STAGE_BYTES;
            const int k0 = kbeg + j * GT_BK;
            load_tile<GT_BM>(p.A, p.lda, m0, p.M, k0, kend, !p.ta, sa, tid);
            load_tile<GT_BN>(p.B, p.ldb, n0, p.N, k0, kend, p.tb != 0, sa + GT_A_BYTES, tid);
            fence_async_smem();
            mbar_arrive(full + 8 * slot);
        }
    } else if (lane == 0) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = idesc_bf16(GT_BM, GT_BN, p.ta ? 1 : 0, p.tb ? 0 : 1);
        for (int j = 0; j < nstage; ++j) {
            const int slot = j % GT_STAGES, ph = (j / GT_STAGES) & 1;
            mbar_wait(full + 8 * slot, ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + slot * GT_STAGE_BYTES), sb = sa + GT_A_BYTES;
#pragma unroll
            for (int kk = 0; kk < GT_BK / 16; ++kk) {
                const uint64_t ad = smem_desc(sa + kk * 2 * (GT_BM / 8) * 128, (GT_BM / 8) * 128, 128);
                const uint64_t bd = smem_desc(sb + kk * 2 * (GT_BN / 8) * 128, (GT_BN / 8) * 128, 128);
                mma_bf16(tmem, ad, bd, idesc, (j | kk) != 0);
            }
            mma_commit(empty + 8 * slot);
        }
        mma_commit(done);
    }
    // ---------------- epilogue: warps 0..3, TMEM lane quadrant = warp ----------------
    if (warp < 4) {
        mbar_wait(done, 0);
        tc_fence_after();
        const int m = m0 + warp * 32 + lane;
        size_t row = (size_t)m;
        if (p.swapB > 0 && m < p.M) row = (size_t)(m % p.swapB) * p.swapT + (size_t)(m / p.swapB);
        float* crow = p.C + row * p.ldc;
        const bool atomic = p.ksplit > 1;
        const bool add_bias = p.bias != nullptr && split == 0;
        const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                            ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
#pragma unroll 1
        for (int c0 = 0; c0 < GT_BN; c0 += 32) {
            if (n0 + c0 >= p.N) break;            // warp-uniform
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            tmem_ld_wait();
            if (m >= p.M) continue;
            const int nb = n0 + c0;
            if (!atomic && vec_ok && nb + 32 <= p.N) {
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    float4 o = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                                           __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
                    if (add_bias) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + nb) + g);
                        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                    }
                    float4* dst = reinterpret_cast<float4*>(crow + nb) + g;
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
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, GT_BN);
}

__global__ void zero_rows_kernel(float* C, int M, int N, int ldc) {
    const int64_t n = (int64_t)M * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        C[(i / N) * ldc + (i % N)] = 0.f;
}

}  // namespace

bool gemm_tc_supported(int M, int N, int K, int lda, int ldb, int ldc, int transa, int transb) {
    (void)lda; (void)ldb; (void)ldc; (void)transa; (void)transb;
    return M >= 1 && N >= 1 && K >= 1;
}

size_t gemm_tc_workspace(int, int, int, int, int, int) { return 256; }

int gemm_tc(const float* A, int lda, const float* B, int ldb, const float* bias, int M, int N, int K, int transa,
            int transb, int accumulate, int precision, float* C, int ldc, int swapB, int swapT, void* workspace,
            size_t workspace_bytes, cudaStream_t st) {
    (void)precision; (void)workspace; (void)workspace_bytes;
    GtParams p;
    p.A = A; p.B = B; p.bias = bias; p.C = C; p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
    p.ta = transa; p.tb = transb; p.accumulate = accumulate; p.swapB = swapB; p.swapT = swapT;
    const int tm = (M + GT_BM - 1) / GT_BM, tn = (N + GT_BN - 1) / GT_BN;
    const int kstages = (K + GT_BK - 1) / GT_BK;
    int ksplit = 1;
    if (tm * tn < kNumSMs / 2 && kstages >= 8) ksplit = std::max(1, std::min(kNumSMs / (tm * tn), kstages / 4));
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
    dim3 grid(tn, tm, ksplit);
    AMSS_LAUNCH(gemm_tc_kernel, grid, GT_THREADS, smem, st, p);
    return AMSS_OK;
}

}  // namespace amss
