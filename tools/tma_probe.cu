// Probe: TMA tensor-map loads (SWIZZLE_128B) feeding tcgen05.mma with swizzled shared-memory descriptors, for K-major and
// MN-major bf16 operands taken straight from row-major global matrices.  Pins the descriptor fields gemm_tc.cu relies on:
//   K-major : box {64 k, R rows}; rows 128 B apart, 8-row groups SBO = 1024 B; K step of 16 = +32 B on the start address
//   MN-major: boxes {64 mn, 64 k}; k rows 128 B apart, 8-k groups SBO = 1024 B, 64-wide mn atoms LBO = 8192 B apart;
//             K step of 16 = +2048 B on the start address
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I adaptive-multispeaker-separation_b200/csrc -o tools/tma_probe.bin tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "tc.cuh"

using namespace amss::tc;

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

constexpr int BM = 128, BN = 256, BK = 64;

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                        int a_mn, int b_mn, int KB, int m0, int n0, float* C) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t full = smem_u32(&bars[0]), done = smem_u32(&bars[1]);
    if (tid == 0) { mbar_init(full, 1); mbar_init(done, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t sa = smem_u32(smem), sb = sa + BM * BK * 2;
    const uint32_t idesc = idesc_bf16(BM, BN, a_mn, b_mn);
    for (int kb = 0; kb < KB; ++kb) {
        if (tid == 0) {
            mbar_expect_tx(full, (BM + BN) * BK * 2);
            if (!a_mn) tma_load_2d(sa, &mapA, kb * BK, m0, full);
            else for (int i = 0; i < BM / 64; ++i) tma_load_2d(sa + i * 8192, &mapA, m0 + i * 64, kb * BK, full);
            if (!b_mn) tma_load_2d(sb, &mapB, kb * BK, n0, full);
            else for (int i = 0; i < BN / 64; ++i) tma_load_2d(sb + i * 8192, &mapB, n0 + i * 64, kb * BK, full);
        }
        mbar_wait(full, kb & 1);
        tc_fence_after();
        if (tid == 0) {
            for (int kk = 0; kk < BK / 16; ++kk) {
                const uint64_t ad = a_mn ? smem_desc_sw128(sa + kk * 2048, 8192, 1024) : smem_desc_sw128(sa + kk * 32, 16, 1024);
                const uint64_t bd = b_mn ? smem_desc_sw128(sb + kk * 2048, 8192, 1024) : smem_desc_sw128(sb + kk * 32, 16, 1024);
                mma_bf16(tmem, ad, bd, idesc, !(kb == 0 && kk == 0));
            }
            mma_commit(done);
        }
        mbar_wait(done, kb & 1);
        tc_fence_after();
    }
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) C[(size_t)(warp * 32 + lane) * BN + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static uint16_t f2bf(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    u += 0x7FFF + ((u >> 16) & 1);
    return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    const int KB = 3, K = KB * BK - 8;        // ragged K: the last box is zero filled past K
    const int R_A = 296, R_B = 400;           // operand extents along M / N (both tiles run past the end: zero fill)
    const int m0 = 256, n0 = 256;
    int nfail = 0;
    for (int cs = 0; cs < 4; ++cs) {
        const int a_mn = cs & 1, b_mn = cs >> 1;
        // logical X[r][k]; K-major storage: src[r*ld + k] (ld = K); MN-major storage: src[k*ld + r] (ld = R)
        std::vector<uint16_t> hA((size_t)R_A * K), hB((size_t)R_B * K);
        std::vector<float> fA((size_t)R_A * K), fB((size_t)R_B * K);
        srand(7 + cs);
        for (int r = 0; r < R_A; ++r) for (int k = 0; k < K; ++k) {
            const uint16_t h = f2bf((float)(rand() % 2001 - 1000) / 1000.f);
            fA[(size_t)r * K + k] = bf2f(h);
            hA[a_mn ? (size_t)k * R_A + r : (size_t)r * K + k] = h;
        }
        for (int r = 0; r < R_B; ++r) for (int k = 0; k < K; ++k) {
            const uint16_t h = f2bf((float)(rand() % 2001 - 1000) / 1000.f);
            fB[(size_t)r * K + k] = bf2f(h);
            hB[b_mn ? (size_t)k * R_B + r : (size_t)r * K + k] = h;
        }
        uint16_t *dA, *dB; float* dC;
        cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dC, BM * BN * 4);
        cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
        CUtensorMap mA, mB;
        auto make = [&](CUtensorMap* m, void* ptr, int R, int mn, int rows_box) {
            cuuint64_t dims[2], strides[1]; cuuint32_t box[2], es[2] = {1, 1};
            if (!mn) { dims[0] = K; dims[1] = R; strides[0] = (cuuint64_t)K * 2; box[0] = 64; box[1] = rows_box; }
            else { dims[0] = R; dims[1] = K; strides[0] = (cuuint64_t)R * 2; box[0] = 64; box[1] = 64; }
            return encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        };
        CUresult r1 = make(&mA, dA, R_A, a_mn, BM), r2 = make(&mB, dB, R_B, b_mn, BN);
        if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("case %d: encode failed %d %d\n", cs, (int)r1, (int)r2); ++nfail; continue; }
        const size_t smem = (BM + BN) * BK * 2 + 1024;
        cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe_kernel<<<1, 128, smem>>>(mA, mB, a_mn, b_mn, KB, m0, n0, dC);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> hC(BM * BN);
        cudaMemcpy(hC.data(), dC, BM * BN * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < BM; ++m) for (int n = 0; n < BN; ++n) {
            double ref = 0;
            if (m0 + m < R_A && n0 + n < R_B)
                for (int k = 0; k < K; ++k) ref += (double)fA[(size_t)(m0 + m) * K + k] * fB[(size_t)(n0 + n) * K + k];
            maxerr = fmax(maxerr, fabs(ref - hC[m * BN + n])); maxref = fmax(maxref, fabs(ref));
        }
        const bool ok = e == cudaSuccess && maxerr < 1e-3 * maxref;
        printf("TMA+SW128 case A %s / B %s: %s (cuda %s, max err %.3g, max |ref| %.3g)\n", a_mn ? "MN-major" : "K-major ",
               b_mn ? "MN-major" : "K-major ", ok ? "MATCH" : "MISMATCH", cudaGetErrorString(e), maxerr, maxref);
        if (!ok) ++nfail;
        if (e != cudaSuccess) return 2;
        cudaFree(dA); cudaFree(dB); cudaFree(dC);
    }
    printf("SUMMARY: %d of 4 cases failed\n", nfail);
    return nfail ? 1 : 0;
}
