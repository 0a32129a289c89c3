// ld_probe -- hardware probe (test tooling, not part of the library) for two facts the BLSTM recurrence relies on:
//   1. the register layout of tcgen05.ld.16x256b.x2 (which TMEM lane / column lands in which thread / register) and
//      that the lane field of the address selects the upper 16 lanes of a warp's quadrant;
//   2. the cost of a chain of small TS-mode MMAs (A in TMEM, N = 16) as a function of the issue order: K steps into one
//      accumulator back to back vs. round-robin over several accumulators (the two orders of the backward recurrence).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I../adaptive-multispeaker-separation_b200/csrc -o ld_probe.bin ld_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc.cuh"

using namespace amss::tc;

__global__ void __launch_bounds__(128, 1) ld_probe_kernel(uint32_t* out) {
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    // value at (TMEM lane l, column c) = l * 256 + c
    for (int c0 = 0; c0 < 32; c0 += 8) {
        uint32_t w[8];
        for (int j = 0; j < 8; ++j) w[j] = (uint32_t)(tid * 256 + c0 + j);
        tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + c0, w);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    for (int half = 0; half < 2; ++half) {
        uint32_t v[8];
        tmem_ld_16x256b_x2(tmem + ((uint32_t)(warp * 32 + half * 16) << 16), v);
        tmem_ld_wait();
        for (int j = 0; j < 8; ++j) out[((half * 4 + warp) * 32 + lane) * 8 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 32);
}

// MMA chain timing: MT accumulators x KS K-steps, TS mode, N = 16, B operand = zeros in shared memory.
__global__ void __launch_bounds__(128, 1) mma_time_kernel(long long* out, int MT, int KS, int order, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 8192 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    {   // zero the A region
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 64; c < 512; c += 8) tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + c, w);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        const bool leader = elect_one();
        const uint32_t idesc = idesc_bf16(128, 16, 0, 0);
        const uint32_t baddr = smem_u32(smem);
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
            if (order == 0) {          // k outer, accumulators inner (round-robin)
                for (int kk = 0; kk < KS; ++kk)
                    for (int m = 0; m < MT; ++m) {
                        const uint64_t bd = smem_desc(baddr + kk * 512, 256, 128);
                        if (leader) mma_bf16_ts(tmem + m * 16, tmem + 64 + m * (8 * KS) + kk * 8, bd, idesc, kk > 0);
                    }
            } else {                   // accumulator outer, k inner (back-to-back chain)
                for (int m = 0; m < MT; ++m)
                    for (int kk = 0; kk < KS; ++kk) {
                        const uint64_t bd = smem_desc(baddr + kk * 512, 256, 128);
                        if (leader) mma_bf16_ts(tmem + m * 16, tmem + 64 + m * (8 * KS) + kk * 8, bd, idesc, kk > 0);
                    }
            }
            if (leader) mma_commit(smem_u32(&bar));
            const long long t1 = clock64();
            mbar_wait(smem_u32(&bar), r & 1);
            const long long t2 = clock64();
            if (leader) { out[2 * r] = t1 - t0; out[2 * r + 1] = t2 - t0; }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// Same chain issued by NW warps in parallel: warp w issues the K steps of accumulators w, w+NW, ... (acc-outer order) and commits
// to its own mbarrier; reports the time until ALL warps' MMAs have completed (measured by warp 0).
__global__ void __launch_bounds__(128, 1) mma_time_multi_kernel(long long* out, int MT, int KS, int NW, int reps, int M) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 8192 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    {
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 64; c < 512; c += 8) tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + c, w);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const bool leader = elect_one();
    const uint32_t idesc = idesc_bf16(M, 16, 0, 0);
    const uint32_t baddr = smem_u32(smem);
    for (int r = 0; r < reps; ++r) {
        __syncthreads();
        const long long t0 = clock64();
        if (warp < NW) {
            for (int m = warp; m < MT; m += NW)
                for (int kk = 0; kk < KS; ++kk) {
                    const uint64_t bd = smem_desc(baddr + kk * 512, 256, 128);
                    if (leader) mma_bf16_ts(tmem + m * 16, tmem + 64 + m * (8 * KS) + kk * 8, bd, idesc, kk > 0);
                }
            if (leader) mma_commit(smem_u32(&bar[warp]));
        }
        __syncwarp();
        if (warp == 0) {
            for (int w = 0; w < NW; ++w) mbar_wait(smem_u32(&bar[w]), r & 1);
            const long long t2 = clock64();
            if (leader) out[r] = t2 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    uint32_t* d_out;
    cudaMalloc(&d_out, 2 * 4 * 32 * 8 * 4);
    ld_probe_kernel<<<1, 128>>>(d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("ld probe: CUDA ERROR %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<uint32_t> h(2 * 4 * 32 * 8);
    cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost);
    // hypothesis: thread t, register j (x2: j = 4*rep + r): lane = base + t/4 + 8*(r/2), column = 8*rep + 2*(t%4) + (r%2)
    int bad = 0;
    for (int half = 0; half < 2; ++half)
        for (int w = 0; w < 4; ++w)
            for (int t = 0; t < 32; ++t)
                for (int j = 0; j < 8; ++j) {
                    const uint32_t v = h[((half * 4 + w) * 32 + t) * 8 + j];
                    const int rep = j / 4, r = j % 4;
                    const uint32_t want = (uint32_t)((w * 32 + half * 16 + t / 4 + 8 * (r / 2)) * 256 + 8 * rep + 2 * (t % 4) + (r % 2));
                    if (v != want) {
                        if (bad < 16) printf("  mismatch half %d warp %d thread %d reg %d: lane %u col %u (hypothesis lane %u col %u)\n", half, w, t,
                                             j, v / 256, v % 256, want / 256, want % 256);
                        ++bad;
                    }
                }
    printf("tcgen05.ld.16x256b.x2 layout hypothesis (lane = base + t/4 + 8*(r/2), col = 8*rep + 2*(t%%4) + r%%2; address lane +16 = upper half): %s (%d mismatches)\n",
           bad ? "MISMATCH" : "OK", bad);
    if (bad) {
        for (int t = 0; t < 8; ++t) {
            printf("  warp 0 half 0 thread %d:", t);
            for (int j = 0; j < 8; ++j) { const uint32_t v = h[(t)*8 + j]; printf(" (l%u,c%u)", v / 256, v % 256); }
            printf("\n");
        }
    }

    long long* d_t;
    cudaMalloc(&d_t, 64 * 8);
    {   // bring the clocks up: ~200 ms of MMA chains on every SM before the timed launches
        cudaFuncSetAttribute(mma_time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
        long long* d_w;
        cudaMalloc(&d_w, 148 * 4096 * 16);
        for (int i = 0; i < 20; ++i) mma_time_kernel<<<1, 128, 16384>>>(d_w, 1, 19, 1, 2000);
        cudaDeviceSynchronize();
    }
    cudaFuncSetAttribute(mma_time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    const int cfgs[][3] = {{1, 19, 1}, {1, 8, 1}, {3, 8, 0}, {3, 8, 1}, {2, 8, 0}, {2, 8, 1}, {2, 19, 0}, {2, 19, 1}, {1, 1, 1}, {1, 2, 1}, {1, 4, 1}};
    for (auto& c : cfgs) {
        mma_time_kernel<<<1, 128, 16384>>>(d_t, c[0], c[1], c[2], 8);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mma timing: CUDA ERROR %s\n", cudaGetErrorString(e)); return 2; }
        long long ht[16];
        cudaMemcpy(ht, d_t, sizeof(ht), cudaMemcpyDeviceToHost);
        printf("MMA chain TS N=16: %d accumulators x %2d k-steps, %s: issue %lld clk, issue+complete %lld clk (%.1f clk per MMA)\n", c[0], c[1],
               c[2] ? "acc-outer" : "k-outer  ", ht[14], ht[15], (double)ht[15] / (c[0] * c[1]));
    }
    cudaFuncSetAttribute(mma_time_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    const int mc[][4] = {{3, 8, 1, 128}, {3, 8, 3, 128}, {2, 19, 1, 128}, {2, 19, 2, 128}, {4, 8, 2, 128}, {4, 8, 4, 128}, {3, 8, 1, 64}, {6, 8, 1, 64}, {6, 8, 2, 64}};
    for (auto& c : mc) {
        mma_time_multi_kernel<<<1, 128, 16384>>>(d_t, c[0], c[1], c[2], 8, c[3]);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("multi-issuer timing: CUDA ERROR %s\n", cudaGetErrorString(e)); return 2; }
        long long ht[8];
        cudaMemcpy(ht, d_t, sizeof(ht), cudaMemcpyDeviceToHost);
        printf("MMA chains TS M=%d N=16: %d accumulators x %2d k-steps issued by %d warp(s): all complete after %lld clk (%.1f clk per MMA)\n", c[3], c[0],
               c[1], c[2], ht[7], (double)ht[7] / (c[0] * c[1]));
    }
    return 0;
}
