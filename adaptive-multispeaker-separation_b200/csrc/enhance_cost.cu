// Tail of the enhance layer (models/network.py:640-693), fused: the Conv1D output [B,S,TF] -> softmax over the S sources
// (or tanh / identity, --nonlinearity) -> masks * X_input -> the S x S table of squared distances to the sources'
// magnitudes that the permutation-invariant cost needs:
//     D[b][s][k] = sum_j (X_non_mix[b,j,s] - est[b,k,j])^2,   est[b,k,j] = mask[b,j,k] * X_input[b,j].
// The host picks the best of the S! permutations from the [B,S,S] table (a handful of scalars); the backward kernel
// turns the chosen permutation into d cost / d logits in one pass.  HBM-bound: 4*(2S+1)*TF bytes per mixture forward.
#include "common.cuh"
#include <algorithm>

namespace amss {
namespace {

constexpr int EC_THREADS = 256;
constexpr int EC_MAXS = 4;

__device__ __forceinline__ void ec_masks(const float* z, int S, int nonlin, float* m) {
    if (nonlin == 0) {                       // softmax over the sources
        float mx = z[0];
#pragma unroll
        for (int s = 1; s < EC_MAXS; ++s) if (s < S) mx = fmaxf(mx, z[s]);
        float tot = 0.f;
#pragma unroll
        for (int s = 0; s < EC_MAXS; ++s) if (s < S) { m[s] = expf(z[s] - mx); tot += m[s]; }
#pragma unroll
        for (int s = 0; s < EC_MAXS; ++s) if (s < S) m[s] /= tot;
    } else {
#pragma unroll
        for (int s = 0; s < EC_MAXS; ++s) if (s < S) m[s] = nonlin == 1 ? tanhf(z[s]) : z[s];
    }
}

// grid = (chunks, B).  part[b][chunk][s][k]
__global__ void __launch_bounds__(EC_THREADS)
enhance_table_kernel(const float* __restrict__ logits, const float* __restrict__ X, const float* __restrict__ tgt,
                     int S, int64_t TF, int nonlin, float* __restrict__ masks, float* __restrict__ part) {
    __shared__ float red[32];
    const int b = blockIdx.y;
    float acc[EC_MAXS][EC_MAXS];
#pragma unroll
    for (int s = 0; s < EC_MAXS; ++s)
#pragma unroll
        for (int k = 0; k < EC_MAXS; ++k) acc[s][k] = 0.f;
    for (int64_t j = blockIdx.x * (int64_t)EC_THREADS + threadIdx.x; j < TF; j += (int64_t)gridDim.x * EC_THREADS) {
        float z[EC_MAXS], m[EC_MAXS], t[EC_MAXS];
#pragma unroll
        for (int s = 0; s < EC_MAXS; ++s)
            if (s < S) { z[s] = logits[((size_t)b * S + s) * TF + j]; t[s] = tgt[((size_t)b * TF + j) * S + s]; }
        ec_masks(z, S, nonlin, m);
        const float x = X[(size_t)b * TF + j];
#pragma unroll
        for (int k = 0; k < EC_MAXS; ++k)
            if (k < S) {
                if (masks) masks[((size_t)b * TF + j) * S + k] = m[k];
                const float e = m[k] * x;
#pragma unroll
                for (int s = 0; s < EC_MAXS; ++s)
                    if (s < S) { const float d = t[s] - e; acc[s][k] = fmaf(d, d, acc[s][k]); }
            }
    }
    for (int s = 0; s < S; ++s)
        for (int k = 0; k < S; ++k) {
            const float v = block_sum(acc[s][k], red);
            if (threadIdx.x == 0) part[(((size_t)b * gridDim.x + blockIdx.x) * S + s) * S + k] = v;
        }
}
__global__ void enhance_table_finish_kernel(const float* __restrict__ part, int B, int chunks, int SS, float* __restrict__ table) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * SS) {
        const int b = i / SS, q = i - b * SS;
        float a = 0.f;
        for (int c = 0; c < chunks; ++c) a += part[((size_t)b * chunks + c) * SS + q];
        table[i] = a;
    }
}
// perm[b][s] = the estimate paired with source s in the chosen permutation; dcost_b[b] = d cost / d (cost of mixture b)
__global__ void __launch_bounds__(EC_THREADS)
enhance_cost_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ X, const float* __restrict__ tgt,
                        const int* __restrict__ perm, const float* __restrict__ dcost_b, int S, int64_t TF, int nonlin,
                        float* __restrict__ dlogits) {
    const int b = blockIdx.y;
    int inv[EC_MAXS];                        // inv[k] = the source whose target estimate k is compared with
#pragma unroll
    for (int s = 0; s < EC_MAXS; ++s) inv[s] = 0;
    for (int s = 0; s < S; ++s) inv[perm[b * S + s]] = s;
    const float g = dcost_b[b];
    for (int64_t j = blockIdx.x * (int64_t)EC_THREADS + threadIdx.x; j < TF; j += (int64_t)gridDim.x * EC_THREADS) {
        float z[EC_MAXS], m[EC_MAXS], dm[EC_MAXS];
#pragma unroll
        for (int s = 0; s < EC_MAXS; ++s) if (s < S) z[s] = logits[((size_t)b * S + s) * TF + j];
        ec_masks(z, S, nonlin, m);
        const float x = X[(size_t)b * TF + j];
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < EC_MAXS; ++k)
            if (k < S) {
                const float t = tgt[((size_t)b * TF + j) * S + inv[k]];
                dm[k] = g * (-2.f) * (t - m[k] * x) * x;
                dot = fmaf(m[k], dm[k], dot);
            }
#pragma unroll
        for (int k = 0; k < EC_MAXS; ++k)
            if (k < S) {
                float dz;
                if (nonlin == 0) dz = m[k] * (dm[k] - dot);
                else if (nonlin == 1) dz = dm[k] * (1.f - m[k] * m[k]);
                else dz = dm[k];
                dlogits[((size_t)b * S + k) * TF + j] = dz;
            }
    }
}

int ec_chunks(int B, int64_t TF) {
    const int64_t tiles = (TF + EC_THREADS - 1) / EC_THREADS;
    const int64_t want = std::max<int64_t>(1, (2 * kNumSMs + B - 1) / B);
    return (int)std::min(tiles, want);
}

}  // namespace
}  // namespace amss

using namespace amss;

extern "C" size_t amss_enhance_cost_workspace_bytes(int B, int64_t TF, int S) {
    return (size_t)B * ec_chunks(B, TF) * S * S * 4 + 256;
}

extern "C" int amss_enhance_cost_table(const float* logits, const float* X_input, const float* X_non_mix, int B, int S,
                                       int64_t TF, int nonlinearity, float* masks, float* table, void* workspace,
                                       size_t workspace_bytes, void* stream) {
    AMSS_REQUIRE(logits && X_input && X_non_mix && table && workspace, "enhance_cost_table: null pointer");
    AMSS_REQUIRE(B > 0 && S >= 1 && S <= EC_MAXS && TF > 0 && nonlinearity >= 0 && nonlinearity <= 2, "enhance_cost_table: bad arguments");
    if (workspace_bytes < amss_enhance_cost_workspace_bytes(B, TF, S)) { set_error("enhance_cost_table: workspace too small"); return AMSS_ERR_WORKSPACE; }
    const int chunks = ec_chunks(B, TF);
    AMSS_LAUNCH(enhance_table_kernel, dim3(chunks, B), EC_THREADS, 0, stream, logits, X_input, X_non_mix, S, TF, nonlinearity,
                masks, (float*)workspace);
    AMSS_LAUNCH(enhance_table_finish_kernel, (B * S * S + 127) / 128, 128, 0, stream, (const float*)workspace, B, chunks, S * S, table);
    return AMSS_OK;
}

extern "C" int amss_enhance_cost_bwd(const float* logits, const float* X_input, const float* X_non_mix, const int32_t* perm,
                                     const float* dcost_b, int B, int S, int64_t TF, int nonlinearity, float* dlogits,
                                     void* stream) {
    AMSS_REQUIRE(logits && X_input && X_non_mix && perm && dcost_b && dlogits, "enhance_cost_bwd: null pointer");
    AMSS_REQUIRE(B > 0 && S >= 1 && S <= EC_MAXS && TF > 0 && nonlinearity >= 0 && nonlinearity <= 2, "enhance_cost_bwd: bad arguments");
    AMSS_LAUNCH(enhance_cost_bwd_kernel, dim3(ec_chunks(B, TF), B), EC_THREADS, 0, stream, logits, X_input, X_non_mix, perm,
                dcost_b, S, TF, nonlinearity, dlogits);
    return AMSS_OK;
}
