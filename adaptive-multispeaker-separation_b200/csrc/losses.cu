// Loss kernels of the separator: DPCL affinity loss (models/dpcl.py:41-86), L41 sigmoid-dot loss
// (models/L41.py:150-178), per-group L2 normalisation (utils/ops.py:318-324), the plugged-mode
// label arg-max (models/network.py:369-378) and the waveform statistics of the Adapt pretraining
// cost (models/adapt.py:323-330, models/network.py:196-221).
// All reductions are two-level (per-CTA partials, fixed-order finalize): deterministic.
#include "common.cuh"
#include <algorithm>

namespace amss {
namespace {

constexpr int LS_THREADS = 256;
constexpr int LS_MAXS = 4;

// ---- DPCL forward --------------------------------------------------------------------------
// Y is one-hot, so with N_s = #bins of speaker s and w_i = N_{l_i}^{-1/2}:
//   V^T D V = sum_i w_i v_i v_i^T                                   (E x E, symmetric)
//   V^T D Y[:, s] = N_s^{-1/2} sum_{i in s} v_i                      (E)
//   Y^T D Y = diag(sqrt(N_s))
// Pass 1 counts N_s; pass 2 accumulates the weighted Gram matrix (upper-triangular 4x4 register
// blocks, LS_PT points staged per tile, point groups spread over the CTA) and the per-speaker
// column sums: part[b][chunk][E*E + S*E]; pass 3 reduces the chunks in a fixed order.
constexpr int LS_PT = 256;      // points per tile

__global__ void dpcl_count_kernel(const uint8_t* __restrict__ labels, int64_t TF, int S, float* __restrict__ counts) {
    __shared__ float red[32];
    const int b = blockIdx.x;
    int c[LS_MAXS] = {0, 0, 0, 0};
    const uint8_t* lb = labels + (size_t)b * TF;
    int64_t i0 = 0;
    if ((reinterpret_cast<uintptr_t>(lb) & 15) == 0) {    // 16 labels per load, byte-wise compare + popcount
        const uint4* l4 = reinterpret_cast<const uint4*>(lb);
        const int64_t n4 = TF / 16;
        for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
            const uint4 v = __ldg(l4 + i);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int s = 0; s < LS_MAXS; ++s) c[s] += __popc(__vcmpeq4(w[k], 0x01010101u * s)) >> 3;
        }
        i0 = n4 * 16;
    }
    for (int64_t i = i0 + threadIdx.x; i < TF; i += blockDim.x) {
        const int l = lb[i];
#pragma unroll
        for (int s = 0; s < LS_MAXS; ++s) c[s] += (l == s);
    }
    for (int s = 0; s < LS_MAXS; ++s) {
        const float v = block_sum((float)c[s], red);      // exact: counts < 2^24
        if (threadIdx.x == 0) counts[b * LS_MAXS + s] = s < S ? v : 0.f;
    }
}

// Weighted labels (--function_mask, models/network.py:381-389): Y[i,:] = u_i e_{l_i}, so (Y Y^T 1)_i = u_i U_{l_i} with
// U_s = sum_{i in s} u_i, and
//   V^T D V = sum_i (u_i U_{l_i})^{-1/2} v_i v_i^T,   V^T D Y[:, s] = U_s^{-1/2} sum_{i in s} sqrt(u_i) v_i,
//   Y^T D Y = diag(U_s^{-1/2} sum_{i in s} u_i^{3/2}).
// counts[b][s] = U_s, counts[B*LS_MAXS + b*LS_MAXS + s] = sum u^{3/2}; strided per-thread sums + block_sum: fixed order.
__global__ void dpcl_wcount_kernel(const uint8_t* __restrict__ labels, const float* __restrict__ weights, int64_t TF,
                                   int S, int B, float* __restrict__ counts) {
    __shared__ float red[32];
    const int b = blockIdx.x;
    float c[LS_MAXS] = {0.f, 0.f, 0.f, 0.f}, c15[LS_MAXS] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t i = threadIdx.x; i < TF; i += blockDim.x) {
        const int l = labels[(size_t)b * TF + i];
        const float u = weights[(size_t)b * TF + i], u15 = u * sqrtf(u);
#pragma unroll
        for (int s = 0; s < LS_MAXS; ++s) { c[s] += (l == s) ? u : 0.f; c15[s] += (l == s) ? u15 : 0.f; }
    }
    for (int s = 0; s < LS_MAXS; ++s) {
        const float v = block_sum(c[s], red);
        const float v15 = block_sum(c15[s], red);
        if (threadIdx.x == 0) {
            counts[b * LS_MAXS + s] = s < S ? v : 0.f;
            counts[(B + b) * LS_MAXS + s] = s < S ? v15 : 0.f;
        }
    }
}

__global__ void __launch_bounds__(LS_THREADS)
dpcl_gram_kernel(const float* __restrict__ V, const uint8_t* __restrict__ labels, const float* __restrict__ counts,
                 const float* __restrict__ weights, int64_t TF, int E, int S, float* __restrict__ part) {
    extern __shared__ __align__(16) unsigned char ls_smem[];
    const int EP = (E + 3) & ~3;
    float* xs = reinterpret_cast<float*>(ls_smem);              // [LS_PT][EP]   (reused for the final reduction)
    float* wsm = xs + LS_PT * EP;                               // [LS_PT] weights of the Gram term
    float* usm = wsm + LS_PT;                                   // [LS_PT] sqrt(u) of the column sums (1 without label weights)
    uint8_t* ls = reinterpret_cast<uint8_t*>(usm + LS_PT);      // [LS_PT] labels
    const int b = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x, tid = threadIdx.x;
    const int nb = EP / 4, nt = nb * (nb + 1) / 2;
    int G = LS_THREADS / nt; if (G > 8) G = 8; if (G < 1) G = 1;
    const bool active = tid < nt * G;
    const int blk = tid % nt, gq = tid / nt;
    int bi = 0, bj = 0;
    { int r = blk; while (r >= nb - bi) { r -= nb - bi; ++bi; } bj = bi + r; }   // upper triangle, row-major
    float wS[LS_MAXS];
#pragma unroll
    for (int s = 0; s < LS_MAXS; ++s) { const float n = counts[b * LS_MAXS + s]; wS[s] = n > 0.f ? rsqrtf(n) : 0.f; }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float macc[LS_MAXS] = {0.f, 0.f, 0.f, 0.f};
    const int me = tid % EP, mg = tid / EP;                     // column-sum role: column me, point group mg (< 4)
    const int64_t ntiles = (TF + LS_PT - 1) / LS_PT;
    for (int64_t tile = chunk; tile < ntiles; tile += chunks) {
        const int64_t p0 = tile * LS_PT;
        const int np = (int)((TF - p0) < LS_PT ? (TF - p0) : LS_PT);
        const float* src = V + ((size_t)b * TF + p0) * E;
        __syncthreads();
        if (EP == E) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
            float4* d4 = reinterpret_cast<float4*>(xs);
            for (int i = tid; i < np * (E / 4); i += LS_THREADS) d4[i] = __ldg(s4 + i);
        } else {
            for (int i = tid; i < np * EP; i += LS_THREADS) { const int p = i / EP, e = i - p * EP; xs[i] = e < E ? src[p * E + e] : 0.f; }
        }
        for (int i = tid; i < np; i += LS_THREADS) {
            const int l = labels[(size_t)b * TF + p0 + i];
            ls[i] = (uint8_t)l;
            const float wl = l == 0 ? wS[0] : (l == 1 ? wS[1] : (l == 2 ? wS[2] : wS[3]));
            if (weights) {
                const float u = weights[(size_t)b * TF + p0 + i];
                wsm[i] = wl * rsqrtf(u);                        // u == 0: inf, as 1/sqrt(0) in the reference
                usm[i] = sqrtf(u);
            } else {
                wsm[i] = wl;
                usm[i] = 1.f;
            }
        }
        __syncthreads();
        if (active) {
            for (int p = gq; p < np; p += G) {
                const float w = wsm[p];
                const float4 a = *reinterpret_cast<const float4*>(xs + p * EP + bi * 4);
                const float4 c = *reinterpret_cast<const float4*>(xs + p * EP + bj * 4);
                const float av[4] = {w * a.x, w * a.y, w * a.z, w * a.w};
                const float cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], cv[j], acc[i][j]);
            }
        }
        if (mg < 4) {
            for (int p = mg; p < np; p += 4) {
                const int l = ls[p];
                const float v = xs[p * EP + me] * usm[p];
#pragma unroll
                for (int s = 0; s < LS_MAXS; ++s) macc[s] += (l == s) ? v : 0.f;
            }
        }
    }
    // fixed-order reduction over the point groups, then one partial per (b, chunk)
    __syncthreads();
    float* red = xs;                                            // [G][nt][16] then [4][LS_MAXS][EP]
    float* mred = xs + 8 * nt * 16;
    if (active) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) red[(gq * nt + blk) * 16 + i * 4 + j] = acc[i][j];
    }
    if (mg < 4) {
#pragma unroll
        for (int s = 0; s < LS_MAXS; ++s) mred[(mg * LS_MAXS + s) * EP + me] = macc[s];
    }
    __syncthreads();
    float* dst = part + ((size_t)b * chunks + chunk) * ((size_t)E * E + (size_t)S * E);
    if (tid < nt) {
        int ri = 0, rj = 0;
        { int r = tid; while (r >= nb - ri) { r -= nb - ri; ++ri; } rj = ri + r; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float a = 0.f;
                for (int g = 0; g < G; ++g) a += red[(g * nt + tid) * 16 + i * 4 + j];
                const int ei = ri * 4 + i, ej = rj * 4 + j;
                if (ei < E && ej < E) { dst[ei * E + ej] = a; dst[ej * E + ei] = a; }
            }
    }
    for (int i = tid; i < S * E; i += LS_THREADS) {
        const int s = i / E, e = i - s * E;
        dst[E * E + i] = mred[(0 * LS_MAXS + s) * EP + e] + mred[(1 * LS_MAXS + s) * EP + e] +
                         mred[(2 * LS_MAXS + s) * EP + e] + mred[(3 * LS_MAXS + s) * EP + e];
    }
}

// One CTA per batch row: reduce the partials, form the three Frobenius norms, keep what the
// backward needs:  stats[b] = { An[E*E] = 2*A/||A||, Bn[S*E] = 2*Bm[:,s]/||Bm||, dinv[S], loss_b }.
// c15 != NULL (weighted labels): ||Y^T D Y||_F^2 = sum_s (c15_s)^2 / U_s.
__global__ void dpcl_finalize_kernel(const float* __restrict__ part, const float* __restrict__ counts,
                                     const float* __restrict__ c15, int chunks, int E, int S, float* __restrict__ stats) {
    __shared__ float red[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int stride = E * E + S * E;
    const int sstride = E * E + S * E + S + 1;
    float* out = stats + (size_t)b * sstride;
    float sa = 0.f, sb = 0.f;
    for (int i = tid; i < stride; i += blockDim.x) {
        float a = 0.f;
        for (int ch = 0; ch < chunks; ++ch) a += part[((size_t)b * chunks + ch) * stride + i];
        if (i >= E * E) {
            const float n = counts[b * LS_MAXS + (i - E * E) / E];
            a = n > 0.f ? a / sqrtf(n) : 0.f;
            sb = fmaf(a, a, sb);
        } else {
            sa = fmaf(a, a, sa);
        }
        out[i] = a;
    }
    sa = block_sum(sa, red);
    sb = block_sum(sb, red);
    float sc = 0.f;
    for (int s = 0; s < S; ++s) {                                  // ||diag(sqrt(N_s))||_F^2 = sum N_s
        const float n = counts[b * LS_MAXS + s];
        if (c15) { const float c = c15[b * LS_MAXS + s]; sc += n > 0.f ? c * c / n : 0.f; }
        else sc += n;
    }
    const float na = sqrtf(sa), nbm = sqrtf(sb), nc = sqrtf(sc);
    __syncthreads();
    for (int i = tid; i < E * E; i += blockDim.x) out[i] = 2.f * out[i] / na;
    for (int i = tid; i < S * E; i += blockDim.x) out[E * E + i] = 2.f * out[E * E + i] / nbm;
    if (tid < S) { const float n = counts[b * LS_MAXS + tid]; out[E * E + S * E + tid] = n > 0.f ? 1.f / sqrtf(n) : 0.f; }
    if (tid == 0) out[E * E + S * E + S] = na - 2.f * nbm + nc;
}

__global__ void mean_of_stat_kernel(const float* __restrict__ stats, int B, int sstride, int off,
                                    float* __restrict__ loss) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float a = 0.f;
        for (int b = 0; b < B; ++b) a += stats[(size_t)b * sstride + off];
        loss[0] = a / (float)B;
    }
}

// dV_i = (dloss/B) * D_i * ( An v_i - Bn[:, l_i] ),  An = 2A/||A||, Bn = 2Bm/||Bm||   (A symmetric)
// Register-blocked [256 points x E] x [E x E] product: thread = 4 points x OB outputs; the points are
// staged transposed (point-contiguous) so the 4 point values are one LDS.128, the An row is a
// warp-uniform broadcast; results go back through shared memory for coalesced stores.
constexpr int LS_BP = 256, LS_BPT = 260, LS_BWD_THREADS = 320;
// EC > 0: embedding size known at compile time (E = 40, the reference default): 5 output blocks of 8 -> the An row
// block is two LDS.128 per inner step (FFMA-bound), constant divisions, full unrolling.  Generic E: 4 blocks of OB.
// Persistent: CTAs walk the flat (mixture, tile) list, so the grid is exactly the resident slots (no wave tail).
// inv_norm != NULL fuses the backward of tf.nn.l2_normalize (utils/ops.py:323-324) into the same pass:
//   out_i = inv_i * (dV_i - v_i <v_i, dV_i>)   (dz, the gradient w.r.t. the un-normalised embeddings)
// so dV is never written and v is read once.
template <int OB, int EC>
__global__ void __launch_bounds__(LS_BWD_THREADS)
dpcl_bwd_kernel(const float* __restrict__ V, const uint8_t* __restrict__ labels, const float* __restrict__ dloss,
                const float* __restrict__ stats, const float* __restrict__ inv_norm, const float* __restrict__ weights,
                int B, int64_t TF, int Ert, int S, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char ls_smem[];
    const int E = EC > 0 ? EC : Ert;
    const int EO = E | 1;                                       // odd row pitch of the result staging
    float* xt = reinterpret_cast<float*>(ls_smem);              // [E][LS_BPT]  transposed points
    float* os = xt + E * LS_BPT;                                // [LS_BP][EO]  results
    float* An = os + LS_BP * EO;                                // [E][E]
    float* Bn = An + E * E;                                     // [S][E]
    float* dinv = Bn + S * E;                                   // [S]
    const int tid = threadIdx.x;
    const int sstride = E * E + S * E + S + 1;
    const float gscale = dloss[0] / (float)B;
    constexpr int NOB = EC > 0 ? EC / OB : 4;                   // output blocks per point group
    const int OBr = EC > 0 ? OB : (E + 3) / 4;                  // outputs per thread (<= OB)
    const int ob = tid % NOB, pg = tid / NOB;                   // 64 point groups of 4 points
    const bool worker = pg < LS_BP / 4;
    const int o0 = ob * OBr;
    const int64_t ntiles = (TF + LS_BP - 1) / LS_BP;
    int bcur = -1;
    for (int64_t w = blockIdx.x; w < (int64_t)B * ntiles; w += gridDim.x) {
        const int b = (int)(w / ntiles);
        const int64_t p0 = (w - (int64_t)b * ntiles) * LS_BP;
        const int np = (int)((TF - p0) < LS_BP ? (TF - p0) : LS_BP);
        const float* src = V + ((size_t)b * TF + p0) * E;
        __syncthreads();
        if (b != bcur) {
            const float* st = stats + (size_t)b * sstride;
            for (int i = tid; i < E * E + S * E + S; i += LS_BWD_THREADS) An[i] = st[i];
            bcur = b;
        }
        for (int i = tid; i < LS_BP * E; i += LS_BWD_THREADS) {
            const int p = i / E, e = i - p * E;
            xt[e * LS_BPT + p] = p < np ? __ldg(src + i) : 0.f;
        }
        __syncthreads();
        if (worker) {
            float acc[4][OB];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < OB; ++j) acc[i][j] = 0.f;
#pragma unroll 8
            for (int e2 = 0; e2 < E; ++e2) {
                const float4 v = *reinterpret_cast<const float4*>(xt + e2 * LS_BPT + pg * 4);
                const float* ar = An + e2 * E + o0;              // An symmetric: column block of row e2
                float a[OB];
                if (EC > 0 && OB == 8) {
                    const float4 a0 = *reinterpret_cast<const float4*>(ar), a1 = *reinterpret_cast<const float4*>(ar + 4);
                    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4 % OB] = a1.x; a[5 % OB] = a1.y; a[6 % OB] = a1.z; a[7 % OB] = a1.w;
                } else {
#pragma unroll
                    for (int j = 0; j < OB; ++j) a[j] = (j < OBr && o0 + j < E) ? ar[j] : 0.f;
                }
#pragma unroll
                for (int j = 0; j < OB; ++j) {
                    acc[0][j] = fmaf(a[j], v.x, acc[0][j]); acc[1][j] = fmaf(a[j], v.y, acc[1][j]);
                    acc[2][j] = fmaf(a[j], v.z, acc[2][j]); acc[3][j] = fmaf(a[j], v.w, acc[3][j]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int p = pg * 4 + i;
                if (p < np) {
                    const int l = labels[(size_t)b * TF + p0 + p];
                    const float d = gscale * dinv[l];
                    if (weights) {                              // weighted labels: D_i = dinv / sqrt(u), D_i u_i = dinv * sqrt(u)
                        const float u = weights[(size_t)b * TF + p0 + p], ru = rsqrtf(u), su = sqrtf(u);
#pragma unroll
                        for (int j = 0; j < OB; ++j)
                            if (EC > 0 || (j < OBr && o0 + j < E))
                                os[p * EO + o0 + j] = d * (acc[i][j] * ru - Bn[l * E + o0 + j] * su);
                    } else {
#pragma unroll
                        for (int j = 0; j < OB; ++j)
                            if (EC > 0 || (j < OBr && o0 + j < E)) os[p * EO + o0 + j] = d * (acc[i][j] - Bn[l * E + o0 + j]);
                    }
                }
            }
        }
        __syncthreads();
        if (inv_norm != nullptr) {                              // fused l2_normalize backward, one thread per point
            if (tid < np) {
                float inv = inv_norm[(size_t)b * TF + p0 + tid];
                float dot = 0.f;
#pragma unroll 8
                for (int e = 0; e < E; ++e) dot = fmaf(xt[e * LS_BPT + tid], os[tid * EO + e], dot);
                if (inv < 0.f) { inv = -inv; dot = 0.f; }       // clamped branch of l2_normalize: linear map
#pragma unroll 8
                for (int e = 0; e < E; ++e) os[tid * EO + e] = inv * (os[tid * EO + e] - xt[e * LS_BPT + tid] * dot);
            }
            __syncthreads();
        }
        float* dst = out + ((size_t)b * TF + p0) * E;
        for (int i = tid; i < np * E; i += LS_BWD_THREADS) { const int pp = i / E; dst[i] = os[pp * EO + (i - pp * E)]; }
    }
}

// ---- per-group L2 normalisation ----------------------------------------------------------------
__global__ void l2norm_fwd_kernel(const float* __restrict__ z, int64_t rows, int E, float* __restrict__ v,
                                  float* __restrict__ inv_norm) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nwarps) {
        const float* x = z + r * E;
        float ss = 0.f;
        for (int e = lane; e < E; e += 32) ss = fmaf(x[e], x[e], ss);
        ss = warp_sum(ss);
        const float inv = rsqrtf(fmaxf(ss, 1e-12f));
        for (int e = lane; e < E; e += 32) v[r * E + e] = x[e] * inv;
        if (inv_norm && lane == 0) inv_norm[r] = (ss >= 1e-12f) ? inv : -inv;   // sign marks the clamped branch
    }
}
// dz = inv * (dv - v * <v, dv>) ; in the clamped branch (sum z^2 < eps) the map is linear: dz = inv * dv
__global__ void l2norm_bwd_kernel(const float* __restrict__ v, const float* __restrict__ inv_norm,
                                  const float* __restrict__ dv, int64_t rows, int E, float* __restrict__ dz) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nwarps) {
        float dot = 0.f;
        for (int e = lane; e < E; e += 32) dot = fmaf(v[r * E + e], dv[r * E + e], dot);
        dot = warp_sum(dot);
        float inv = inv_norm[r];
        if (inv < 0.f) { inv = -inv; dot = 0.f; }
        for (int e = lane; e < E; e += 32) dz[r * E + e] = inv * (dv[r * E + e] - v[r * E + e] * dot);
    }
}

// Vectorised variants for E % 4 == 0, E <= 128: E/4 lanes per row (one float4 each), 32/(E/4) rows per
// warp per iteration -> fully coalesced 16-byte accesses; segmented shuffle reduction inside the row group.
__device__ __forceinline__ float group_sum(float s, int gl, int lpr, int leader) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float t = __shfl_down_sync(0xffffffffu, s, off);
        if (off < lpr && gl + off < lpr) s += t;
    }
    return __shfl_sync(0xffffffffu, s, leader);
}
__global__ void l2norm_fwd_vec_kernel(const float4* __restrict__ z, int64_t rows, int lpr, float4* __restrict__ v,
                                      float* __restrict__ inv_norm) {
    const int lane = threadIdx.x & 31, rpw = 32 / lpr;
    const int g = lane / lpr, gl = lane - g * lpr;
    const bool on = g < rpw;
    const int leader = on ? g * lpr : 0;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r0 = warp * rpw; r0 < rows; r0 += nwarps * rpw) {
        const int64_t r = r0 + g;
        const bool ok = on && r < rows;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) x = __ldcs(z + r * lpr + gl);
        const float ss = group_sum(x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w, gl, lpr, leader);
        const float inv = rsqrtf(fmaxf(ss, 1e-12f));
        if (ok) {
            v[r * lpr + gl] = make_float4(x.x * inv, x.y * inv, x.z * inv, x.w * inv);
            if (inv_norm && gl == 0) inv_norm[r] = (ss >= 1e-12f) ? inv : -inv;
        }
    }
}
__global__ void l2norm_bwd_vec_kernel(const float4* __restrict__ v, const float* __restrict__ inv_norm,
                                      const float4* __restrict__ dv, int64_t rows, int lpr, float4* __restrict__ dz) {
    const int lane = threadIdx.x & 31, rpw = 32 / lpr;
    const int g = lane / lpr, gl = lane - g * lpr;
    const bool on = g < rpw;
    const int leader = on ? g * lpr : 0;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r0 = warp * rpw; r0 < rows; r0 += nwarps * rpw) {
        const int64_t r = r0 + g;
        const bool ok = on && r < rows;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), d = a;
        float inv = 0.f;
        if (ok) { a = __ldcs(v + r * lpr + gl); d = __ldcs(dv + r * lpr + gl); inv = inv_norm[r]; }
        float dot = group_sum(a.x * d.x + a.y * d.y + a.z * d.z + a.w * d.w, gl, lpr, leader);
        if (inv < 0.f) { inv = -inv; dot = 0.f; }
        if (ok) dz[r * lpr + gl] = make_float4(inv * (d.x - a.x * dot), inv * (d.y - a.y * dot), inv * (d.z - a.z * dot),
                                               inv * (d.w - a.w * dot));
    }
}

// dbias[n] = sum_m dZ[m][n]: grid (ceil(N/32), chunks) -> partial[chunks][N], then a fixed-order sum
__global__ void colsum_partial_kernel(const float* __restrict__ dZ, int64_t M, int N, float* __restrict__ part) {
    __shared__ float tile[8][33];
    const int n = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ty = threadIdx.x >> 5;   // 8 row lanes
    float a = 0.f;
    if (n < N)
        for (int64_t m = blockIdx.y * 8 + ty; m < M; m += (int64_t)gridDim.y * 8) a += dZ[m * N + n];
    tile[ty][threadIdx.x & 31] = a;
    __syncthreads();
    if (ty == 0 && n < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += tile[i][threadIdx.x & 31];
        part[(size_t)blockIdx.y * N + n] = s;
    }
}
// same for a bf16 matrix (N even): a thread owns two adjacent columns, a warp reads 128 contiguous bytes of a row
__global__ void colsum_partial_bf16_kernel(const uint32_t* __restrict__ dZ2, int64_t M, int N2, float* __restrict__ part) {
    __shared__ float2 tile[8][33];
    const int n2 = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ty = threadIdx.x >> 5;
    float2 a = make_float2(0.f, 0.f);
    if (n2 < N2)
        for (int64_t m = blockIdx.y * 8 + ty; m < M; m += (int64_t)gridDim.y * 8) {
            const uint32_t w = __ldg(dZ2 + m * N2 + n2);
            a.x += __uint_as_float(w << 16);
            a.y += __uint_as_float(w & 0xFFFF0000u);
        }
    tile[ty][threadIdx.x & 31] = a;
    __syncthreads();
    if (ty == 0 && n2 < N2) {
        float2 s = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s.x += tile[i][threadIdx.x & 31].x; s.y += tile[i][threadIdx.x & 31].y; }
        part[(size_t)blockIdx.y * (2 * N2) + 2 * n2] = s.x;
        part[(size_t)blockIdx.y * (2 * N2) + 2 * n2 + 1] = s.y;
    }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, int chunks, int N, float* __restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) {
        float s = 0.f;
        for (int c = 0; c < chunks; ++c) s += part[(size_t)c * N + n];
        out[n] = s;
    }
}

// ---- L41 ---------------------------------------------------------------------------------------
// cost = mean_{b,i,s} softplus(-y * <spk[b,s], emb[b,i]>), y = +1 if labels[b,i]==s else -1
__global__ void __launch_bounds__(LS_THREADS)
l41_fwd_kernel(const float* __restrict__ emb, const uint8_t* __restrict__ labels, const float* __restrict__ spk,
               const float* __restrict__ weights, int64_t TF, int E, int S, float* __restrict__ part) {
    extern __shared__ __align__(16) unsigned char ls_smem[];
    __shared__ float red[32];
    const int EP = E + 1;
    float* xs = reinterpret_cast<float*>(ls_smem);
    float* sp = xs + LS_THREADS * EP;
    const int b = blockIdx.y, tid = threadIdx.x;
    for (int i = tid; i < S * E; i += LS_THREADS) sp[i] = spk[(size_t)b * S * E + i];
    float acc = 0.f;
    const int64_t ntiles = (TF + LS_THREADS - 1) / LS_THREADS;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p0 = tile * LS_THREADS;
        const int np = (int)((TF - p0) < LS_THREADS ? (TF - p0) : LS_THREADS);
        const float* src = emb + ((size_t)b * TF + p0) * E;
        __syncthreads();
        for (int i = tid; i < np * E; i += LS_THREADS) { const int p = i / E, e = i - p * E; xs[p * EP + e] = src[i]; }
        __syncthreads();
        if (tid < np) {
            const int l = labels[(size_t)b * TF + p0 + tid];
            // y = +-1 (one_hot(argmax, S, 1, -1)) times the optional label weight (function_mask / silence_loss,
            // models/network.py:381-396)
            const float w = weights ? weights[(size_t)b * TF + p0 + tid] : 1.f;
            for (int s = 0; s < S; ++s) {
                float d = 0.f;
                for (int e = 0; e < E; ++e) d = fmaf(sp[s * E + e], xs[tid * EP + e], d);
                const float x = ((l == s) ? d : -d) * w;
                // -log(sigmoid(x)) = softplus(-x)
                acc += (x > 0.f) ? log1pf(expf(-x)) : (-x + log1pf(expf(x)));
            }
        }
    }
    acc = block_sum(acc, red);
    if (tid == 0) part[(size_t)b * gridDim.x + blockIdx.x] = acc;
}
__global__ void l41_fwd_final_kernel(const float* __restrict__ part, int n, float denom, float* __restrict__ loss) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float a = 0.f;
        for (int i = 0; i < n; ++i) a += part[i];
        loss[0] = a / denom;
    }
}
// demb_i = sum_s c_is spk_s ;  dspk_s = sum_i c_is emb_i ;  c_is = -y * sigmoid(-y*dot) * dloss / (B*TF*S)
__global__ void __launch_bounds__(LS_THREADS)
l41_bwd_kernel(const float* __restrict__ emb, const uint8_t* __restrict__ labels, const float* __restrict__ spk,
               const float* __restrict__ weights, const float* __restrict__ dloss, int B, int64_t TF, int E, int S,
               float* __restrict__ demb, float* __restrict__ dspk_part) {
    extern __shared__ __align__(16) unsigned char ls_smem[];
    const int EP = E + 1;
    float* xs = reinterpret_cast<float*>(ls_smem);              // [LS_THREADS][EP]
    float* sp = xs + LS_THREADS * EP;                           // [S][E]
    float* cs = sp + S * E;                                     // [LS_THREADS][S] coefficients
    const int b = blockIdx.y, tid = threadIdx.x;
    for (int i = tid; i < S * E; i += LS_THREADS) sp[i] = spk[(size_t)b * S * E + i];
    const float g = dloss[0] / ((float)B * (float)TF * (float)S);
    float dacc[LS_MAXS];
#pragma unroll
    for (int s = 0; s < LS_MAXS; ++s) dacc[s] = 0.f;
    const int64_t ntiles = (TF + LS_THREADS - 1) / LS_THREADS;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p0 = tile * LS_THREADS;
        const int np = (int)((TF - p0) < LS_THREADS ? (TF - p0) : LS_THREADS);
        const float* src = emb + ((size_t)b * TF + p0) * E;
        __syncthreads();
        for (int i = tid; i < np * E; i += LS_THREADS) { const int p = i / E, e = i - p * E; xs[p * EP + e] = src[i]; }
        __syncthreads();
        if (tid < np) {
            const int l = labels[(size_t)b * TF + p0 + tid];
            const float w = weights ? weights[(size_t)b * TF + p0 + tid] : 1.f;
            for (int s = 0; s < S; ++s) {
                float d = 0.f;
                for (int e = 0; e < E; ++e) d = fmaf(sp[s * E + e], xs[tid * EP + e], d);
                const float y = ((l == s) ? 1.f : -1.f) * w;
                cs[tid * S + s] = -y * g / (1.f + expf(y * d));
            }
        }
        __syncthreads();
        // dspk: thread e < E accumulates over the tile's points (fixed order)
        if (tid < E) {
            for (int p = 0; p < np; ++p) {
                const float v = xs[p * EP + tid];
#pragma unroll
                for (int s = 0; s < LS_MAXS; ++s) if (s < S) dacc[s] = fmaf(cs[p * S + s], v, dacc[s]);
            }
        }
        // demb: coalesced write
        for (int i = tid; i < np * E; i += LS_THREADS) {
            const int p = i / E, e = i - p * E;
            float a = 0.f;
            for (int s = 0; s < S; ++s) a = fmaf(cs[p * S + s], sp[s * E + e], a);
            demb[((size_t)b * TF + p0) * E + i] = a;
        }
    }
    if (tid < E) {
#pragma unroll
        for (int s = 0; s < LS_MAXS; ++s)
            if (s < S) dspk_part[(((size_t)b * gridDim.x + blockIdx.x) * S + s) * E + tid] = dacc[s];
    }
}
__global__ void l41_dspk_final_kernel(const float* __restrict__ part, int B, int chunks, int SE,
                                      float* __restrict__ dspk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * SE) {
        const int b = i / SE, j = i - b * SE;
        float a = 0.f;
        for (int c = 0; c < chunks; ++c) a += part[((size_t)b * chunks + c) * SE + j];
        dspk[i] = a;
    }
}

// labels[b][j] = argmax_s |front_y[B + b*S + s][j]|   (network.py:372-378; first index wins ties)
__global__ void plugged_labels_kernel(const float* __restrict__ y, int B, int S, int64_t TN,
                                      uint8_t* __restrict__ labels) {
    const int64_t n = (int64_t)B * TN;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / TN, j = i - b * TN;
        float best = -1.f;
        int bi = 0;
        for (int s = 0; s < S; ++s) {
            const float v = fabsf(y[((size_t)B + b * S + s) * TN + j]);
            if (v > best) { best = v; bi = s; }
        }
        labels[i] = (uint8_t)bi;
    }
}

// the same, four bins per thread (TN % 4 == 0, aligned buffers)
__global__ void plugged_labels_vec_kernel(const float4* __restrict__ y, int B, int S, int64_t TN4, uint32_t* __restrict__ labels) {
    const int64_t n = (int64_t)B * TN4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / TN4, j = i - b * TN4;
        float best[4] = {-1.f, -1.f, -1.f, -1.f};
        uint32_t bi[4] = {0, 0, 0, 0};
        for (int s = 0; s < S; ++s) {
            const float4 v = __ldg(y + ((size_t)B + b * S + s) * TN4 + j);
            const float a[4] = {fabsf(v.x), fabsf(v.y), fabsf(v.z), fabsf(v.w)};
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (a[q] > best[q]) { best[q] = a[q]; bi[q] = (uint32_t)s; }
        }
        labels[i] = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
    }
}

// stats[r] = { <t,t>, <a,a>, <t,a>, <t-a,t-a> } over L   (fixed-order two-level reduction)
__global__ void wave_stats_kernel(const float* __restrict__ tg, const float* __restrict__ ap, int64_t L, int ap_div,
                                  float* __restrict__ stats) {
    __shared__ float red[32];
    const int r = blockIdx.x;
    const size_t ar = (size_t)(r / ap_div);
    float tt = 0.f, aa = 0.f, ta = 0.f, ee = 0.f;
    for (int64_t i = threadIdx.x; i < L; i += blockDim.x) {
        const float t = tg[(size_t)r * L + i], a = ap[ar * L + i], d = t - a;
        tt = fmaf(t, t, tt); aa = fmaf(a, a, aa); ta = fmaf(t, a, ta); ee = fmaf(d, d, ee);
    }
    tt = block_sum(tt, red); aa = block_sum(aa, red); ta = block_sum(ta, red); ee = block_sum(ee, red);
    if (threadIdx.x == 0) { stats[r * 4 + 0] = tt; stats[r * 4 + 1] = aa; stats[r * 4 + 2] = ta; stats[r * 4 + 3] = ee; }
}

// d approx[r][i] = ca[r] * approx[r][i] + ct[r] * target[r][i],  ca = 2 d<a,a> + 2 d<t-a,t-a>,  ct = d<t,a> - 2 d<t-a,t-a>
// (the gradient of the four statistics w.r.t. the approximation; the targets are data)
__global__ void wave_stats_bwd_kernel(const float* __restrict__ tg, const float* __restrict__ ap, const float* __restrict__ dstats,
                                      int R, int64_t L, float* __restrict__ dap) {
    const int64_t n = (int64_t)R * L;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / L);
        const float d1 = dstats[r * 4 + 1], d2 = dstats[r * 4 + 2], d3 = dstats[r * 4 + 3];
        dap[i] = (2.f * d1 + 2.f * d3) * ap[i] + (d2 - 2.f * d3) * tg[i];
    }
}

int ls_chunks(int B, int64_t TF, int tile) {
    int64_t nt = (TF + tile - 1) / tile;
    int64_t want = (2 * kNumSMs + B - 1) / B;
    if (want < 1) want = 1;
    return (int)(nt < want ? nt : want);
}

}  // namespace
}  // namespace amss

using namespace amss;

extern "C" size_t amss_dpcl_workspace_bytes(int B, int64_t TF, int E, int S) {
    const size_t stride = (size_t)E * E + (size_t)S * E;
    const size_t sstride = (size_t)E * E + (size_t)S * E + S + 1;
    return align_up((size_t)B * sstride * 4, 256) + align_up((size_t)2 * B * LS_MAXS * 4, 256) +
           align_up((size_t)B * ls_chunks(B, TF, LS_PT) * stride * 4, 256);
}

namespace amss {
bool dpcl_gram_tc_supported(int E, int S);
int dpcl_gram_tc_chunks(int B);
int dpcl_gram_tc(const float* V, const uint8_t* labels, const float* counts, int B, int64_t TF, int E, int S, int chunks,
                 float* part, cudaStream_t st);
}  // namespace amss

static int dpcl_fwd_impl(const float* V, const uint8_t* labels, const float* weights, int B, int64_t TF, int E, int S,
                         int precision, float* loss, void* workspace, size_t workspace_bytes, void* stream) {
    AMSS_REQUIRE(V && labels && loss && workspace, "dpcl_loss_fwd: null pointer");
    AMSS_REQUIRE(S >= 1 && S <= LS_MAXS, "dpcl_loss_fwd: S=%d outside [1,%d]", S, LS_MAXS);
    AMSS_REQUIRE(E >= 1 && E <= 64, "dpcl_loss_fwd: E=%d outside [1,64]", E);
    if (workspace_bytes < amss_dpcl_workspace_bytes(B, TF, E, S)) { set_error("dpcl_loss_fwd: workspace too small"); return AMSS_ERR_WORKSPACE; }
    const int sstride = E * E + S * E + S + 1;
    float* stats = (float*)workspace;
    float* counts = (float*)((char*)workspace + align_up((size_t)B * sstride * 4, 256));
    float* part = (float*)((char*)counts + align_up((size_t)2 * B * LS_MAXS * 4, 256));
    if (weights) AMSS_LAUNCH(dpcl_wcount_kernel, B, 256, 0, stream, labels, weights, TF, S, B, counts);
    else AMSS_LAUNCH(dpcl_count_kernel, B, 256, 0, stream, labels, TF, S, counts);
    int chunks = ls_chunks(B, TF, LS_PT);
    const bool tc = !weights && precision == AMSS_PREC_BF16 && dpcl_gram_tc_supported(E, S) &&
                    (reinterpret_cast<uintptr_t>(V) & 15) == 0;
    if (tc) {
        chunks = std::min(chunks, dpcl_gram_tc_chunks(B));
        int rc = dpcl_gram_tc(V, labels, counts, B, TF, E, S, chunks, part, (cudaStream_t)stream);
        if (rc != AMSS_OK) return rc;
    } else {
        const int EP = (E + 3) & ~3, nb = EP / 4, nt = nb * (nb + 1) / 2;
        size_t smem1 = (size_t)LS_PT * EP * 4 + 2 * LS_PT * 4 + LS_PT;
        smem1 = std::max(smem1, ((size_t)8 * nt * 16 + (size_t)4 * LS_MAXS * EP) * 4);
        AMSS_CUDA(cudaFuncSetAttribute(dpcl_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        dim3 grid(chunks, B);
        AMSS_LAUNCH(dpcl_gram_kernel, grid, LS_THREADS, smem1, stream, V, labels, counts, weights, TF, E, S, part);
    }
    AMSS_LAUNCH(dpcl_finalize_kernel, B, 256, 0, stream, part, counts,
                weights ? (const float*)(counts + (size_t)B * LS_MAXS) : (const float*)nullptr, chunks, E, S, stats);
    AMSS_LAUNCH(mean_of_stat_kernel, 1, 32, 0, stream, stats, B, sstride, sstride - 1, loss);
    return AMSS_OK;
}

extern "C" int amss_dpcl_loss_fwd(const float* V, const uint8_t* labels, int B, int64_t TF, int E, int S, float* loss,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    return dpcl_fwd_impl(V, labels, nullptr, B, TF, E, S, AMSS_PREC_FP32, loss, workspace, workspace_bytes, stream);
}

extern "C" int amss_dpcl_loss_fwd_prec(const float* V, const uint8_t* labels, int B, int64_t TF, int E, int S,
                                       int precision, float* loss, void* workspace, size_t workspace_bytes,
                                       void* stream) {
    return dpcl_fwd_impl(V, labels, nullptr, B, TF, E, S, precision, loss, workspace, workspace_bytes, stream);
}

extern "C" int amss_dpcl_loss_weighted_fwd(const float* V, const uint8_t* labels, const float* weights, int B, int64_t TF,
                                           int E, int S, float* loss, void* workspace, size_t workspace_bytes,
                                           void* stream) {
    AMSS_REQUIRE(weights, "dpcl_loss_weighted_fwd: null weights");
    return dpcl_fwd_impl(V, labels, weights, B, TF, E, S, AMSS_PREC_FP32, loss, workspace, workspace_bytes, stream);
}

namespace {
int dpcl_bwd_launch(const float* V, const uint8_t* labels, const float* dloss, const float* inv_norm, const float* weights,
                    int B, int64_t TF, int E, int S, float* out, const void* workspace, void* stream) {
    const size_t smem = ((size_t)E * LS_BPT + (size_t)LS_BP * (E | 1) + (size_t)E * E + (size_t)S * E + S) * 4;
    const int64_t work = (int64_t)B * ((TF + LS_BP - 1) / LS_BP);
    const int grid = (int)std::min<int64_t>(work, 2 * kNumSMs);      // 2 resident CTAs per SM (89 KB of shared memory each)
    if (E == 40) {
        AMSS_CUDA(cudaFuncSetAttribute(dpcl_bwd_kernel<8, 40>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        AMSS_LAUNCH((dpcl_bwd_kernel<8, 40>), grid, LS_BWD_THREADS, smem, stream, V, labels, dloss, (const float*)workspace,
                    inv_norm, weights, B, TF, E, S, out);
    } else {
        AMSS_CUDA(cudaFuncSetAttribute(dpcl_bwd_kernel<16, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        AMSS_LAUNCH((dpcl_bwd_kernel<16, 0>), grid, LS_BWD_THREADS, smem, stream, V, labels, dloss, (const float*)workspace,
                    inv_norm, weights, B, TF, E, S, out);
    }
    return AMSS_OK;
}
}  // namespace

extern "C" int amss_dpcl_loss_bwd(const float* V, const uint8_t* labels, const float* dloss, int B, int64_t TF, int E,
                                  int S, float* dV, const void* workspace, void* stream) {
    AMSS_REQUIRE(V && labels && dloss && dV && workspace, "dpcl_loss_bwd: null pointer");
    AMSS_REQUIRE(S >= 1 && S <= LS_MAXS, "dpcl_loss_bwd: S out of range");
    AMSS_REQUIRE(E >= 1 && E <= 64, "dpcl_loss_bwd: E=%d outside [1,64]", E);
    return dpcl_bwd_launch(V, labels, dloss, nullptr, nullptr, B, TF, E, S, dV, workspace, stream);
}

extern "C" int amss_dpcl_loss_weighted_bwd(const float* V, const uint8_t* labels, const float* weights, const float* dloss,
                                           const float* inv_norm, int B, int64_t TF, int E, int S, float* dV,
                                           const void* workspace, void* stream) {
    AMSS_REQUIRE(V && labels && weights && dloss && dV && workspace, "dpcl_loss_weighted_bwd: null pointer");
    AMSS_REQUIRE(S >= 1 && S <= LS_MAXS, "dpcl_loss_weighted_bwd: S out of range");
    AMSS_REQUIRE(E >= 1 && E <= 64, "dpcl_loss_weighted_bwd: E=%d outside [1,64]", E);
    return dpcl_bwd_launch(V, labels, dloss, inv_norm, weights, B, TF, E, S, dV, workspace, stream);
}

namespace amss {
bool dpcl_bwd_tc_supported(int E, int S);
int dpcl_bwd_tc(const float* V, const uint8_t* labels, const float* dloss, const float* stats, const float* inv_norm, int B,
                int64_t TF, int E, int S, float* dz, uint16_t* dz_bf16, cudaStream_t st);
}  // namespace amss

extern "C" int amss_dpcl_loss_bwd_normalized(const float* V, const uint8_t* labels, const float* dloss,
                                             const float* inv_norm, int B, int64_t TF, int E, int S, int precision,
                                             float* dz, const void* workspace, void* stream) {
    AMSS_REQUIRE(V && labels && dloss && inv_norm && dz && workspace, "dpcl_loss_bwd_normalized: null pointer");
    AMSS_REQUIRE(S >= 1 && S <= LS_MAXS, "dpcl_loss_bwd_normalized: S out of range");
    AMSS_REQUIRE(E >= 1 && E <= 64, "dpcl_loss_bwd_normalized: E=%d outside [1,64]", E);
    if (precision == AMSS_PREC_BF16 && dpcl_bwd_tc_supported(E, S) &&
        ((reinterpret_cast<uintptr_t>(V) | reinterpret_cast<uintptr_t>(dz)) & 15) == 0)
        return dpcl_bwd_tc(V, labels, dloss, (const float*)workspace, inv_norm, B, TF, E, S, dz, nullptr, (cudaStream_t)stream);
    return dpcl_bwd_launch(V, labels, dloss, inv_norm, nullptr, B, TF, E, S, dz, workspace, stream);
}

extern "C" int amss_dpcl_loss_bwd_normalized_bf16(const float* V, const uint8_t* labels, const float* dloss,
                                                  const float* inv_norm, int B, int64_t TF, int E, int S,
                                                  uint16_t* dz_bf16, const void* workspace, void* stream) {
    AMSS_REQUIRE(V && labels && dloss && inv_norm && dz_bf16 && workspace, "dpcl_loss_bwd_normalized_bf16: null pointer");
    if (!(dpcl_bwd_tc_supported(E, S) && E % 8 == 0) ||
        ((reinterpret_cast<uintptr_t>(V) | reinterpret_cast<uintptr_t>(dz_bf16)) & 15) != 0) {
        set_error("dpcl_loss_bwd_normalized_bf16: needs E %% 8 == 0, 8 <= E <= 64, 16-byte aligned buffers (E=%d S=%d)", E, S);
        return AMSS_ERR_UNSUPPORTED;
    }
    return dpcl_bwd_tc(V, labels, dloss, (const float*)workspace, inv_norm, B, TF, E, S, nullptr, dz_bf16, (cudaStream_t)stream);
}

extern "C" int amss_l2norm_fwd(const float* z, int64_t rows, int E, float* v, float* inv_norm, void* stream) {
    AMSS_REQUIRE(z && v && rows > 0 && E > 0, "l2norm_fwd: bad arguments");
    if (E % 4 == 0 && E <= 128 && ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(v)) & 15) == 0) {
        AMSS_LAUNCH(l2norm_fwd_vec_kernel, 16 * kNumSMs, 256, 0, stream, (const float4*)z, rows, E / 4, (float4*)v, inv_norm);
        return AMSS_OK;
    }
    AMSS_LAUNCH(l2norm_fwd_kernel, 8 * kNumSMs, 256, 0, stream, z, rows, E, v, inv_norm);
    return AMSS_OK;
}
extern "C" int amss_l2norm_bwd(const float* v, const float* inv_norm, const float* dv, int64_t rows, int E, float* dz,
                               void* stream) {
    AMSS_REQUIRE(v && inv_norm && dv && dz, "l2norm_bwd: null pointer");
    if (E % 4 == 0 && E <= 128 &&
        ((reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(dv) | reinterpret_cast<uintptr_t>(dz)) & 15) == 0) {
        AMSS_LAUNCH(l2norm_bwd_vec_kernel, 16 * kNumSMs, 256, 0, stream, (const float4*)v, inv_norm, (const float4*)dv, rows,
                    E / 4, (float4*)dz);
        return AMSS_OK;
    }
    AMSS_LAUNCH(l2norm_bwd_kernel, 8 * kNumSMs, 256, 0, stream, v, inv_norm, dv, rows, E, dz);
    return AMSS_OK;
}

extern "C" size_t amss_colsum_workspace_bytes(int64_t M, int N) {
    (void)M;
    return (size_t)64 * N * 4;
}
extern "C" int amss_colsum(const float* dZ, int64_t M, int N, float* dbias, void* workspace, size_t workspace_bytes,
                           void* stream) {
    AMSS_REQUIRE(dZ && dbias && workspace, "colsum: null pointer");
    int chunks = (int)std::min<int64_t>(64, (M + 7) / 8);
    if (workspace_bytes < (size_t)chunks * N * 4) { set_error("colsum: workspace too small"); return AMSS_ERR_WORKSPACE; }
    dim3 grid((N + 31) / 32, chunks);
    AMSS_LAUNCH(colsum_partial_kernel, grid, 256, 0, stream, dZ, M, N, (float*)workspace);
    AMSS_LAUNCH(colsum_final_kernel, (N + 255) / 256, 256, 0, stream, (const float*)workspace, chunks, N, dbias);
    return AMSS_OK;
}

extern "C" int amss_colsum_bf16(const uint16_t* dZ, int64_t M, int N, float* dbias, void* workspace, size_t workspace_bytes,
                                void* stream) {
    AMSS_REQUIRE(dZ && dbias && workspace, "colsum_bf16: null pointer");
    AMSS_REQUIRE(N % 2 == 0 && (reinterpret_cast<uintptr_t>(dZ) & 3) == 0, "colsum_bf16: N must be even and dZ 4-byte aligned");
    int chunks = (int)std::min<int64_t>(64, (M + 7) / 8);
    if (workspace_bytes < (size_t)chunks * N * 4) { set_error("colsum_bf16: workspace too small"); return AMSS_ERR_WORKSPACE; }
    dim3 grid((N / 2 + 31) / 32, chunks);
    AMSS_LAUNCH(colsum_partial_bf16_kernel, grid, 256, 0, stream, (const uint32_t*)dZ, M, N / 2, (float*)workspace);
    AMSS_LAUNCH(colsum_final_kernel, (N + 255) / 256, 256, 0, stream, (const float*)workspace, chunks, N, dbias);
    return AMSS_OK;
}

extern "C" size_t amss_l41_workspace_bytes(int B, int64_t TF, int E, int S) {
    const int chunks = ls_chunks(B, TF, LS_THREADS);
    return align_up((size_t)B * chunks * 4, 256) + align_up((size_t)B * chunks * S * E * 4, 256);
}
extern "C" int amss_l41_loss_fwd(const float* emb, const uint8_t* labels, const float* spk, const float* weights, int B,
                                 int64_t TF, int E, int S, float* loss, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    AMSS_REQUIRE(emb && labels && spk && loss && workspace, "l41_loss_fwd: null pointer");
    AMSS_REQUIRE(S >= 1 && S <= LS_MAXS, "l41_loss_fwd: S out of range");
    if (workspace_bytes < amss_l41_workspace_bytes(B, TF, E, S)) { set_error("l41_loss_fwd: workspace too small"); return AMSS_ERR_WORKSPACE; }
    const int chunks = ls_chunks(B, TF, LS_THREADS);
    const size_t smem = ((size_t)LS_THREADS * (E + 1) + (size_t)S * E) * 4;
    AMSS_CUDA(cudaFuncSetAttribute(l41_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(chunks, B);
    AMSS_LAUNCH(l41_fwd_kernel, grid, LS_THREADS, smem, stream, emb, labels, spk, weights, TF, E, S, (float*)workspace);
    AMSS_LAUNCH(l41_fwd_final_kernel, 1, 32, 0, stream, (const float*)workspace, B * chunks,
                (float)B * (float)TF * (float)S, loss);
    return AMSS_OK;
}
extern "C" int amss_l41_loss_bwd(const float* emb, const uint8_t* labels, const float* spk, const float* weights,
                                 const float* dloss, int B, int64_t TF, int E, int S, float* demb, float* dspk,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    AMSS_REQUIRE(emb && labels && spk && dloss && demb && dspk && workspace, "l41_loss_bwd: null pointer");
    AMSS_REQUIRE(S >= 1 && S <= LS_MAXS, "l41_loss_bwd: S out of range");
    if (workspace_bytes < amss_l41_workspace_bytes(B, TF, E, S)) { set_error("l41_loss_bwd: workspace too small"); return AMSS_ERR_WORKSPACE; }
    const int chunks = ls_chunks(B, TF, LS_THREADS);
    float* part = (float*)((char*)workspace + align_up((size_t)B * chunks * 4, 256));
    const size_t smem = ((size_t)LS_THREADS * (E + 1) + (size_t)S * E + (size_t)LS_THREADS * S) * 4;
    AMSS_CUDA(cudaFuncSetAttribute(l41_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(chunks, B);
    AMSS_LAUNCH(l41_bwd_kernel, grid, LS_THREADS, smem, stream, emb, labels, spk, weights, dloss, B, TF, E, S, demb, part);
    AMSS_LAUNCH(l41_dspk_final_kernel, (B * S * E + 255) / 256, 256, 0, stream, part, B, chunks, S * E, dspk);
    return AMSS_OK;
}

extern "C" int amss_plugged_labels(const float* front_y, int B, int S, int64_t TN, uint8_t* labels, void* stream) {
    AMSS_REQUIRE(front_y && labels && S >= 1 && S < 256, "plugged_labels: bad arguments");
    if ((TN & 3) == 0 && ((reinterpret_cast<uintptr_t>(front_y) | reinterpret_cast<uintptr_t>(labels)) & 15) == 0) {
        const int64_t n4 = (int64_t)B * (TN / 4);
        AMSS_LAUNCH(plugged_labels_vec_kernel, (int)std::min<int64_t>((n4 + 255) / 256, 16 * kNumSMs), 256, 0, stream,
                    reinterpret_cast<const float4*>(front_y), B, S, TN / 4, reinterpret_cast<uint32_t*>(labels));
        return AMSS_OK;
    }
    AMSS_LAUNCH(plugged_labels_kernel, 4 * kNumSMs, 256, 0, stream, front_y, B, S, TN, labels);
    return AMSS_OK;
}

extern "C" int amss_wave_stats(const float* target, const float* approx, int R, int64_t L, float* stats,
                               void* stream) {
    AMSS_REQUIRE(target && approx && stats && R > 0, "wave_stats: bad arguments");
    AMSS_LAUNCH(wave_stats_kernel, R, 512, 0, stream, target, approx, L, 1, stats);
    return AMSS_OK;
}

extern "C" int amss_wave_stats_bwd(const float* target, const float* approx, const float* dstats, int R, int64_t L,
                                   float* dapprox, void* stream) {
    AMSS_REQUIRE(target && approx && dstats && dapprox && R > 0 && L > 0, "wave_stats_bwd: bad arguments");
    AMSS_LAUNCH(wave_stats_bwd_kernel, 8 * kNumSMs, 256, 0, stream, target, approx, dstats, R, L, dapprox);
    return AMSS_OK;
}

extern "C" int amss_wave_stats_rows(const float* target, const float* approx, int R, int64_t L, int approx_div, float* stats,
                                    void* stream) {
    AMSS_REQUIRE(target && approx && stats && R > 0 && approx_div > 0, "wave_stats_rows: bad arguments");
    AMSS_LAUNCH(wave_stats_kernel, R, 512, 0, stream, target, approx, L, approx_div, stats);
    return AMSS_OK;
}
