// Loss kernels of the separator: DPCL affinity loss (models/dpcl.py:41-86), L41 sigmoid-dot loss
// (models/L41.py:150-178), per-group L2 normalisation (utils/ops.py:318-324), the plugged-mode
// label arg-max (models/network.py:369-378) and the waveform statistics of the Adapt pretraining
// cost (models/adapt.py:323-330, models/network.py:196-221).
// All reductions are two-level (per-CTA partials, fixed-order finalize): deterministic.
#include "common.cuh"
#include <algorithm>

namespace amss {
namespace {

constexpr int LS_TILE = 128;    // points staged per tile
constexpr int LS_THREADS = 256;
constexpr int LS_MAXS = 4;
constexpr int LS_RB = 4;        // register block (RB x RB entries of the E x E Gram matrix per thread)

// ---- DPCL forward --------------------------------------------------------------------------
// Y is one-hot, so with N_s = #bins of speaker s and D_i = N_{l_i}^{-1/2}:
//   V^T D V = sum_s N_s^{-1/2} G_s,  G_s = sum_{i in s} v_i v_i^T   (E x E)
//   V^T D Y[:, s] = N_s^{-1/2} sum_{i in s} v_i                      (E)
//   Y^T D Y = diag(sqrt(N_s))
// part[b][chunk][s][E*E + E + 1] = (G_s, m_s, N_s) partials.
__global__ void __launch_bounds__(LS_THREADS)
dpcl_gram_kernel(const float* __restrict__ V, const uint8_t* __restrict__ labels, int64_t TF, int E, int S,
                 float* __restrict__ part) {
    extern __shared__ __align__(16) unsigned char ls_smem[];
    float* xs = reinterpret_cast<float*>(ls_smem);              // [LS_TILE][E+1]
    uint8_t* ls = reinterpret_cast<uint8_t*>(xs + LS_TILE * (E + 1));
    const int b = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x, tid = threadIdx.x, EP = E + 1;
    const int nb = (E + LS_RB - 1) / LS_RB;                     // blocks per side
    // thread -> (bi, bj) block of the Gram matrix (threads beyond nb*nb idle in the Gram part)
    const int bi = tid / nb, bj = tid % nb;
    const bool active = bi < nb;
    float acc[LS_MAXS][LS_RB][LS_RB];
    float macc[LS_MAXS];
    float cnt[LS_MAXS];
#pragma unroll
    for (int s = 0; s < LS_MAXS; ++s) {
        macc[s] = 0.f; cnt[s] = 0.f;
#pragma unroll
        for (int i = 0; i < LS_RB; ++i)
#pragma unroll
            for (int j = 0; j < LS_RB; ++j) acc[s][i][j] = 0.f;
    }
    const int64_t ntiles = (TF + LS_TILE - 1) / LS_TILE;
    for (int64_t tile = chunk; tile < ntiles; tile += chunks) {
        const int64_t p0 = tile * LS_TILE;
        const int np = (int)((TF - p0) < LS_TILE ? (TF - p0) : LS_TILE);
        const float* src = V + ((size_t)b * TF + p0) * E;
        __syncthreads();
        for (int i = tid; i < np * E; i += LS_THREADS) { const int p = i / E, e = i - p * E; xs[p * EP + e] = src[i]; }
        for (int i = tid; i < np; i += LS_THREADS) ls[i] = labels[(size_t)b * TF + p0 + i];
        __syncthreads();
        if (active) {
            for (int p = 0; p < np; ++p) {
                const int l = ls[p];
                float vi[LS_RB], vj[LS_RB];
#pragma unroll
                for (int i = 0; i < LS_RB; ++i) {
                    const int ei = bi * LS_RB + i, ej = bj * LS_RB + i;
                    vi[i] = ei < E ? xs[p * EP + ei] : 0.f;
                    vj[i] = ej < E ? xs[p * EP + ej] : 0.f;
                }
#pragma unroll
                for (int s = 0; s < LS_MAXS; ++s)
                    if (s < S) {
                        const float w = (l == s) ? 1.f : 0.f;
#pragma unroll
                        for (int i = 0; i < LS_RB; ++i) {
                            const float wi = w * vi[i];
#pragma unroll
                            for (int j = 0; j < LS_RB; ++j) acc[s][i][j] = fmaf(wi, vj[j], acc[s][i][j]);
                        }
                    }
            }
        }
        // column sums and counts: thread e < E owns m_s[e]; thread E owns the counts
        if (tid < E) {
            for (int p = 0; p < np; ++p) {
                const int l = ls[p];
                const float v = xs[p * EP + tid];
#pragma unroll
                for (int s = 0; s < LS_MAXS; ++s) if (s < S) macc[s] += (l == s) ? v : 0.f;
            }
        } else if (tid == E) {
            for (int p = 0; p < np; ++p) {
                const int l = ls[p];
#pragma unroll
                for (int s = 0; s < LS_MAXS; ++s) if (s < S) cnt[s] += (l == s) ? 1.f : 0.f;
            }
        }
    }
    const int stride = E * E + E + 1;
    float* dst = part + ((size_t)b * chunks + chunk) * S * stride;
#pragma unroll
    for (int s = 0; s < LS_MAXS; ++s)
        if (s < S) {
            if (active) {
#pragma unroll
                for (int i = 0; i < LS_RB; ++i)
#pragma unroll
                    for (int j = 0; j < LS_RB; ++j) {
                        const int ei = bi * LS_RB + i, ej = bj * LS_RB + j;
                        if (ei < E && ej < E) dst[s * stride + ei * E + ej] = acc[s][i][j];
                    }
            }
            if (tid < E) dst[s * stride + E * E + tid] = macc[s];
            else if (tid == E) dst[s * stride + E * E + E] = cnt[s];
        }
}

// One CTA per batch row: reduce the partials, form the three Frobenius norms, keep what the
// backward needs:  stats[b] = { An[E*E] = 2*A/||A||, Bn[S*E] = 2*Bm[:,s]/||Bm||, dinv[S], loss_b }.
__global__ void dpcl_finalize_kernel(const float* __restrict__ part, int chunks, int E, int S,
                                     float* __restrict__ stats) {
    extern __shared__ __align__(16) unsigned char ls_smem[];
    __shared__ float red[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int stride = E * E + E + 1;
    float* g = reinterpret_cast<float*>(ls_smem);               // [S][stride]
    for (int i = tid; i < S * stride; i += blockDim.x) {
        float a = 0.f;
        for (int ch = 0; ch < chunks; ++ch) a += part[((size_t)b * chunks + ch) * S * stride + i];
        g[i] = a;
    }
    __syncthreads();
    const int sstride = E * E + S * E + S + 1;
    float* out = stats + (size_t)b * sstride;
    float sa = 0.f, sb = 0.f;
    for (int i = tid; i < E * E; i += blockDim.x) {
        float a = 0.f;
        for (int s = 0; s < S; ++s) { const float n = g[s * stride + E * E + E]; if (n > 0.f) a += g[s * stride + i] / sqrtf(n); }
        out[i] = a;
        sa = fmaf(a, a, sa);
    }
    for (int i = tid; i < S * E; i += blockDim.x) {
        const int s = i / E, e = i - s * E;
        const float n = g[s * stride + E * E + E];
        const float v = n > 0.f ? g[s * stride + E * E + e] / sqrtf(n) : 0.f;
        out[E * E + i] = v;
        sb = fmaf(v, v, sb);
    }
    sa = block_sum(sa, red);
    sb = block_sum(sb, red);
    float sc = 0.f;
    for (int s = 0; s < S; ++s) sc += g[s * stride + E * E + E];   // ||diag(sqrt(N_s))||_F^2 = sum N_s
    const float na = sqrtf(sa), nbm = sqrtf(sb), nc = sqrtf(sc);
    __syncthreads();
    for (int i = tid; i < E * E; i += blockDim.x) out[i] = 2.f * out[i] / na;
    for (int i = tid; i < S * E; i += blockDim.x) out[E * E + i] = 2.f * out[E * E + i] / nbm;
    if (tid < S) { const float n = g[tid * stride + E * E + E]; out[E * E + S * E + tid] = n > 0.f ? 1.f / sqrtf(n) : 0.f; }
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
__global__ void __launch_bounds__(LS_THREADS)
dpcl_bwd_kernel(const float* __restrict__ V, const uint8_t* __restrict__ labels, const float* __restrict__ dloss,
                const float* __restrict__ stats, int B, int64_t TF, int E, int S, float* __restrict__ dV) {
    extern __shared__ __align__(16) unsigned char ls_smem[];
    const int EP = E + 1;
    float* xs = reinterpret_cast<float*>(ls_smem);              // [LS_THREADS][E+1]
    float* An = xs + LS_THREADS * EP;                           // [E][E]
    float* Bn = An + E * E;                                     // [S][E]
    float* dinv = Bn + S * E;                                   // [S]
    const int b = blockIdx.y, tid = threadIdx.x;
    const int sstride = E * E + S * E + S + 1;
    const float* st = stats + (size_t)b * sstride;
    for (int i = tid; i < E * E + S * E + S; i += LS_THREADS) An[i] = st[i];
    const float gscale = dloss[0] / (float)B;
    const int64_t ntiles = (TF + LS_THREADS - 1) / LS_THREADS;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t p0 = tile * LS_THREADS;
        const int np = (int)((TF - p0) < LS_THREADS ? (TF - p0) : LS_THREADS);
        const float* src = V + ((size_t)b * TF + p0) * E;
        __syncthreads();
        for (int i = tid; i < np * E; i += LS_THREADS) { const int p = i / E, e = i - p * E; xs[p * EP + e] = src[i]; }
        __syncthreads();
        if (tid < np) {
            const int l = labels[(size_t)b * TF + p0 + tid];
            const float d = gscale * dinv[l];
            float* xp = xs + tid * EP;
            // out[e] = sum_e2 An[e][e2] v[e2]; computed into registers in chunks of 8 rows
            for (int e0 = 0; e0 < E; e0 += 8) {
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = 0.f;
                for (int e2 = 0; e2 < E; ++e2) {
                    const float v = xp[e2];
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (e0 + i < E) o[i] = fmaf(An[(e0 + i) * E + e2], v, o[i]);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (e0 + i < E) dV[((size_t)b * TF + p0 + tid) * E + e0 + i] = d * (o[i] - Bn[l * E + e0 + i]);
            }
        }
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
               int64_t TF, int E, int S, float* __restrict__ part) {
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
            for (int s = 0; s < S; ++s) {
                float d = 0.f;
                for (int e = 0; e < E; ++e) d = fmaf(sp[s * E + e], xs[tid * EP + e], d);
                const float x = (l == s) ? d : -d;
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
               const float* __restrict__ dloss, int B, int64_t TF, int E, int S, float* __restrict__ demb,
               float* __restrict__ dspk_part) {
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
            for (int s = 0; s < S; ++s) {
                float d = 0.f;
                for (int e = 0; e < E; ++e) d = fmaf(sp[s * E + e], xs[tid * EP + e], d);
                const float y = (l == s) ? 1.f : -1.f;
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

// stats[r] = { <t,t>, <a,a>, <t,a>, <t-a,t-a> } over L   (fixed-order two-level reduction)
__global__ void wave_stats_kernel(const float* __restrict__ tg, const float* __restrict__ ap, int64_t L,
                                  float* __restrict__ stats) {
    __shared__ float red[32];
    const int r = blockIdx.x;
    float tt = 0.f, aa = 0.f, ta = 0.f, ee = 0.f;
    for (int64_t i = threadIdx.x; i < L; i += blockDim.x) {
        const float t = tg[(size_t)r * L + i], a = ap[(size_t)r * L + i], d = t - a;
        tt = fmaf(t, t, tt); aa = fmaf(a, a, aa); ta = fmaf(t, a, ta); ee = fmaf(d, d, ee);
    }
    tt = block_sum(tt, red); aa = block_sum(aa, red); ta = block_sum(ta, red); ee = block_sum(ee, red);
    if (threadIdx.x == 0) { stats[r * 4 + 0] = tt; stats[r * 4 + 1] = aa; stats[r * 4 + 2] = ta; stats[r * 4 + 3] = ee; }
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
    const size_t stride = (size_t)E * E + E + 1;
    const size_t sstride = (size_t)E * E + (size_t)S * E + S + 1;
    return align_up((size_t)B * sstride * 4, 256) + align_up((size_t)B * ls_chunks(B, TF, LS_TILE) * S * stride * 4, 256);
}

extern "C" int amss_dpcl_loss_fwd(const float* V, const uint8_t* labels, int B, int64_t TF, int E, int S, float* loss,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    AMSS_REQUIRE(V && labels && loss && workspace, "dpcl_loss_fwd: null pointer");
    AMSS_REQUIRE(S >= 1 && S <= LS_MAXS, "dpcl_loss_fwd: S=%d outside [1,%d]", S, LS_MAXS);
    const int nb = (E + LS_RB - 1) / LS_RB;
    AMSS_REQUIRE(E >= 1 && nb * nb <= LS_THREADS, "dpcl_loss_fwd: E=%d too large (max %d)", E, 16 * LS_RB);
    if (workspace_bytes < amss_dpcl_workspace_bytes(B, TF, E, S)) { set_error("dpcl_loss_fwd: workspace too small"); return AMSS_ERR_WORKSPACE; }
    const int sstride = E * E + S * E + S + 1;
    float* stats = (float*)workspace;
    float* part = (float*)((char*)workspace + align_up((size_t)B * sstride * 4, 256));
    const int chunks = ls_chunks(B, TF, LS_TILE);
    const size_t smem1 = (size_t)LS_TILE * (E + 1) * 4 + LS_TILE;
    dim3 grid(chunks, B);
    AMSS_LAUNCH(dpcl_gram_kernel, grid, LS_THREADS, smem1, stream, V, labels, TF, E, S, part);
    const size_t smem2 = (size_t)S * (E * E + E + 1) * 4;
    AMSS_LAUNCH(dpcl_finalize_kernel, B, 256, smem2, stream, part, chunks, E, S, stats);
    AMSS_LAUNCH(mean_of_stat_kernel, 1, 32, 0, stream, stats, B, sstride, sstride - 1, loss);
    return AMSS_OK;
}

extern "C" int amss_dpcl_loss_bwd(const float* V, const uint8_t* labels, const float* dloss, int B, int64_t TF, int E,
                                  int S, float* dV, const void* workspace, void* stream) {
    AMSS_REQUIRE(V && labels && dloss && dV && workspace, "dpcl_loss_bwd: null pointer");
    AMSS_REQUIRE(S >= 1 && S <= LS_MAXS, "dpcl_loss_bwd: S out of range");
    const size_t smem = ((size_t)LS_THREADS * (E + 1) + (size_t)E * E + (size_t)S * E + S) * 4;
    AMSS_REQUIRE(smem <= 200 * 1024, "dpcl_loss_bwd: E too large");
    AMSS_CUDA(cudaFuncSetAttribute(dpcl_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ls_chunks(B, TF, LS_THREADS), B);
    AMSS_LAUNCH(dpcl_bwd_kernel, grid, LS_THREADS, smem, stream, V, labels, dloss, (const float*)workspace, B, TF, E,
                S, dV);
    return AMSS_OK;
}

extern "C" int amss_l2norm_fwd(const float* z, int64_t rows, int E, float* v, float* inv_norm, void* stream) {
    AMSS_REQUIRE(z && v && rows > 0 && E > 0, "l2norm_fwd: bad arguments");
    AMSS_LAUNCH(l2norm_fwd_kernel, 8 * kNumSMs, 256, 0, stream, z, rows, E, v, inv_norm);
    return AMSS_OK;
}
extern "C" int amss_l2norm_bwd(const float* v, const float* inv_norm, const float* dv, int64_t rows, int E, float* dz,
                               void* stream) {
    AMSS_REQUIRE(v && inv_norm && dv && dz, "l2norm_bwd: null pointer");
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

extern "C" size_t amss_l41_workspace_bytes(int B, int64_t TF, int E, int S) {
    const int chunks = ls_chunks(B, TF, LS_THREADS);
    return align_up((size_t)B * chunks * 4, 256) + align_up((size_t)B * chunks * S * E * 4, 256);
}
extern "C" int amss_l41_loss_fwd(const float* emb, const uint8_t* labels, const float* spk, int B, int64_t TF, int E,
                                 int S, float* loss, void* workspace, size_t workspace_bytes, void* stream) {
    AMSS_REQUIRE(emb && labels && spk && loss && workspace, "l41_loss_fwd: null pointer");
    AMSS_REQUIRE(S >= 1 && S <= LS_MAXS, "l41_loss_fwd: S out of range");
    if (workspace_bytes < amss_l41_workspace_bytes(B, TF, E, S)) { set_error("l41_loss_fwd: workspace too small"); return AMSS_ERR_WORKSPACE; }
    const int chunks = ls_chunks(B, TF, LS_THREADS);
    const size_t smem = ((size_t)LS_THREADS * (E + 1) + (size_t)S * E) * 4;
    AMSS_CUDA(cudaFuncSetAttribute(l41_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(chunks, B);
    AMSS_LAUNCH(l41_fwd_kernel, grid, LS_THREADS, smem, stream, emb, labels, spk, TF, E, S, (float*)workspace);
    AMSS_LAUNCH(l41_fwd_final_kernel, 1, 32, 0, stream, (const float*)workspace, B * chunks,
                (float)B * (float)TF * (float)S, loss);
    return AMSS_OK;
}
extern "C" int amss_l41_loss_bwd(const float* emb, const uint8_t* labels, const float* spk, const float* dloss, int B,
                                 int64_t TF, int E, int S, float* demb, float* dspk, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    AMSS_REQUIRE(emb && labels && spk && dloss && demb && dspk && workspace, "l41_loss_bwd: null pointer");
    AMSS_REQUIRE(S >= 1 && S <= LS_MAXS, "l41_loss_bwd: S out of range");
    if (workspace_bytes < amss_l41_workspace_bytes(B, TF, E, S)) { set_error("l41_loss_bwd: workspace too small"); return AMSS_ERR_WORKSPACE; }
    const int chunks = ls_chunks(B, TF, LS_THREADS);
    float* part = (float*)((char*)workspace + align_up((size_t)B * chunks * 4, 256));
    const size_t smem = ((size_t)LS_THREADS * (E + 1) + (size_t)S * E + (size_t)LS_THREADS * S) * 4;
    AMSS_CUDA(cudaFuncSetAttribute(l41_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(chunks, B);
    AMSS_LAUNCH(l41_bwd_kernel, grid, LS_THREADS, smem, stream, emb, labels, spk, dloss, B, TF, E, S, demb, part);
    AMSS_LAUNCH(l41_dspk_final_kernel, (B * S * E + 255) / 256, 256, 0, stream, part, B, chunks, S * E, dspk);
    return AMSS_OK;
}

extern "C" int amss_plugged_labels(const float* front_y, int B, int S, int64_t TN, uint8_t* labels, void* stream) {
    AMSS_REQUIRE(front_y && labels && S >= 1 && S < 256, "plugged_labels: bad arguments");
    AMSS_LAUNCH(plugged_labels_kernel, 4 * kNumSMs, 256, 0, stream, front_y, B, S, TN, labels);
    return AMSS_OK;
}

extern "C" int amss_wave_stats(const float* target, const float* approx, int R, int64_t L, float* stats,
                               void* stream) {
    AMSS_REQUIRE(target && approx && stats && R > 0, "wave_stats: bad arguments");
    AMSS_LAUNCH(wave_stats_kernel, R, 512, 0, stream, target, approx, L, stats);
    return AMSS_OK;
}
