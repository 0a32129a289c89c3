// The elementwise / reduction terms of the Adapt pre-training graph over the front output y[B(S+1),Tp,N]
// (mixture rows first, then the B*S source rows), fused into one forward and one backward pass:
//   * p_hat[j] = sum over ALL rows |y[r,j]|, sparse_constraint = sum_j kl_div(rho, p_hat[j])    (models/adapt.py:127-132,
//     utils/ops.py:46-54: p log(clip(p)/clip(p_hat)) + (1-p) log(clip(1-p)/clip(1-p_hat)), clip to [1e-10, 1]);
//   * overlapping = mean_b mean_pairs mean_j 1 - |a-b| / (max(a,b) + 1e-8), a,b = |y| of two sources   (adapt.py:141-160);
//   * the pre-training separator: 'mask' input_mix * (input_non_mix / input_mix), 'perfect' input_mix - sum of the OTHER
//     sources                                                                                     (adapt.py:162-196);
//   * the non-negativity term mean_rows sum_j min(y,0)^2                                          (adapt.py:315-316).
// One thread owns a column j = (t, n) and walks the rows: every access is coalesced across j, y is read once in the
// forward pass and once in the backward pass (0.26 MB per signal): HBM-bound, 4*(2S+1)*Tp*N bytes per mixture forward.
// Reductions are two-level (per-CTA partials, fixed-order finish): deterministic.
#include "common.cuh"
#include <algorithm>

namespace amss {
namespace {

constexpr int AC_THREADS = 256;
constexpr int AC_MAXS = 4;

__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 1e-10f), 1.f); }

// part[blockIdx] = { sum_j kl_j, sum_{b,pairs,j} measure, sum_{rows,j} neg^2 }
__global__ void __launch_bounds__(AC_THREADS)
adapt_terms_fwd_kernel(const float* __restrict__ y, int B, int S, int64_t TN, float rho, int mode,
                       float* __restrict__ sep, float* __restrict__ p_hat, float* __restrict__ part) {
    __shared__ float red[32];
    const int64_t j = blockIdx.x * (int64_t)AC_THREADS + threadIdx.x;
    float kl = 0.f, ov = 0.f, ng = 0.f;
    if (j < TN) {
        float ph = 0.f;
        for (int b = 0; b < B; ++b) {
            const float m = y[(size_t)b * TN + j];
            ph += fabsf(m);
            ng += m < 0.f ? m * m : 0.f;
            float v[AC_MAXS], tot = 0.f;
#pragma unroll
            for (int s = 0; s < AC_MAXS; ++s)
                if (s < S) {
                    v[s] = y[((size_t)B + (size_t)b * S + s) * TN + j];
                    tot += v[s];
                    ph += fabsf(v[s]);
                    ng += v[s] < 0.f ? v[s] * v[s] : 0.f;
                }
#pragma unroll
            for (int s = 0; s < AC_MAXS; ++s)
                if (s < S) {
                    if (sep) sep[((size_t)b * S + s) * TN + j] = mode == 0 ? m * (v[s] / m) : m - (tot - v[s]);
#pragma unroll
                    for (int s2 = s + 1; s2 < AC_MAXS; ++s2)
                        if (s2 < S) {
                            const float a = fabsf(v[s]), c = fabsf(v[s2]);
                            ov += 1.f - fabsf(a - c) / (fmaxf(a, c) + 1e-8f);
                        }
                }
        }
        p_hat[j] = ph;
        kl = rho * logf(clip01(rho) / clip01(ph)) + (1.f - rho) * logf(clip01(1.f - rho) / clip01(1.f - ph));
    }
    kl = block_sum(kl, red); ov = block_sum(ov, red); ng = block_sum(ng, red);
    if (threadIdx.x == 0) { part[blockIdx.x * 3 + 0] = kl; part[blockIdx.x * 3 + 1] = ov; part[blockIdx.x * 3 + 2] = ng; }
}
// terms = { sparse_constraint, overlapping, nonneg }
__global__ void adapt_terms_finish_kernel(const float* __restrict__ part, int n, float ov_div, float ng_div,
                                          float* __restrict__ terms) {
    __shared__ float red[32];
    float kl = 0.f, ov = 0.f, ng = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { kl += part[i * 3]; ov += part[i * 3 + 1]; ng += part[i * 3 + 2]; }
    kl = block_sum(kl, red); ov = block_sum(ov, red); ng = block_sum(ng, red);
    if (threadIdx.x == 0) { terms[0] = kl; terms[1] = ov / ov_div; terms[2] = ng / ng_div; }
}

// dy[r,j] = dterms[0] * dkl/dp_hat[j] * sign(y) + dterms[1] * d overlap + dterms[2] * d nonneg + (separator)^T dsep
__global__ void __launch_bounds__(AC_THREADS)
adapt_terms_bwd_kernel(const float* __restrict__ y, const float* __restrict__ p_hat, const float* __restrict__ dsep,
                       const float* __restrict__ dterms, int B, int S, int64_t TN, float rho, int mode, float ov_div,
                       float ng_div, float* __restrict__ dy) {
    const int64_t j = blockIdx.x * (int64_t)AC_THREADS + threadIdx.x;
    if (j >= TN) return;
    const float ph = p_hat[j];
    // d/dp_hat of p log(clip(p)/clip(ph)) + (1-p) log(clip(1-p)/clip(1-ph)); clip_by_value passes the gradient inside
    // its range only
    float dk = 0.f;
    if (ph >= 1e-10f && ph <= 1.f) dk -= rho / ph;
    { const float q = 1.f - ph; if (q >= 1e-10f && q <= 1.f) dk += (1.f - rho) / q; }
    const float gk = dterms[0] * dk, go = dterms[1] / ov_div, gn = dterms[2] / ng_div;
    auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
    for (int b = 0; b < B; ++b) {
        const float m = y[(size_t)b * TN + j];
        float v[AC_MAXS], g[AC_MAXS], dv[AC_MAXS], gtot = 0.f;
#pragma unroll
        for (int s = 0; s < AC_MAXS; ++s)
            if (s < S) {
                v[s] = y[((size_t)B + (size_t)b * S + s) * TN + j];
                g[s] = dsep ? dsep[((size_t)b * S + s) * TN + j] : 0.f;
                gtot += g[s];
                dv[s] = gk * sgn(v[s]) + (v[s] < 0.f ? gn * 2.f * v[s] : 0.f);
            }
        float dm = gk * sgn(m) + (m < 0.f ? gn * 2.f * m : 0.f);
#pragma unroll
        for (int s = 0; s < AC_MAXS; ++s)
            if (s < S) {
                if (mode == 0) {           // out = m * f, f = v / m:  d/dv = m * (1/m) ; d/dm = f + m * (-v / m^2)
                    dv[s] += g[s] * m * (1.f / m);
                    dm += g[s] * (v[s] / m) + g[s] * m * (-v[s] / (m * m));
                } else {                   // out_s = m - (tot - v_s)
                    dm += g[s];
                    dv[s] += g[s] - gtot;
                }
#pragma unroll
                for (int s2 = s + 1; s2 < AC_MAXS; ++s2)
                    if (s2 < S) {
                        const float a = fabsf(v[s]), c = fabsf(v[s2]);
                        const float mx = fmaxf(a, c) + 1e-8f, df = a - c, ad = fabsf(df);
                        // measure = 1 - |a-c| / mx ; d/da = -(sgn(a-c) mx - |a-c| [a>=c]) / mx^2 (torch.maximum: the
                        // gradient goes to the larger argument, split evenly on ties)
                        const float wa = a > c ? 1.f : (a == c ? 0.5f : 0.f), wc = 1.f - wa;
                        const float da = -(sgn(df) * mx - ad * wa) / (mx * mx);
                        const float dc = -(-sgn(df) * mx - ad * wc) / (mx * mx);
                        dv[s] += go * da * sgn(v[s]);
                        dv[s2] += go * dc * sgn(v[s2]);
                    }
            }
        dy[(size_t)b * TN + j] = dm;
#pragma unroll
        for (int s = 0; s < AC_MAXS; ++s)
            if (s < S) dy[((size_t)B + (size_t)b * S + s) * TN + j] = dv[s];
    }
}

// ---- the scalar tail of Adapt.cost (pre-training branch) ---------------------------------------------------------------
// models/adapt.py:323-337, 374-385 + models/network.py:196-221 (with_perm = False), from the per-(b,s) waveform statistics
//   st[r] = (<t,t>, <a,a>, <t,a>, <t-a,t-a>)  (target vs synthesis)      sm[r] = (<t,t>, <m,m>, <t,m>, .)  (target vs mixture)
//   l2  = mean_B sum_S <t-a,t-a>                       sdr = mean_{B,S} <t,t><a,a> / (<t,a>^2 + 1e-12)
//   cost = l2 | sdr | l2 + sdr  + beta sparse + lambda^2 (|f|^2 + |f2|^2) / 2 + overlap_coef overlapping + nn^2 nonneg
//   sdr_improvement = mean_{B,S} 10 log10(1 / (<t,t><a,a>/<t,a>^2 - 1)) - 10 log10(1 / (<t,t><m,m>/<t,m>^2 - 1))
// and the derivatives of cost w.r.t. every input (the graph is a scalar function of ~B*S*4 + 4 numbers: ~90 elementwise
// launches of 2-4 us each when written with tensor ops).  One CTA, fixed-order sums.
__global__ void adapt_cost_kernel(const float* __restrict__ st, const float* __restrict__ sm, const float* __restrict__ terms,
                                  const float* __restrict__ regsq, int B, int S, int loss_kind, float beta, float lam, float ov,
                                  float nn, float* __restrict__ out, float* __restrict__ dst, float* __restrict__ dterms,
                                  float* __restrict__ dreg) {
    __shared__ float red[32];
    const int R = B * S;
    const float wl2 = (loss_kind == 0 || loss_kind == 2) ? 1.f / (float)B : 0.f;
    const float wsdr = (loss_kind == 1 || loss_kind == 2) ? 1.f / (float)R : 0.f;
    float l2s = 0.f, sdrs = 0.f, vals = 0.f;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const float tn = st[r * 4 + 0], an = st[r * 4 + 1], ta = st[r * 4 + 2], ee = st[r * 4 + 3];
        const float den = ta * ta + 1e-12f;
        l2s += ee;
        sdrs += tn * an / den;
        if (sm) {
            const float mm = sm[r * 4 + 1], tm = sm[r * 4 + 2], tn2 = sm[r * 4 + 0];
            const float sep = 10.f * logf(1.f / ((tn * an) / (ta * ta) - 1.f)) / logf(10.f);
            const float nsep = 10.f * logf(1.f / ((tn2 * mm) / (tm * tm) - 1.f)) / logf(10.f);
            vals += sep - nsep;
        }
        dst[r * 4 + 0] = wsdr * an / den;
        dst[r * 4 + 1] = wsdr * tn / den;
        dst[r * 4 + 2] = -2.f * wsdr * tn * an * ta / (den * den);
        dst[r * 4 + 3] = wl2;
    }
    l2s = block_sum(l2s, red);
    sdrs = block_sum(sdrs, red);
    vals = block_sum(vals, red);
    if (threadIdx.x == 0) {
        const float l2 = l2s / (float)B, sdr = sdrs / (float)R;
        float cost = loss_kind == 0 ? l2 : (loss_kind == 1 ? sdr : l2 + sdr);
        if (beta != 0.f) cost += beta * terms[0];
        if (lam != 0.f) cost += lam * (lam * (0.5f * regsq[0]));        // lambda applied twice (adapt.py:312, :380)
        if (ov != 0.f) cost += ov * terms[1];
        if (nn != 0.f) cost += nn * (nn * terms[2]);                    // applied twice (:316, :384)
        out[0] = cost; out[1] = l2; out[2] = sdr; out[3] = vals / (float)R;
        dterms[0] = beta; dterms[1] = ov; dterms[2] = nn * nn;
        dreg[0] = lam * lam;
    }
}

}  // namespace
}  // namespace amss

using namespace amss;

extern "C" int amss_adapt_cost_fwd(const float* stats, const float* mix_stats, const float* terms, const float* regsq, int B,
                                   int S, int loss_kind, float beta, float lambda, float overlap_coef, float nonneg_coef,
                                   float* out4, float* dstats, float* dterms, float* dreg, void* stream) {
    AMSS_REQUIRE(stats && terms && regsq && out4 && dstats && dterms && dreg && B > 0 && S > 0, "adapt_cost_fwd: bad arguments");
    AMSS_REQUIRE(loss_kind >= 0 && loss_kind <= 2, "adapt_cost_fwd: loss_kind must be 0 (l2), 1 (sdr) or 2 (l2 + sdr)");
    AMSS_LAUNCH(adapt_cost_kernel, 1, 256, 0, stream, stats, mix_stats, terms, regsq, B, S, loss_kind, beta, lambda, overlap_coef,
                nonneg_coef, out4, dstats, dterms, dreg);
    return AMSS_OK;
}

extern "C" size_t amss_adapt_terms_workspace_bytes(int64_t TN) { return (size_t)((TN + AC_THREADS - 1) / AC_THREADS) * 3 * 4 + 256; }

extern "C" int amss_adapt_terms_fwd(const float* y, int B, int S, int64_t TN, float rho, int separation, float* sep,
                                    float* p_hat, float* terms, void* workspace, size_t workspace_bytes, void* stream) {
    AMSS_REQUIRE(y && p_hat && terms && workspace && B > 0 && S >= 1 && S <= AC_MAXS && TN > 0, "adapt_terms_fwd: bad arguments");
    AMSS_REQUIRE(separation == 0 || separation == 1, "adapt_terms_fwd: separation must be 0 (mask) or 1 (perfect)");
    if (workspace_bytes < amss_adapt_terms_workspace_bytes(TN)) { set_error("adapt_terms_fwd: workspace too small"); return AMSS_ERR_WORKSPACE; }
    const int grid = (int)((TN + AC_THREADS - 1) / AC_THREADS);
    const int pairs = std::max(1, S * (S - 1) / 2);
    AMSS_LAUNCH(adapt_terms_fwd_kernel, grid, AC_THREADS, 0, stream, y, B, S, TN, rho, separation, sep, p_hat, (float*)workspace);
    AMSS_LAUNCH(adapt_terms_finish_kernel, 1, 256, 0, stream, (const float*)workspace, grid, (float)B * pairs * (float)TN,
                (float)B * (S + 1), terms);
    return AMSS_OK;
}

extern "C" int amss_adapt_terms_bwd(const float* y, const float* p_hat, const float* dsep, const float* dterms, int B, int S,
                                    int64_t TN, float rho, int separation, float* dy, void* stream) {
    AMSS_REQUIRE(y && p_hat && dterms && dy && B > 0 && S >= 1 && S <= AC_MAXS && TN > 0, "adapt_terms_bwd: bad arguments");
    const int grid = (int)((TN + AC_THREADS - 1) / AC_THREADS);
    const int pairs = std::max(1, S * (S - 1) / 2);
    AMSS_LAUNCH(adapt_terms_bwd_kernel, grid, AC_THREADS, 0, stream, y, p_hat, dsep, dterms, B, S, TN, rho, separation,
                (float)B * pairs * (float)TN, (float)B * (S + 1), dy);
    return AMSS_OK;
}
