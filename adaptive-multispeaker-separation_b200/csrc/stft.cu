// STFT twin of the adaptive front end and its inverse.
//   forward : tf.contrib.signal.stft(x, frame, hop, fft_length=frame)        models/network.py:480-502
//   inverse : (mask*|X|)*exp(j*angle(X)) -> tf.contrib.signal.inverse_stft   models/network.py:584-607
// Both are HBM/launch-bound (about 1.35 MB of traffic per 4 s mixture): one CTA per frame does
// frame + periodic-hann window + radix-2 FFT in shared memory and writes |X| / X (and the
// arg-max-over-speakers label) straight out; the inverse fuses mask, Hermitian fill, inverse FFT,
// the inverse_stft_window_fn normalisation and the overlap-add gather (no atomics, no scratch).
#include "common.cuh"
#include <cstdlib>

namespace amss {
namespace {

// In-place radix-2 DIT FFT of N complex points held bit-reversed in `buf`; N/2 threads.
// tw[k] = exp(-+ 2*pi*i*k/N), k < N/2 (sign chosen by the caller when filling tw).
// Two consecutive stages are fused per barrier (the four elements {i, i+h, i+2h, i+3h} are closed under stages s and
// s+1): every element crosses shared memory once per TWO stages -- the kernels are bound by shared-memory traffic, 9 stages
// of 5 x 8 bytes per thread -- with exactly the operations of the stage-by-stage form (bit-identical results).
__device__ __forceinline__ float2 cmul_(float2 b, float2 w) { return make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x); }
__device__ __forceinline__ void fft_shared(float2* buf, const float2* tw, int N, int logN) {
    const int tid = threadIdx.x;
    int s = 1;
    if (logN & 1) {                                             // odd number of stages: stage 1 on its own
        __syncthreads();
        if (tid < N / 2) {
            const int i0 = tid << 1;
            const float2 a = buf[i0], bw = cmul_(buf[i0 + 1], tw[0]);
            buf[i0] = make_float2(a.x + bw.x, a.y + bw.y);
            buf[i0 + 1] = make_float2(a.x - bw.x, a.y - bw.y);
        }
        s = 2;
    }
    for (; s < logN; s += 2) {
        const int h = 1 << (s - 1);
        __syncthreads();
        if (tid < N / 4) {
            const int grp = tid >> (s - 1), j = tid & (h - 1);
            const int ia = (grp << (s + 1)) + j, ib = ia + h, ic = ia + 2 * h, id = ia + 3 * h;
            const float2 w1 = tw[j << (logN - s)], w2 = tw[j << (logN - s - 1)], w3 = tw[(j + h) << (logN - s - 1)];
            const float2 a = buf[ia], c = buf[ic];
            const float2 bw = cmul_(buf[ib], w1), dw = cmul_(buf[id], w1);
            const float2 a1 = make_float2(a.x + bw.x, a.y + bw.y), b1 = make_float2(a.x - bw.x, a.y - bw.y);
            const float2 c1 = make_float2(c.x + dw.x, c.y + dw.y), d1 = make_float2(c.x - dw.x, c.y - dw.y);
            const float2 cw = cmul_(c1, w2), ew = cmul_(d1, w3);
            buf[ia] = make_float2(a1.x + cw.x, a1.y + cw.y);
            buf[ic] = make_float2(a1.x - cw.x, a1.y - cw.y);
            buf[ib] = make_float2(b1.x + ew.x, b1.y + ew.y);
            buf[id] = make_float2(b1.x - ew.x, b1.y - ew.y);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ float hann_periodic(int i, int N) {
    // tf.contrib.signal.hann_window(N, periodic=True) = 0.5 - 0.5*cos(2*pi*i/N)
    return 0.5f - 0.5f * cospif(2.0f * (float)i / (float)N);
}

__device__ __forceinline__ void fill_twiddles(float2* tw, int N, float sign) {
    for (int k = threadIdx.x; k < N / 2; k += blockDim.x) {
        float s, c;
        sincospif(2.0f * (float)k / (float)N, &s, &c);
        tw[k] = make_float2(c, sign * s);
    }
}

// grid (T, R); block N/2 (>=32).  spec/mag optional.
__global__ void stft_fwd_kernel(const float* __restrict__ x, int64_t L, int N, int logN, int hop, int T,
                                float2* __restrict__ spec, float* __restrict__ mag) {
    extern __shared__ __align__(16) unsigned char st_smem[];
    float2* buf = reinterpret_cast<float2*>(st_smem);
    float2* tw = buf + N;
    const int t = blockIdx.x, r = blockIdx.y, F = N / 2 + 1;
    fill_twiddles(tw, N, -1.f);
    const float* src = x + (size_t)r * L + (size_t)t * hop;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float v = src[i] * hann_periodic(i, N);
        buf[__brev((unsigned)i) >> (32 - logN)] = make_float2(v, 0.f);
    }
    fft_shared(buf, tw, N, logN);
    const size_t o = ((size_t)r * T + t) * F;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        const float2 v = buf[f];
        if (spec) spec[o + f] = v;
        if (mag) mag[o + f] = sqrtf(v.x * v.x + v.y * v.y);
    }
}

// grid (T, B): labels[b,t,f] = argmax_s |stft(non_mix[b,s])|[t,f]   (first index wins ties)
__global__ void stft_labels_kernel(const float* __restrict__ x, int S, int64_t L, int N, int logN, int hop, int T,
                                   uint8_t* __restrict__ labels, float* __restrict__ mag_nm) {
    extern __shared__ __align__(16) unsigned char st_smem[];
    float2* buf = reinterpret_cast<float2*>(st_smem);
    float2* tw = buf + N;
    const int t = blockIdx.x, b = blockIdx.y, F = N / 2 + 1;
    fill_twiddles(tw, N, -1.f);
    // each thread owns bins f = tid and (thread 0) f = N/2
    float best0 = -1.f, best1 = -1.f;
    int lab0 = 0, lab1 = 0;
    const size_t o = ((size_t)b * T + t) * F;
    for (int s = 0; s < S; ++s) {
        const float* src = x + ((size_t)b * S + s) * L + (size_t)t * hop;
        __syncthreads();
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const float v = src[i] * hann_periodic(i, N);
            buf[__brev((unsigned)i) >> (32 - logN)] = make_float2(v, 0.f);
        }
        fft_shared(buf, tw, N, logN);
        {
            const float2 v = buf[threadIdx.x];
            const float m = sqrtf(v.x * v.x + v.y * v.y);
            if (m > best0) { best0 = m; lab0 = s; }
            if (mag_nm) mag_nm[(o + threadIdx.x) * S + s] = m;
        }
        if (threadIdx.x == 0) {
            const float2 v = buf[N / 2];
            const float m = sqrtf(v.x * v.x + v.y * v.y);
            if (m > best1) { best1 = m; lab1 = s; }
            if (mag_nm) mag_nm[(o + N / 2) * S + s] = m;
        }
    }
    labels[o + threadIdx.x] = (uint8_t)lab0;
    if (threadIdx.x == 0) labels[o + N / 2] = (uint8_t)lab1;
}

// ---- forward, several frames per CTA -------------------------------------------------------------------------------------------
// The per-frame kernels above spend most of a CTA's life on its 256 twiddles, its 512 window values and nine barrier-separated
// butterfly stages for ONE real frame.  Here a CTA owns ST_FR consecutive frames of one signal: twiddles and window are
// built once, and two real frames share one complex FFT (z = x_t + i x_{t+1}: X_t[f] = (Z[f] + conj Z[N-f]) / 2,
// X_{t+1}[f] = (Z[f] - conj Z[N-f]) / 2i).  grid (ceil(T / ST_FR), R); block N/2.
constexpr int ST_FR = 16;
// frames per CTA: ST_FR, fewer (an even number >= 2) while the grid would not fill two CTAs per SM
inline int st_frames_per_cta(int T, int R) {
    int fr = ST_FR;
    while (fr > 2 && (int64_t)((T + fr - 1) / fr) * R < 2 * kNumSMs) fr -= 2;
    return fr;
}

// spectra of the two frames packed in `buf` at bin f (0 <= f <= N/2)
__device__ __forceinline__ void unpack_pair(const float2* buf, int f, int N, float2& a, float2& b) {
    const float2 zf = buf[f], zn = buf[(N - f) & (N - 1)];
    a = make_float2(0.5f * (zf.x + zn.x), 0.5f * (zf.y - zn.y));
    b = make_float2(0.5f * (zf.y + zn.y), -0.5f * (zf.x - zn.x));
}

__global__ void stft_fwd_run_kernel(const float* __restrict__ x, int64_t L, int N, int logN, int hop, int T, int FR,
                                    float2* __restrict__ spec, float* __restrict__ mag) {
    extern __shared__ __align__(16) unsigned char st_smem[];
    float2* buf = reinterpret_cast<float2*>(st_smem);
    float2* tw = buf + N;
    float* win = reinterpret_cast<float*>(tw + N / 2);
    const int r = blockIdx.y, F = N / 2 + 1, tid = threadIdx.x;
    const int tA = blockIdx.x * FR, tB = min(tA + FR, T);
    fill_twiddles(tw, N, -1.f);
    for (int i = tid; i < N; i += blockDim.x) win[i] = hann_periodic(i, N);
    for (int t = tA; t < tB; t += 2) {
        const bool two = t + 1 < tB;
        const float* s0 = x + (size_t)r * L + (size_t)t * hop;
        __syncthreads();                                        // window / twiddles ready; the previous pair has been read out
        for (int i = tid; i < N; i += blockDim.x) {
            const float w = win[i];
            buf[__brev((unsigned)i) >> (32 - logN)] = make_float2(s0[i] * w, two ? s0[hop + i] * w : 0.f);
        }
        fft_shared(buf, tw, N, logN);
        for (int f = tid; f < F; f += blockDim.x) {
            float2 a, b;
            unpack_pair(buf, f, N, a, b);
            const size_t o = ((size_t)r * T + t) * F + f;
            if (spec) { spec[o] = a; if (two) spec[o + F] = b; }
            if (mag) { mag[o] = sqrtf(a.x * a.x + a.y * a.y); if (two) mag[o + F] = sqrtf(b.x * b.x + b.y * b.y); }
        }
    }
}

// grid (ceil(T / ST_FR), B): labels[b,t,f] = argmax_s |stft(non_mix[b,s])|[t,f]   (first index wins ties)
__global__ void stft_labels_run_kernel(const float* __restrict__ x, int S, int64_t L, int N, int logN, int hop, int T, int FR,
                                       uint8_t* __restrict__ labels, float* __restrict__ mag_nm) {
    extern __shared__ __align__(16) unsigned char st_smem[];
    float2* buf = reinterpret_cast<float2*>(st_smem);
    float2* tw = buf + N;
    float* win = reinterpret_cast<float*>(tw + N / 2);
    const int b = blockIdx.y, F = N / 2 + 1, tid = threadIdx.x;
    const int tA = blockIdx.x * FR, tB = min(tA + FR, T);
    fill_twiddles(tw, N, -1.f);
    for (int i = tid; i < N; i += blockDim.x) win[i] = hann_periodic(i, N);
    for (int t = tA; t < tB; t += 2) {
        const bool two = t + 1 < tB;
        // each thread owns bins f = tid and (thread 0) f = N/2, of both frames of the pair
        float best[2][2] = {{-1.f, -1.f}, {-1.f, -1.f}};        // [frame][own bin]
        int lab[2][2] = {{0, 0}, {0, 0}};
        const size_t o = ((size_t)b * T + t) * F;
        for (int sidx = 0; sidx < S; ++sidx) {
            const float* s0 = x + ((size_t)b * S + sidx) * L + (size_t)t * hop;
            __syncthreads();
            for (int i = tid; i < N; i += blockDim.x) {
                const float w = win[i];
                buf[__brev((unsigned)i) >> (32 - logN)] = make_float2(s0[i] * w, two ? s0[hop + i] * w : 0.f);
            }
            fft_shared(buf, tw, N, logN);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (q == 1 && tid != 0) break;
                const int f = q == 0 ? tid : N / 2;
                float2 a, c;
                unpack_pair(buf, f, N, a, c);
                const float m0 = sqrtf(a.x * a.x + a.y * a.y), m1 = sqrtf(c.x * c.x + c.y * c.y);
                if (m0 > best[0][q]) { best[0][q] = m0; lab[0][q] = sidx; }
                if (m1 > best[1][q]) { best[1][q] = m1; lab[1][q] = sidx; }
                if (mag_nm) {
                    mag_nm[(o + f) * S + sidx] = m0;
                    if (two) mag_nm[(o + F + f) * S + sidx] = m1;
                }
            }
        }
        labels[o + tid] = (uint8_t)lab[0][0];
        if (two) labels[o + F + tid] = (uint8_t)lab[1][0];
        if (tid == 0) {
            labels[o + N / 2] = (uint8_t)lab[0][1];
            if (two) labels[o + F + N / 2] = (uint8_t)lab[1][1];
        }
    }
}

// grid (nblocks = T-1+frame/hop, B*S); block N/2.  out[bs, m*hop + j], j < hop.
__global__ void istft_masked_kernel(const float2* __restrict__ spec, const int* __restrict__ labels,
                                    const float* __restrict__ masks, int S, int T, int N, int logN, int hop,
                                    float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char st_smem[];
    float2* buf = reinterpret_cast<float2*>(st_smem);
    float2* tw = buf + N;
    float* acc = reinterpret_cast<float*>(tw + N / 2);   // hop floats
    const int m = blockIdx.x, bs = blockIdx.y, b = bs / S, s = bs % S;
    const int F = N / 2 + 1, r = N / hop;
    const int64_t Lout = (int64_t)(T - 1) * hop + N;
    fill_twiddles(tw, N, +1.f);
    for (int j = threadIdx.x; j < hop; j += blockDim.x) acc[j] = 0.f;
    for (int t = m - r + 1; t <= m; ++t) {
        if (t < 0 || t >= T) continue;   // block-uniform
        const size_t o = ((size_t)b * T + t) * F;
        __syncthreads();
        for (int k = threadIdx.x; k < F; k += blockDim.x) {
            const float w = labels ? (labels[o + k] == s ? 1.f : 0.f) : masks[(o + k) * S + s];
            float2 v = spec[o + k];
            v.x *= w; v.y *= w;
            if (k == 0 || k == N / 2) {   // irfft ignores the imaginary part of DC and Nyquist
                buf[__brev((unsigned)k) >> (32 - logN)] = make_float2(v.x, 0.f);
            } else {
                buf[__brev((unsigned)k) >> (32 - logN)] = v;
                buf[__brev((unsigned)(N - k)) >> (32 - logN)] = make_float2(v.x, -v.y);
            }
        }
        fft_shared(buf, tw, N, logN);
        const int base = (m - t) * hop;
        for (int j = threadIdx.x; j < hop; j += blockDim.x) {
            const int i = base + j;
            // inverse_stft_window_fn(hop)(N): w[i] / sum_o w^2[(i mod hop) + o*hop]
            float den = 0.f;
            for (int o2 = 0; o2 < r; ++o2) { const float w = hann_periodic(j + o2 * hop, N); den = fmaf(w, w, den); }
            acc[j] += buf[i].x * (1.0f / (float)N) * (hann_periodic(i, N) / den);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < hop; j += blockDim.x) {
        const int64_t u = (int64_t)m * hop + j;
        if (u < Lout) out[(size_t)bs * Lout + u] = acc[j];
    }
}

// The same inverse with every frame's inverse FFT evaluated ONCE, two frames per transform: a CTA owns IR_Q consecutive
// output blocks of one signal and walks the frames that touch them in time order.  The (Hermitian-filled) spectra of frames
// t and t+1 are combined as Z = X_t + i X_{t+1}: one complex inverse FFT returns frame t in the real and frame t+1 in the
// imaginary part.  Every frame adds its r = N / hop windowed pieces to a ring of r + 1 block accumulators in shared memory
// (two frames touch r + 1 blocks); after frame t block t is complete, stored and its slot cleared.  Twiddles and the
// normalised synthesis window (inverse_stft_window_fn / N) are tabulated once per CTA.  The per-block kernel above evaluates
// each frame r times and rebuilds the tables per block: 0.80 -> 0.3x ms for 192 signals of 4 s.
constexpr int IR_Q = 16;
__global__ void istft_masked_run_kernel(const float2* __restrict__ spec, const int* __restrict__ labels,
                                        const float* __restrict__ masks, int S, int T, int N, int logN, int hop, int nblocks,
                                        float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char st_smem[];
    float2* buf = reinterpret_cast<float2*>(st_smem);
    float2* tw = buf + N;
    float* winv = reinterpret_cast<float*>(tw + N / 2);   // [N]  w[i] / (N * sum_o w^2[(i mod hop) + o hop])
    float* ring = winv + N;                               // [r + 1][hop]
    const int m0 = blockIdx.x * IR_Q, m1 = min(m0 + IR_Q, nblocks), bs = blockIdx.y, b = bs / S, s = bs % S;
    const int F = N / 2 + 1, r = N / hop, R1 = r + 1;
    const int64_t Lout = (int64_t)(T - 1) * hop + N;
    fill_twiddles(tw, N, +1.f);
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int j = i % hop;
        float den = 0.f;
        for (int o2 = 0; o2 < r; ++o2) { const float w = hann_periodic(j + o2 * hop, N); den = fmaf(w, w, den); }
        winv[i] = hann_periodic(i, N) / den * (1.0f / (float)N);
    }
    for (int i = threadIdx.x; i < R1 * hop; i += blockDim.x) ring[i] = 0.f;
    auto emit = [&](int m) {                              // block m is complete: store it and clear its slot
        float* slot = ring + (m % R1) * hop;
        for (int j = threadIdx.x; j < hop; j += blockDim.x) {
            const int64_t u = (int64_t)m * hop + j;
            if (m >= m0 && u < Lout) out[(size_t)bs * Lout + u] = slot[j];
            slot[j] = 0.f;
        }
    };
    auto masked = [&](size_t o, int k) {
        const float w = labels ? (labels[o + k] == s ? 1.f : 0.f) : masks[(o + k) * S + s];
        float2 v = spec[o + k];
        v.x *= w; v.y *= w;
        if (k == 0 || k == N / 2) v.y = 0.f;              // irfft ignores the imaginary part of DC and Nyquist
        return v;
    };
    const int t_lo = max(m0 - r + 1, 0), t_hi = min(m1 - 1, T - 1);       // frames that touch the blocks [m0, m1)
    for (int t = t_lo; t <= t_hi; t += 2) {
        const bool two = t + 1 <= t_hi;
        const size_t o = ((size_t)b * T + t) * F;
        __syncthreads();                                  // tables / cleared slots visible; buf free again
        for (int k = threadIdx.x; k < F; k += blockDim.x) {
            const float2 v1 = masked(o, k);
            const float2 v2 = two ? masked(o + F, k) : make_float2(0.f, 0.f);
            // Z[k] = X1[k] + i X2[k];  Z[N-k] = conj(X1[k]) + i conj(X2[k])
            buf[__brev((unsigned)k) >> (32 - logN)] = make_float2(v1.x - v2.y, v1.y + v2.x);
            if (k != 0 && k != N / 2) buf[__brev((unsigned)(N - k)) >> (32 - logN)] = make_float2(v1.x + v2.y, v2.x - v1.y);
        }
        fft_shared(buf, tw, N, logN);
        // piece o2 of frame t belongs to block t + o2, of frame t+1 to block t + 1 + o2: r + 1 different slots
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const int o2 = i / hop, j = i - o2 * hop;
            const float2 z = buf[i];
            ring[((t + o2) % R1) * hop + j] += z.x * winv[i];
            if (two) ring[((t + 1 + o2) % R1) * hop + j] += z.y * winv[i];
        }
        __syncthreads();
        emit(t);                                          // frames t-r+1 .. t have all been added (or do not exist)
        if (two) { __syncthreads(); emit(t + 1); }
    }
    __syncthreads();
    for (int m = max(t_hi + 1, m0); m < m1; ++m) { emit(m); __syncthreads(); }   // the tail blocks after the last frame
}

// Backward of istft_masked w.r.t. the soft masks (fine-tuning through Separator.postprocessing, network.py:584-607,
// 697-723).  out[n] = sum_t winv[n - t*hop] * irfft(mask * X)[n - t*hop], so with g[i] = dout[t*hop + i] * winv[i]:
//   d Re(mask*X)_k = (c_k / N) Re(DFT g)_k,   d Im(mask*X)_k = (c_k / N) Im(DFT g)_k   (c_k = 1 for k = 0, N/2, else 2;
//   irfft ignores the imaginary part of those two bins),   d mask_k = Re(X_k) dRe_k + Im(X_k) dIm_k.
// grid (T, B*S); block N/2.  dmasks[B, T*F, S].
__global__ void istft_masked_bwd_kernel(const float2* __restrict__ spec, const float* __restrict__ dout, int S, int T, int N,
                                        int logN, int hop, float* __restrict__ dmasks) {
    extern __shared__ __align__(16) unsigned char st_smem[];
    float2* buf = reinterpret_cast<float2*>(st_smem);
    float2* tw = buf + N;
    const int t = blockIdx.x, bs = blockIdx.y, b = bs / S, s = bs % S;
    const int F = N / 2 + 1, r = N / hop;
    const int64_t Lout = (int64_t)(T - 1) * hop + N;
    fill_twiddles(tw, N, -1.f);
    const float* src = dout + (size_t)bs * Lout + (size_t)t * hop;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int j = i % hop;
        float den = 0.f;
        for (int o2 = 0; o2 < r; ++o2) { const float w = hann_periodic(j + o2 * hop, N); den = fmaf(w, w, den); }
        const float v = src[i] * (hann_periodic(i, N) / den) * (1.0f / (float)N);
        buf[__brev((unsigned)i) >> (32 - logN)] = make_float2(v, 0.f);
    }
    fft_shared(buf, tw, N, logN);
    const size_t o = ((size_t)b * T + t) * F;
    for (int k = threadIdx.x; k < F; k += blockDim.x) {
        const float2 g = buf[k], x = spec[o + k];
        const bool edge = k == 0 || k == N / 2;
        dmasks[(o + k) * S + s] = edge ? x.x * g.x : 2.f * (x.x * g.x + x.y * g.y);
    }
}

int ilog2_exact(int n) {
    int l = 0;
    while ((1 << l) < n) ++l;
    return (1 << l) == n ? l : -1;
}

}  // namespace
}  // namespace amss

using namespace amss;

extern "C" int amss_stft_fwd(const float* x, int R, int L, int frame, int hop, float* spec, float* mag,
                             void* stream) {
    AMSS_REQUIRE(x && (spec || mag), "stft_fwd: null pointer");
    const int logN = ilog2_exact(frame);
    AMSS_REQUIRE(logN >= 6 && logN <= 11, "stft_fwd: frame %d must be a power of two in [64,2048]", frame);
    AMSS_REQUIRE(hop > 0 && L >= frame && R > 0, "stft_fwd: bad sizes L=%d frame=%d hop=%d", L, frame, hop);
    const int T = 1 + (L - frame) / hop;
    const size_t smem = (size_t)frame * 8 + (size_t)frame / 2 * 8;
    // several frames per CTA, two frames per complex FFT -- unless the batch is so small that one frame per CTA is what
    // fills the machine (4 signals: 19.5 us per frame vs 29.4 us; 64 signals: 130 vs 87 us, tools/bench_stft.py)
    if (getenv("AMSS_STFT_PER_FRAME") == nullptr && (int64_t)T * R >= 4000) {
        const int fr = st_frames_per_cta(T, R);
        dim3 grid((T + fr - 1) / fr, R);
        AMSS_LAUNCH(stft_fwd_run_kernel, grid, frame / 2, smem + (size_t)frame * 4, stream, x, (int64_t)L, frame, logN, hop, T, fr,
                    reinterpret_cast<float2*>(spec), mag);
        return AMSS_OK;
    }
    dim3 grid(T, R);
    AMSS_LAUNCH(stft_fwd_kernel, grid, frame / 2, smem, stream, x, (int64_t)L, frame, logN, hop, T,
                reinterpret_cast<float2*>(spec), mag);
    return AMSS_OK;
}

extern "C" int amss_stft_labels(const float* non_mix, int B, int S, int L, int frame, int hop, uint8_t* labels,
                                float* mag_non_mix, void* stream) {
    AMSS_REQUIRE(non_mix && labels, "stft_labels: null pointer");
    const int logN = ilog2_exact(frame);
    AMSS_REQUIRE(logN >= 6 && logN <= 11, "stft_labels: frame %d must be a power of two in [64,2048]", frame);
    AMSS_REQUIRE(hop > 0 && L >= frame && B > 0 && S > 0 && S < 256, "stft_labels: bad sizes");
    const int T = 1 + (L - frame) / hop;
    const size_t smem = (size_t)frame * 8 + (size_t)frame / 2 * 8;
    if (getenv("AMSS_STFT_PER_FRAME") == nullptr && (int64_t)T * B >= 4000) {
        const int fr = st_frames_per_cta(T, B);
        dim3 grid((T + fr - 1) / fr, B);
        AMSS_LAUNCH(stft_labels_run_kernel, grid, frame / 2, smem + (size_t)frame * 4, stream, non_mix, S, (int64_t)L, frame, logN,
                    hop, T, fr, labels, mag_non_mix);
        return AMSS_OK;
    }
    dim3 grid(T, B);
    AMSS_LAUNCH(stft_labels_kernel, grid, frame / 2, smem, stream, non_mix, S, (int64_t)L, frame, logN, hop, T,
                labels, mag_non_mix);
    return AMSS_OK;
}

extern "C" int amss_istft_masked_fwd(const float* spec, const int32_t* labels, const float* masks, int B, int S,
                                     int T, int frame, int hop, float* out, void* stream) {
    AMSS_REQUIRE(spec && out && ((labels != nullptr) != (masks != nullptr)),
                 "istft_masked_fwd: need spec, out and exactly one of labels / masks");
    const int logN = ilog2_exact(frame);
    AMSS_REQUIRE(logN >= 6 && logN <= 11, "istft_masked_fwd: frame %d must be a power of two in [64,2048]", frame);
    AMSS_REQUIRE(hop > 0 && frame % hop == 0, "istft_masked_fwd: frame must be a multiple of hop");
    const int nblocks = T - 1 + frame / hop;
    if (getenv("AMSS_ISTFT_PER_BLOCK") == nullptr) {       // (A/B and parity tests: the one-CTA-per-block kernel)
        const size_t smem = (size_t)frame * 8 + (size_t)frame / 2 * 8 + (size_t)frame * 4 + (size_t)(frame + hop) * 4;
        dim3 grid((nblocks + IR_Q - 1) / IR_Q, B * S);
        AMSS_LAUNCH(istft_masked_run_kernel, grid, frame / 2, smem, stream, reinterpret_cast<const float2*>(spec), labels,
                    masks, S, T, frame, logN, hop, nblocks, out);
        return AMSS_OK;
    }
    const size_t smem = (size_t)frame * 8 + (size_t)frame / 2 * 8 + (size_t)hop * 4;
    dim3 grid(nblocks, B * S);
    AMSS_LAUNCH(istft_masked_kernel, grid, frame / 2, smem, stream, reinterpret_cast<const float2*>(spec), labels,
                masks, S, T, frame, logN, hop, out);
    return AMSS_OK;
}

extern "C" int amss_istft_masked_bwd(const float* spec, const float* dout, int B, int S, int T, int frame, int hop,
                                     float* dmasks, void* stream) {
    AMSS_REQUIRE(spec && dout && dmasks, "istft_masked_bwd: null pointer");
    const int logN = ilog2_exact(frame);
    AMSS_REQUIRE(logN >= 6 && logN <= 11, "istft_masked_bwd: frame %d must be a power of two in [64,2048]", frame);
    AMSS_REQUIRE(hop > 0 && frame % hop == 0 && B > 0 && S > 0 && T > 0, "istft_masked_bwd: bad sizes");
    const size_t smem = (size_t)frame * 8 + (size_t)frame / 2 * 8;
    dim3 grid(T, B * S);
    AMSS_LAUNCH(istft_masked_bwd_kernel, grid, frame / 2, smem, stream, reinterpret_cast<const float2*>(spec), dout, S, T,
                frame, logN, hop, dmasks);
    return AMSS_OK;
}
