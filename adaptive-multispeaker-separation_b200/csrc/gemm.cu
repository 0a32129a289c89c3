// Dense GEMM entry point.  C[M,N] (ldc) = op(A) op(B) (+ bias[N]) (+ C if accumulate).
// Replaces tf.nn.conv1d(k=1) (utils/ops.py:501-503), the hoisted input projection of
// BasicLSTMCell (utils/ops.py:372-380) and their autograd transposes.
//   AMSS_PREC_FP32 : fp32 SIMT tiles (128x128x16, 8x8 per thread)  -- the parity path
//   AMSS_PREC_BF16 : tcgen05 bf16 tiles with TMEM accumulators (gemm_tc.cu) when the shape is
//                    supported, else an error (never a silent fallback).
#include "common.cuh"
#include <algorithm>
#include <cuda_bf16.h>

namespace amss {
bool gemm_tc_supported(int M, int N, int K, int lda, int ldb, int ldc, int transa, int transb);
size_t gemm_tc_workspace(int M, int N, int K, int transa, int transb, int precision);
int gemm_tc(const float* A, int lda, const float* B, int ldb, const float* bias, int M, int N, int K, int transa,
            int transb, int accumulate, int precision, float* C, int ldc, int swapB, int swapT, void* workspace,
            size_t workspace_bytes, cudaStream_t st);

int convert_bf16(const float* src, int rows, int cols, int ld, uint16_t* dst, int ldd, cudaStream_t st);
int gemm_bf16(const uint16_t* A, int lda, int a_mn, const uint16_t* B, int ldb, int b_mn, const float* bias, int M, int N,
              int K, int accumulate, float* C, int ldc, int swapB, int swapT, int norm_E, float* inv, cudaStream_t st);

namespace {

constexpr int GM_BM = 128, GM_BN = 128, GM_BK = 16, GM_THREADS = 256;

// swapB/swapT > 0: output row m (= t*swapB + b, time-major) is written to row b*swapT + t.
template <int TA, int TB>
__global__ void __launch_bounds__(GM_THREADS)
sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
             const float* __restrict__ bias, int M, int N, int K, int accumulate, float* __restrict__ C, int ldc,
             int swapB, int swapT) {
    __shared__ __align__(16) float As[GM_BK][GM_BM + 4];
    __shared__ __align__(16) float Bs[GM_BK][GM_BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * GM_BM, n0 = blockIdx.x * GM_BN;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += GM_BK) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int idx = tid + it * GM_THREADS;
            int m, k;
            if (TA) { m = idx & 127; k = idx >> 7; } else { k = idx & 15; m = idx >> 4; }
            const int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < M && gk < K) v = TA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
            As[k][m] = v;
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int idx = tid + it * GM_THREADS;
            int n, k;
            if (TB) { k = idx & 15; n = idx >> 4; } else { n = idx & 127; k = idx >> 7; }
            const int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < N && gk < K) v = TB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GM_BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
        if (m >= M) continue;
        size_t row = (size_t)m;
        if (swapB > 0) row = (size_t)(m % swapB) * swapT + (size_t)(m / swapB);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
            if (n >= N) continue;
            float v = acc[i][j];
            if (bias) v += bias[n];
            float* c = C + row * ldc + n;
            *c = accumulate ? (*c + v) : v;
        }
    }
}

__global__ void transpose_01_kernel(const float* __restrict__ in, int D0, int D1, int C, float* __restrict__ out) {
    // in[D0][D1][C] -> out[D1][D0][C]
    const int64_t n = (int64_t)D0 * D1 * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t r = i / C;
        const int d1 = (int)(r % D1), d0 = (int)(r / D1);
        out[((size_t)d1 * D0 + d0) * C + c] = in[i];
    }
}

// the same in 16-byte units (C % 4 == 0, aligned buffers, fewer than 2^31 units: 32-bit index arithmetic)
__global__ void transpose_01_vec_kernel(const float4* __restrict__ in, int D0, int D1, int C4, float4* __restrict__ out) {
    const uint32_t n = (uint32_t)D0 * (uint32_t)D1 * (uint32_t)C4;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t c = i % (uint32_t)C4, r = i / (uint32_t)C4;
        const uint32_t d1 = r % (uint32_t)D1, d0 = r / (uint32_t)D1;
        out[((size_t)d1 * D0 + d0) * C4 + c] = __ldg(in + i);
    }
}

// in[D0][D1][C] fp32 -> out[D1][D0][ldd] bf16 (ldd = C padded to 8, zero filled): the [T,B,C] -> [B,T,C] hand-over into
// the embedding head fused with its operand conversion (one pass instead of transpose + convert)
__global__ void transpose_01_bf16_kernel(const float* __restrict__ in, int D0, int D1, int C, int ldd, uint4* __restrict__ out) {
    const int upr = ldd >> 3;
    const int64_t units = (int64_t)D0 * D1 * upr;
    const bool vec_ok = (C & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < units; u += (int64_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(u % upr) * 8;
        const int64_t r = u / upr;                       // output row d1*D0 + d0
        const int d0 = (int)(r % D0), d1 = (int)(r / D0);
        const float* s = in + ((size_t)d0 * D1 + d1) * C + c0;
        float v[8];
        if (c0 + 8 <= C && vec_ok) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(s)), b = __ldg(reinterpret_cast<const float4*>(s) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = c0 + e < C ? __ldg(s + e) : 0.f;
        }
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
        out[u] = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                            *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
    }
}

}  // namespace

int sgemm_launch(const float* A, int lda, const float* B, int ldb, const float* bias, int M, int N, int K, int transa,
                 int transb, int accumulate, float* C, int ldc, int swapB, int swapT, cudaStream_t st) {
    dim3 grid((N + GM_BN - 1) / GM_BN, (M + GM_BM - 1) / GM_BM);
    if (!transa && !transb) {
        AMSS_LAUNCH((sgemm_kernel<0, 0>), grid, GM_THREADS, 0, st, A, lda, B, ldb, bias, M, N, K, accumulate, C, ldc, swapB, swapT);
    } else if (transa && !transb) {
        AMSS_LAUNCH((sgemm_kernel<1, 0>), grid, GM_THREADS, 0, st, A, lda, B, ldb, bias, M, N, K, accumulate, C, ldc, swapB, swapT);
    } else if (!transa && transb) {
        AMSS_LAUNCH((sgemm_kernel<0, 1>), grid, GM_THREADS, 0, st, A, lda, B, ldb, bias, M, N, K, accumulate, C, ldc, swapB, swapT);
    } else {
        AMSS_LAUNCH((sgemm_kernel<1, 1>), grid, GM_THREADS, 0, st, A, lda, B, ldb, bias, M, N, K, accumulate, C, ldc, swapB, swapT);
    }
    return AMSS_OK;
}

// Internal GEMM used by the BLSTM and head code: picks the tensor-core path for bf16.
int gemm_dispatch(const float* A, int lda, const float* B, int ldb, const float* bias, int M, int N, int K, int transa,
                  int transb, int accumulate, int precision, float* C, int ldc, int swapB, int swapT, void* workspace,
                  size_t workspace_bytes, cudaStream_t st) {
    if (precision == AMSS_PREC_BF16) {
        if (!gemm_tc_supported(M, N, K, lda, ldb, ldc, transa, transb)) {
            set_error("gemm: bf16 tensor-core path does not support M=%d N=%d K=%d lda=%d ldb=%d ldc=%d ta=%d tb=%d",
                      M, N, K, lda, ldb, ldc, transa, transb);
            return AMSS_ERR_UNSUPPORTED;
        }
        return gemm_tc(A, lda, B, ldb, bias, M, N, K, transa, transb, accumulate, precision, C, ldc, swapB, swapT,
                       workspace, workspace_bytes, st);
    }
    return sgemm_launch(A, lda, B, ldb, bias, M, N, K, transa, transb, accumulate, C, ldc, swapB, swapT, st);
}

int transpose_01(const float* in, int D0, int D1, int C, float* out, cudaStream_t st) {
    const int64_t n4 = (int64_t)D0 * D1 * (C / 4);
    if ((C & 3) == 0 && n4 < (1ll << 31) && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        const int grid = (int)std::min<int64_t>((n4 + 255) / 256, 16 * kNumSMs);
        AMSS_LAUNCH(transpose_01_vec_kernel, grid, 256, 0, st, reinterpret_cast<const float4*>(in), D0, D1, C / 4,
                    reinterpret_cast<float4*>(out));
        return AMSS_OK;
    }
    AMSS_LAUNCH(transpose_01_kernel, 4 * kNumSMs, 256, 0, st, in, D0, D1, C, out);
    return AMSS_OK;
}

}  // namespace amss

using namespace amss;

extern "C" size_t amss_gemm_workspace_bytes(int M, int N, int K, int transa, int transb, int precision) {
    if (precision == AMSS_PREC_BF16) return gemm_tc_workspace(M, N, K, transa, transb, precision);
    return 256;
}

extern "C" int amss_gemm(const float* A, int lda, const float* B, int ldb, const float* bias, int M, int N, int K,
                         int transa, int transb, int accumulate, int precision, float* C, int ldc, int out_swap_b,
                         int out_swap_t, void* workspace, size_t workspace_bytes, void* stream) {
    AMSS_REQUIRE(A && B && C, "gemm: null pointer");
    AMSS_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: bad sizes M=%d N=%d K=%d", M, N, K);
    AMSS_REQUIRE(lda >= (transa ? M : K) && ldb >= (transb ? K : N) && ldc >= N, "gemm: leading dimension too small");
    AMSS_REQUIRE((out_swap_b > 0) == (out_swap_t > 0), "gemm: out_swap_b/out_swap_t must both be set or both 0");
    AMSS_REQUIRE(out_swap_b == 0 || (int64_t)out_swap_b * out_swap_t == M, "gemm: out_swap_b*out_swap_t != M");
    return gemm_dispatch(A, lda, B, ldb, bias, M, N, K, transa, transb, accumulate, precision, C, ldc, out_swap_b,
                         out_swap_t, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int amss_transpose_01(const float* in, int D0, int D1, int C, float* out, void* stream) {
    AMSS_REQUIRE(in && out && D0 > 0 && D1 > 0 && C > 0, "transpose_01: bad arguments");
    return transpose_01(in, D0, D1, C, out, (cudaStream_t)stream);
}

extern "C" int amss_convert_bf16(const float* src, int rows, int cols, int ld, uint16_t* dst, int ldd, void* stream) {
    AMSS_REQUIRE(src && dst && rows > 0 && cols > 0, "convert_bf16: bad arguments");
    AMSS_REQUIRE(ld >= cols && ldd >= cols && ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                 "convert_bf16: ldd must be a multiple of 8 (>= cols) and dst 16-byte aligned");
    return convert_bf16(src, rows, cols, ld, dst, ldd, (cudaStream_t)stream);
}

extern "C" int amss_gemm_bf16(const uint16_t* A, int lda, int a_mn, const uint16_t* B, int ldb, int b_mn, const float* bias,
                              int M, int N, int K, int accumulate, float* C, int ldc, int out_swap_b, int out_swap_t,
                              int norm_E, float* inv_norm, void* stream) {
    AMSS_REQUIRE(A && B && C, "gemm_bf16: null pointer");
    AMSS_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_bf16: bad sizes M=%d N=%d K=%d", M, N, K);
    AMSS_REQUIRE(lda >= (a_mn ? M : K) && ldb >= (b_mn ? N : K) && ldc >= N, "gemm_bf16: leading dimension too small");
    AMSS_REQUIRE((out_swap_b > 0) == (out_swap_t > 0), "gemm_bf16: out_swap_b/out_swap_t must both be set or both 0");
    AMSS_REQUIRE(out_swap_b == 0 || (int64_t)out_swap_b * out_swap_t == M, "gemm_bf16: out_swap_b*out_swap_t != M");
    return gemm_bf16(A, lda, a_mn ? 1 : 0, B, ldb, b_mn ? 1 : 0, bias, M, N, K, accumulate, C, ldc, out_swap_b,
                     out_swap_t, norm_E, inv_norm, (cudaStream_t)stream);
}

extern "C" int amss_transpose_01_bf16(const float* in, int D0, int D1, int C, uint16_t* out, int ldd, void* stream) {
    AMSS_REQUIRE(in && out && D0 > 0 && D1 > 0 && C > 0, "transpose_01_bf16: bad arguments");
    AMSS_REQUIRE(ldd >= C && ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "transpose_01_bf16: ldd must be a multiple of 8 (>= C) and out 16-byte aligned");
    AMSS_LAUNCH(transpose_01_bf16_kernel, 8 * kNumSMs, 256, 0, stream, in, D0, D1, C, ldd, (uint4*)out);
    return AMSS_OK;
}
