// Bidirectional LSTM layer (utils/ops.py:358-383): BasicLSTMCell(H) forward on x and on
// reverse(x), outputs concatenated.  Gate order i, j(candidate), f, o;
//   c' = c*sigmoid(f + forget_bias) + sigmoid(i)*tanh(j);  h' = tanh(c')*sigmoid(o).
//
// Structure (everything time-major, [T, B, *]):
//   1. hoisted input projection  Zx_d = x @ kernel_d[:I] + bias_d      one GEMM per direction
//   2. persistent recurrent kernel: both directions in ONE cooperative launch.  The 4H gate
//      columns of a direction are split over NC CTAs; a CTA keeps its W_h slice resident in
//      shared memory for all T steps, and the CTAs of a direction meet at a device-scope barrier
//      once per step (h is exchanged through y itself, which lives in L2).
//   3. backward: reverse-time persistent kernel producing dZ (pre-activation gate gradients),
//      then dW_x, dW_h, dbias, dx as GEMMs / column sums over all T*B rows at once.
// This file is the fp32 SIMT recurrent path (parity path); the GEMMs go through gemm_dispatch().
#include "common.cuh"
#include <algorithm>

namespace amss {
int gemm_dispatch(const float* A, int lda, const float* B, int ldb, const float* bias, int M, int N, int K, int transa,
                  int transb, int accumulate, int precision, float* C, int ldc, int swapB, int swapT, void* workspace,
                  size_t workspace_bytes, cudaStream_t st);
bool blstm_rec_tc_supported(int B, int T, int H);
void blstm_tc_set_profile(long long* dev_buf);
void blstm_tc_set_sched(long long* dev_buf);
void blstm_tc_max_clusters(int H, int* out4);
int convert_bf16(const float* src, int rows, int cols, int ld, uint16_t* dst, int ldd, cudaStream_t st);
int gemm_bf16(const uint16_t* A, int lda, int a_mn, const uint16_t* B, int ldb, int b_mn, const float* bias, int M, int N,
              int K, int accumulate, float* C, int ldc, int swapB, int swapT, int norm_E, float* inv, cudaStream_t st);
int blstm_rec_fwd_tc(const float* Wh_fw, const float* Wh_bw, int ldw, float* gates, float* cst, float* y, int B, int T,
                     int H, float forget_bias, cudaStream_t st);
int blstm_rec_bwd_tc_nsub(int B, int H);
int blstm_rec_bwd_tc(const float* Wh_fw, const float* Wh_bw, int ldw, const float* gates, const float* cst,
                     const float* dy, float* dZ, uint16_t* dZb, int ldzb, float* dbpart, int B, int T, int H, cudaStream_t st);
namespace {

constexpr int RC_THREADS = 256;
constexpr int RC_CCH = 256;   // dZ columns staged per chunk in the backward kernel

struct RecGeom {
    int U, NC, BC, HP;
    size_t smem_fwd, smem_bwd;
};
RecGeom rec_geom(int B, int H) {
    RecGeom g;
    int nc = std::min(H, kNumSMs / 2);
    g.U = (H + nc - 1) / nc;
    g.NC = (H + g.U - 1) / g.U;
    g.BC = std::max(1, std::min(B, RC_THREADS / g.U));
    g.HP = H | 1;   // odd row pitch: conflict-free column walks
    g.smem_fwd = ((size_t)H * g.U * 4 + (size_t)g.BC * g.HP) * 4;
    g.smem_bwd = ((size_t)g.U * 4 * H + (size_t)g.BC * (RC_CCH + 1)) * 4;
    return g;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Device-scope barrier among the `n` CTAs of one direction; `target` = n * (#barriers so far).
__device__ __forceinline__ void dir_barrier(unsigned* ctr, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (ld_acquire_u32(ctr) < target) { }
        __threadfence();
    }
    __syncthreads();
}

struct RecFwdParams {
    const float* Wh[2];   // [H][ldw] recurrent rows of the TF kernel, per direction
    int ldw;
    float* gates;         // [2][T][B][4H]: in = Zx (pre-activation input part), out = activated gates
    float* cst;           // [2][T][B][H]
    float* y;             // [T][B][2H]
    unsigned* bar;        // [2]
    int B, T, H, U, NC, BC, HP;
    float forget_bias;
};

__global__ void __launch_bounds__(RC_THREADS, 1) blstm_rec_fwd_kernel(RecFwdParams p) {
    extern __shared__ __align__(16) unsigned char rc_smem[];
    const int H = p.H, U = p.U, B = p.B, T = p.T, HP = p.HP;
    float4* Ws = reinterpret_cast<float4*>(rc_smem);                   // [H][U] of (i,j,f,o)
    float* hs = reinterpret_cast<float*>(Ws + (size_t)H * U);          // [BC][HP]
    const int d = blockIdx.x / p.NC, ci = blockIdx.x % p.NC;
    const int u0 = ci * U, nu = min(U, H - u0);
    const int tid = threadIdx.x;
    const float* Wh = p.Wh[d];
    for (int idx = tid; idx < H * U; idx += RC_THREADS) {
        const int k = idx / U, u = idx - k * U;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (u < nu) {
            const float* r = Wh + (size_t)k * p.ldw + u0 + u;
            w = make_float4(r[0], r[H], r[2 * H], r[3 * H]);
        }
        Ws[idx] = w;
    }
    __syncthreads();
    for (int s = 0; s < T; ++s) {
        const int t = d == 0 ? s : T - 1 - s;
        const int tprev = d == 0 ? t - 1 : t + 1;
        for (int b0 = 0; b0 < B; b0 += p.BC) {
            const int nb = min(p.BC, B - b0);
            __syncthreads();
            if (s > 0) {
                for (int idx = tid; idx < nb * H; idx += RC_THREADS) {
                    const int bl = idx / H, k = idx - bl * H;
                    hs[bl * HP + k] = __ldcg(p.y + ((size_t)tprev * B + b0 + bl) * 2 * H + d * H + k);
                }
            }
            __syncthreads();
            const int bl = tid / U, u = tid - bl * U;
            if (bl < nb && u < nu) {
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                if (s > 0) {
                    const float* hrow = hs + bl * HP;
                    for (int k = 0; k < H; ++k) {
                        const float hv = hrow[k];
                        const float4 w = Ws[k * U + u];
                        a0 = fmaf(hv, w.x, a0); a1 = fmaf(hv, w.y, a1); a2 = fmaf(hv, w.z, a2); a3 = fmaf(hv, w.w, a3);
                    }
                }
                const int b = b0 + bl, hu = u0 + u;
                float* g = p.gates + (((size_t)d * T + t) * B + b) * 4 * H + hu;
                const float zi = a0 + g[0], zj = a1 + g[H], zf = a2 + g[2 * H], zo = a3 + g[3 * H];
                const float cprev = s > 0 ? p.cst[(((size_t)d * T + tprev) * B + b) * H + hu] : 0.f;
                const float gi = 1.f / (1.f + expf(-zi));
                const float gj = tanhf(zj);
                const float gf = 1.f / (1.f + expf(-(zf + p.forget_bias)));
                const float go = 1.f / (1.f + expf(-zo));
                const float c = cprev * gf + gi * gj;
                const float h = tanhf(c) * go;
                g[0] = gi; g[H] = gj; g[2 * H] = gf; g[3 * H] = go;
                p.cst[(((size_t)d * T + t) * B + b) * H + hu] = c;
                p.y[((size_t)t * B + b) * 2 * H + d * H + hu] = h;
            }
        }
        if (s + 1 < T) dir_barrier(p.bar + d, (unsigned)p.NC * (unsigned)(s + 1));
    }
}

struct RecBwdParams {
    const float* Wh[2];
    int ldw;
    const float* gates;   // activated gates from the forward pass
    const float* cst;
    const float* dy;      // [T][B][2H]
    float* dZ;            // [2][T][B][4H]
    float* dcc;           // [2][B][H] cell-gradient carry
    unsigned* bar;
    int B, T, H, U, NC, BC;
};

__global__ void __launch_bounds__(RC_THREADS, 1) blstm_rec_bwd_kernel(RecBwdParams p) {
    extern __shared__ __align__(16) unsigned char rc_smem[];
    const int H = p.H, U = p.U, B = p.B, T = p.T, H4 = 4 * p.H;
    float* Wt = reinterpret_cast<float*>(rc_smem);                     // [U][4H]: Wh[u0+u][:]
    float* dzs = Wt + (size_t)U * H4;                                  // [BC][RC_CCH+1]
    const int d = blockIdx.x / p.NC, ci = blockIdx.x % p.NC;
    const int u0 = ci * U, nu = min(U, H - u0);
    const int tid = threadIdx.x;
    const float* Wh = p.Wh[d];
    for (int idx = tid; idx < U * H4; idx += RC_THREADS) {
        const int u = idx / H4, c = idx - u * H4;
        Wt[idx] = u < nu ? Wh[(size_t)(u0 + u) * p.ldw + c] : 0.f;
    }
    __syncthreads();
    for (int s = T - 1; s >= 0; --s) {
        const int t = d == 0 ? s : T - 1 - s;
        const int tprev = d == 0 ? t - 1 : t + 1;    // forward-order predecessor
        const int tnext = d == 0 ? t + 1 : t - 1;    // forward-order successor (already processed)
        for (int b0 = 0; b0 < B; b0 += p.BC) {
            const int nb = min(p.BC, B - b0);
            const int bl = tid / U, u = tid - bl * U;
            const bool has = bl < nb && u < nu;
            float acc = 0.f;
            if (s < T - 1) {
                for (int c0 = 0; c0 < H4; c0 += RC_CCH) {
                    const int nc = min(RC_CCH, H4 - c0);
                    __syncthreads();
                    for (int idx = tid; idx < nb * nc; idx += RC_THREADS) {
                        const int r = idx / nc, c = idx - r * nc;
                        dzs[r * (RC_CCH + 1) + c] =
                            __ldcg(p.dZ + (((size_t)d * T + tnext) * B + b0 + r) * H4 + c0 + c);
                    }
                    __syncthreads();
                    if (has) {
                        const float* dr = dzs + bl * (RC_CCH + 1);
                        const float* wr = Wt + (size_t)u * H4 + c0;
                        for (int c = 0; c < nc; ++c) acc = fmaf(dr[c], wr[c], acc);
                    }
                }
            }
            if (has) {
                const int b = b0 + bl, hu = u0 + u;
                const float dh = p.dy[((size_t)t * B + b) * 2 * H + d * H + hu] + acc;
                const float* g = p.gates + (((size_t)d * T + t) * B + b) * H4 + hu;
                const float gi = g[0], gj = g[H], gf = g[2 * H], go = g[3 * H];
                const float c = p.cst[(((size_t)d * T + t) * B + b) * H + hu];
                const float cprev = s > 0 ? p.cst[(((size_t)d * T + tprev) * B + b) * H + hu] : 0.f;
                const float tc = tanhf(c);
                float* dcp = p.dcc + ((size_t)d * B + b) * H + hu;
                const float dc = (s < T - 1 ? *dcp : 0.f) + dh * go * (1.f - tc * tc);
                float* dz = p.dZ + (((size_t)d * T + t) * B + b) * H4 + hu;
                dz[0] = dc * gj * gi * (1.f - gi);
                dz[H] = dc * gi * (1.f - gj * gj);
                dz[2 * H] = dc * cprev * gf * (1.f - gf);
                dz[3 * H] = dh * tc * go * (1.f - go);
                *dcp = dc * gf;
            }
        }
        if (s > 0) dir_barrier(p.bar + d, (unsigned)p.NC * (unsigned)(T - s));
    }
}

// dbias[n] = sum_m Z[m][n], deterministic two-level sum: grid (ceil(N/32), CS_CHUNKS) -> part[chunk][N] -> out[N]
constexpr int CS_CHUNKS = 64;
__global__ void colsum_chunk_kernel(const float* __restrict__ Z, int64_t M, int N, float* __restrict__ part) {
    __shared__ float tile[8][33];
    const int n = blockIdx.x * 32 + (threadIdx.x & 31), ty = threadIdx.x >> 5;
    float a = 0.f;
    if (n < N) for (int64_t m = blockIdx.y * 8 + ty; m < M; m += (int64_t)gridDim.y * 8) a += Z[m * N + n];
    tile[ty][threadIdx.x & 31] = a;
    __syncthreads();
    if (ty == 0 && n < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += tile[i][threadIdx.x & 31];
        part[(size_t)blockIdx.y * N + n] = s;
    }
}
__global__ void colsum_sum_kernel(const float* __restrict__ part, int chunks, int N, float* __restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) {
        float s = 0.f;
        for (int c = 0; c < chunks; ++c) s += part[(size_t)c * N + n];
        out[n] = s;
    }
}

struct BlstmWs {
    float *gates, *cst, *dZ, *dcc;
    unsigned* bar;
    void* gemm_ws;
    size_t gemm_ws_bytes, total;
};

// bf16 operand copies for the tensor-core GEMMs (leading dimensions padded to 8 elements for TMA), carved out of the
// GEMM scratch after the column-sum partials: x [TB,Ip], y per direction [2][TB,Hp], dZ [2*TB,H4p], W_x [2][I,H4p].
inline int pad8i(int x) { return (x + 7) & ~7; }
struct Bf16Scratch {
    uint16_t *xb, *yb[2], *dzb, *wb[2];
    int Ip, Hp, H4p;
    size_t bytes;
};
Bf16Scratch bf16_scratch(void* base, int B, int T, int I, int H) {
    Bf16Scratch s;
    s.Ip = pad8i(I); s.Hp = pad8i(H); s.H4p = pad8i(4 * H);
    const size_t TB = (size_t)T * B;
    char* p = (char*)base + align_up((size_t)CS_CHUNKS * 4 * H * 4, 256);
    s.xb = (uint16_t*)p;  p += align_up(TB * s.Ip * 2, 256);
    for (int d = 0; d < 2; ++d) { s.yb[d] = (uint16_t*)p; p += align_up(TB * s.Hp * 2, 256); }
    s.dzb = (uint16_t*)p; p += align_up(2 * TB * s.H4p * 2, 256);
    for (int d = 0; d < 2; ++d) { s.wb[d] = (uint16_t*)p; p += align_up((size_t)I * s.H4p * 2, 256); }
    s.bytes = (size_t)(p - (char*)base);
    return s;
}

size_t gates_bytes(int B, int T, int H) { return align_up((size_t)2 * T * B * 4 * H * 4, 256); }
size_t cst_bytes(int B, int T, int H) { return align_up((size_t)2 * T * B * H * 4, 256); }

}  // namespace
}  // namespace amss

using namespace amss;

extern "C" size_t amss_gemm_workspace_bytes(int M, int N, int K, int transa, int transb, int precision);

// saved = [activated gates][cell states][bf16 copy of x, [T*B, pad8(I)] -- written and read on the tensor-core path only]
extern "C" size_t amss_blstm_saved_bytes(int B, int T, int I, int H) {
    return gates_bytes(B, T, H) + cst_bytes(B, T, H) + align_up((size_t)T * B * pad8i(I) * 2, 256);
}

extern "C" size_t amss_blstm_workspace_bytes(int B, int T, int I, int H, int precision) {
    // forward without `saved` needs gates+cst; backward needs dZ + carry; both need counters + GEMM scratch
    size_t g = 0;
    g = std::max(g, amss_gemm_workspace_bytes(T * B, 4 * H, I, 0, 0, precision));
    g = std::max(g, amss_gemm_workspace_bytes(I, 4 * H, T * B, 1, 0, precision));
    g = std::max(g, amss_gemm_workspace_bytes(H, 4 * H, T * B, 1, 0, precision));
    g = std::max(g, amss_gemm_workspace_bytes(T * B, I, 4 * H, 0, 1, precision));
    g = std::max(g, (size_t)CS_CHUNKS * 4 * H * 4);   // column-sum partials share the GEMM scratch
    if (precision == AMSS_PREC_BF16) g = std::max(g, bf16_scratch(nullptr, B, T, I, H).bytes);
    return 256 + gates_bytes(B, T, H) + cst_bytes(B, T, H) + align_up((size_t)2 * B * H * 4, 256) + align_up(g, 256);
}

extern "C" int amss_blstm_fwd(const float* x, const float* kernel_fw, const float* bias_fw, const float* kernel_bw,
                              const float* bias_bw, int B, int T, int I, int H, float forget_bias, int precision,
                              float* y, void* saved, void* workspace, size_t workspace_bytes, void* stream) {
    AMSS_REQUIRE(x && kernel_fw && kernel_bw && bias_fw && bias_bw && y && workspace, "blstm_fwd: null pointer");
    AMSS_REQUIRE(B > 0 && T > 0 && I > 0 && H > 0, "blstm_fwd: bad sizes");
    if (workspace_bytes < amss_blstm_workspace_bytes(B, T, I, H, precision)) { set_error("blstm_fwd: workspace too small"); return AMSS_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    const RecGeom geo = rec_geom(B, H);
    AMSS_REQUIRE(geo.smem_fwd <= 220 * 1024, "blstm_fwd: H=%d needs %zu B of shared memory", H, geo.smem_fwd);
    char* ws = (char*)workspace;
    unsigned* bar = (unsigned*)ws;
    char* after = ws + 256;
    float* gates = saved ? (float*)saved : (float*)after;
    float* cst = saved ? (float*)((char*)saved + gates_bytes(B, T, H)) : (float*)(after + gates_bytes(B, T, H));
    void* gws = after + gates_bytes(B, T, H) + cst_bytes(B, T, H) + align_up((size_t)2 * B * H * 4, 256);
    const size_t gws_bytes = workspace_bytes - (size_t)((char*)gws - ws);
    const float* kern[2] = {kernel_fw, kernel_bw};
    const float* bias[2] = {bias_fw, bias_bw};
    if (precision == AMSS_PREC_BF16) {
        // hoisted input projection on the tensor cores: x is converted once for both directions
        // (the copy goes into `saved` when there is one: the backward pass reads it instead of converting x again)
        const Bf16Scratch s = bf16_scratch(gws, B, T, I, H);
        uint16_t* xb = saved ? (uint16_t*)((char*)saved + gates_bytes(B, T, H) + cst_bytes(B, T, H)) : s.xb;
        int rc = convert_bf16(x, T * B, I, I, xb, s.Ip, st);
        if (rc != AMSS_OK) return rc;
        for (int d = 0; d < 2; ++d) {
            rc = convert_bf16(kern[d], I, 4 * H, 4 * H, s.wb[d], s.H4p, st);
            if (rc != AMSS_OK) return rc;
            rc = gemm_bf16(xb, s.Ip, 0, s.wb[d], s.H4p, 1, bias[d], T * B, 4 * H, I, 0,
                           gates + (size_t)d * T * B * 4 * H, 4 * H, 0, 0, 0, nullptr, st);
            if (rc != AMSS_OK) return rc;
        }
    } else {
        for (int d = 0; d < 2; ++d) {
            int rc = gemm_dispatch(x, I, kern[d], 4 * H, bias[d], T * B, 4 * H, I, 0, 0, 0, precision,
                                   gates + (size_t)d * T * B * 4 * H, 4 * H, 0, 0, gws, gws_bytes, st);
            if (rc != AMSS_OK) return rc;
        }
    }
    if (precision == AMSS_PREC_BF16 && blstm_rec_tc_supported(B, T, H))
        return blstm_rec_fwd_tc(kernel_fw + (size_t)I * 4 * H, kernel_bw + (size_t)I * 4 * H, 4 * H, gates, cst, y, B, T, H,
                                forget_bias, st);
    AMSS_CUDA(cudaMemsetAsync(bar, 0, 256, st));
    RecFwdParams p;
    p.Wh[0] = kernel_fw + (size_t)I * 4 * H;
    p.Wh[1] = kernel_bw + (size_t)I * 4 * H;
    p.ldw = 4 * H; p.gates = gates; p.cst = cst; p.y = y; p.bar = bar;
    p.B = B; p.T = T; p.H = H; p.U = geo.U; p.NC = geo.NC; p.BC = geo.BC; p.HP = geo.HP;
    p.forget_bias = forget_bias;
    AMSS_CUDA(cudaFuncSetAttribute(blstm_rec_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)geo.smem_fwd));
    void* args[] = {&p};
    AMSS_CUDA(cudaLaunchCooperativeKernel((void*)blstm_rec_fwd_kernel, dim3(2 * geo.NC), dim3(RC_THREADS), args,
                                          geo.smem_fwd, st));
    count_launch();
    return AMSS_OK;
}

extern "C" int amss_blstm_bwd(const float* x, const float* kernel_fw, const float* kernel_bw, const float* y,
                              const float* dy, const void* saved, int B, int T, int I, int H, int precision,
                              float* dx, float* dkernel_fw, float* dbias_fw, float* dkernel_bw, float* dbias_bw,
                              void* workspace, size_t workspace_bytes, void* stream) {
    AMSS_REQUIRE(x && kernel_fw && kernel_bw && y && dy && saved && workspace, "blstm_bwd: null pointer");
    AMSS_REQUIRE(dkernel_fw && dkernel_bw && dbias_fw && dbias_bw, "blstm_bwd: null gradient output");
    if (workspace_bytes < amss_blstm_workspace_bytes(B, T, I, H, precision)) { set_error("blstm_bwd: workspace too small"); return AMSS_ERR_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    const RecGeom geo = rec_geom(B, H);
    AMSS_REQUIRE(geo.smem_bwd <= 220 * 1024, "blstm_bwd: H=%d needs %zu B of shared memory", H, geo.smem_bwd);
    char* ws = (char*)workspace;
    unsigned* bar = (unsigned*)ws;
    float* dZ = (float*)(ws + 256);
    float* dcc = (float*)(ws + 256 + gates_bytes(B, T, H) + cst_bytes(B, T, H));
    void* gws = (char*)dcc + align_up((size_t)2 * B * H * 4, 256);
    const size_t gws_bytes = workspace_bytes - (size_t)((char*)gws - ws);
    const float* gates = (const float*)saved;
    const float* cst = (const float*)((const char*)saved + gates_bytes(B, T, H));
    const bool use_tc = precision == AMSS_PREC_BF16 && blstm_rec_tc_supported(B, T, H);
    // Tensor-core recurrence: its writer warps emit dZ directly as the bf16 GEMM operand plus per-cluster column sums
    // (the bias gradients), so the fp32 dZ, its conversion pass and the column-sum passes disappear.
    int nsub = 0;
    bool fused_dz = false;
    if (use_tc) {
        nsub = blstm_rec_bwd_tc_nsub(B, H);
        fused_dz = 2 * nsub <= CS_CHUNKS;                 // the partials live in the column-sum scratch
        const Bf16Scratch s0 = bf16_scratch(gws, B, T, I, H);
        int rc = blstm_rec_bwd_tc(kernel_fw + (size_t)I * 4 * H, kernel_bw + (size_t)I * 4 * H, 4 * H, gates, cst, dy,
                                  fused_dz ? nullptr : dZ, fused_dz ? s0.dzb : nullptr, s0.H4p, fused_dz ? (float*)gws : nullptr,
                                  B, T, H, st);
        if (rc != AMSS_OK) return rc;
    } else {
    AMSS_CUDA(cudaMemsetAsync(bar, 0, 256, st));
    RecBwdParams p;
    p.Wh[0] = kernel_fw + (size_t)I * 4 * H;
    p.Wh[1] = kernel_bw + (size_t)I * 4 * H;
    p.ldw = 4 * H; p.gates = gates; p.cst = cst; p.dy = dy; p.dZ = dZ; p.dcc = dcc; p.bar = bar;
    p.B = B; p.T = T; p.H = H; p.U = geo.U; p.NC = geo.NC; p.BC = geo.BC;
    AMSS_CUDA(cudaFuncSetAttribute(blstm_rec_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)geo.smem_bwd));
    void* args[] = {&p};
    AMSS_CUDA(cudaLaunchCooperativeKernel((void*)blstm_rec_bwd_kernel, dim3(2 * geo.NC), dim3(RC_THREADS), args,
                                          geo.smem_bwd, st));
    count_launch();
    }
    const float* kern[2] = {kernel_fw, kernel_bw};
    float* dkern[2] = {dkernel_fw, dkernel_bw};
    float* dbias[2] = {dbias_fw, dbias_bw};
    const int H4 = 4 * H;
    const bool bf = precision == AMSS_PREC_BF16;
    Bf16Scratch s{};
    if (bf) {
        // every operand is converted to bf16 ONCE (x, each direction's half of y, dZ of both directions, W_x); the
        // t-1 / t+1 shifted views of dW_h are row offsets (multiples of the padded leading dimension: TMA-aligned)
        s = bf16_scratch(gws, B, T, I, H);
        s.xb = (uint16_t*)((char*)saved + gates_bytes(B, T, H) + cst_bytes(B, T, H));      // written by amss_blstm_fwd (bf16)
        int rc = AMSS_OK;
        for (int d = 0; d < 2 && rc == AMSS_OK && T > 1; ++d) rc = convert_bf16(y + d * H, T * B, H, 2 * H, s.yb[d], s.Hp, st);
        if (rc == AMSS_OK && !fused_dz) rc = convert_bf16(dZ, 2 * T * B, H4, H4, s.dzb, s.H4p, st);
        for (int d = 0; d < 2 && rc == AMSS_OK && dx; ++d) rc = convert_bf16(kern[d], I, H4, H4, s.wb[d], s.H4p, st);
        if (rc != AMSS_OK) return rc;
    }
    for (int d = 0; d < 2; ++d) {
        const float* dZd = dZ + (size_t)d * T * B * H4;
        const uint16_t* dzb = bf ? s.dzb + (size_t)d * T * B * s.H4p : nullptr;
        int rc;
        // dW_x = x^T dZ
        if (bf) rc = gemm_bf16(s.xb, s.Ip, 1, dzb, s.H4p, 1, nullptr, I, H4, T * B, 0, dkern[d], H4, 0, 0, 0, nullptr, st);
        else rc = gemm_dispatch(x, I, dZd, H4, nullptr, I, H4, T * B, 1, 0, 0, precision, dkern[d], H4, 0, 0, gws, gws_bytes, st);
        if (rc != AMSS_OK) return rc;
        // dW_h = h_prev^T dZ : forward dir pairs y[t-1] with dZ[t]; backward dir pairs y[t+1] with dZ[t]
        float* dWh = dkern[d] + (size_t)I * H4;
        if (T > 1) {
            if (bf) {
                rc = gemm_bf16(s.yb[d] + (d == 0 ? 0 : (size_t)B * s.Hp), s.Hp, 1, dzb + (d == 0 ? (size_t)B * s.H4p : 0), s.H4p, 1,
                               nullptr, H, H4, (T - 1) * B, 0, dWh, H4, 0, 0, 0, nullptr, st);
            } else {
                const float* hA = d == 0 ? y : y + (size_t)B * 2 * H + H;
                const float* zB = d == 0 ? dZd + (size_t)B * H4 : dZd;
                rc = gemm_dispatch(hA, 2 * H, zB, H4, nullptr, H, H4, (T - 1) * B, 1, 0, 0, precision, dWh, H4, 0, 0, gws,
                                   gws_bytes, st);
            }
            if (rc != AMSS_OK) return rc;
        } else {
            AMSS_CUDA(cudaMemsetAsync(dWh, 0, (size_t)H * H4 * 4, st));
        }
        if (fused_dz) {
            AMSS_LAUNCH(colsum_sum_kernel, (H4 + 255) / 256, 256, 0, st, (const float*)gws + (size_t)d * nsub * H4, nsub, H4, dbias[d]);
        } else {
            dim3 cg((H4 + 31) / 32, CS_CHUNKS);
            AMSS_LAUNCH(colsum_chunk_kernel, cg, 256, 0, st, dZd, (int64_t)T * B, H4, (float*)gws);
            AMSS_LAUNCH(colsum_sum_kernel, (H4 + 255) / 256, 256, 0, st, (const float*)gws, CS_CHUNKS, H4, dbias[d]);
        }
        if (dx) {
            // dx (+)= dZ W_x^T
            if (bf) rc = gemm_bf16(dzb, s.H4p, 0, s.wb[d], s.H4p, 0, nullptr, T * B, I, H4, d, dx, I, 0, 0, 0, nullptr, st);
            else rc = gemm_dispatch(dZd, H4, kern[d], H4, nullptr, T * B, I, H4, 0, 1, d, precision, dx, I, 0, 0, gws, gws_bytes, st);
            if (rc != AMSS_OK) return rc;
        }
    }
    return AMSS_OK;
}

// Diagnostics: clock64() stamps of the tcgen05 recurrence (CTA 0, steps 100..103, 12 slots per step) are
// written to dev_buf (>= 48 int64) by subsequent amss_blstm_fwd calls; NULL switches it off.
extern "C" int amss_debug_blstm_profile(long long* dev_buf) {
    blstm_tc_set_profile(dev_buf);
    return AMSS_OK;
}
// Diagnostics: schedule trace of the tcgen05 recurrence kernels: per CTA {SM id, globaltimer ns at start, at end, at the end of the prologue},
// forward launches at [0, 4*4096), backward at [4*4096, 8*4096) of dev_buf (>= 8*4096 int64); NULL switches it off.
extern "C" int amss_debug_blstm_sched(long long* dev_buf) {
    blstm_tc_set_sched(dev_buf);
    return AMSS_OK;
}
// Diagnostics: co-resident clusters (cudaOccupancyMaxActiveClusters) of the tensor-core recurrence kernels for H hidden
// units per direction: out[4] = {forward NB=16, forward NB=32, backward NB=16, backward NB=32} (host memory).
extern "C" int amss_debug_blstm_clusters(int H, int* out4) {
    AMSS_REQUIRE(out4 && blstm_rec_tc_supported(1, 1, H), "debug_blstm_clusters: bad arguments");
    blstm_tc_max_clusters(H, out4);
    return AMSS_OK;
}
