/*
 * amss.h -- C ABI of libamss_b200.so: the sm_100a kernels behind the hot path of
 * Totoketchup/Adaptive-MultiSpeaker-Separation (adaptive conv filterbank / STFT twin ->
 * stacked BLSTM embeddings -> k-means masks -> waveform inversion, fwd + bwd + AMSGrad).
 *
 * The reference has no FFI: its operator boundary is Python/TensorFlow graph ops.  Each
 * entry point below replaces the TF op(s) one reference call site lowers to; the call
 * site is cited as file:line relative to the reference repository root.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  Every pointer is a DEVICE pointer
 *     unless the name ends in _host.  The caller owns every buffer (including the
 *     scratch ones, whose sizes come from the *_workspace_bytes queries); the library
 *     never allocates device memory.  Process-wide state is limited to: the lazily
 *     applied cudaFuncSetAttribute settings of its kernels, the cached occupancy /
 *     cluster queries, the launch counter read by amss_launch_count(), and the
 *     diagnostic pointers set by the amss_debug_* calls (NULL unless a tool sets
 *     them).  Calls may be issued from several host threads on different streams; the
 *     amss_debug_* switches are not synchronised with running launches.
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it.
 *   - All tensors are row-major contiguous with the layouts written in the comments.
 *   - Return value: AMSS_OK (0) or a negative AMSS_ERR_* code; amss_last_error()
 *     returns a thread-local message for the last failing call.  Nothing throws.
 *   - `precision`: AMSS_PREC_FP32 = fp32-equivalent arithmetic (SIMT fp32 or 3xbf16
 *     split tensor-core products, used for the 1e-3 parity gate), AMSS_PREC_BF16 =
 *     bf16 operands / fp32 accumulate on tcgen05.
 */
#ifndef AMSS_H_
#define AMSS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define AMSS_OK 0
#define AMSS_ERR_INVALID_ARG (-1)
#define AMSS_ERR_CUDA (-2)
#define AMSS_ERR_UNSUPPORTED (-3)
#define AMSS_ERR_WORKSPACE (-4)

#define AMSS_PREC_FP32 0
#define AMSS_PREC_BF16 1

#define AMSS_POOL_MAX 0     /* --with_max_pool      models/adapt.py:114-117 */
#define AMSS_POOL_AVG 1     /* --with_average_pool  models/adapt.py:118-120 */
#define AMSS_POOL_STRIDE 2  /* strided conv         models/adapt.py:121-122 */

int amss_version(void);
const char* amss_last_error(void);
/* Number of kernels launched by this library in the calling process (all threads). */
uint64_t amss_launch_count(void);

/* ------------------------------------------------------------------------------------ *
 * Adaptive front end  (models/adapt.py:95-134)
 * ------------------------------------------------------------------------------------ */
/* filt[W,N] = |window[k]| * bases[k,n]                      models/adapt.py:106, :234   */
int amss_filterbank_make_filter(const float* window, const float* bases, int W, int N,
                                float* filt, void* stream);
/* d(window), d(bases) from d(filt)  (autograd of the line above)                        */
int amss_filterbank_make_filter_bwd(const float* window, const float* bases,
                                    const float* dfilt, int W, int N, float* dwindow,
                                    float* dbases, void* stream);

/* tf.nn.conv2d(SAME, stride 1) + tf.nn.max_pool_with_argmax(VALID)   adapt.py:115-117
 * (or avg-pool :118-120 / strided conv :121-122), fused: the [Bt,L,N] tensor is never
 * written.  x[Bt,L], filt[W,N] -> y[Bt,Tp,N], argmax_i64[Bt,Tp,N] (per-sample flat
 * index t*N+n, TF convention; may be NULL unless mode==AMSS_POOL_MAX).
 * Tp = (L-pool)/hop+1 (max), L/pool (avg), ceil(L/hop) (stride).                        */
int amss_filterbank_analysis_fwd(const float* x, const float* filt, int Bt, int L, int W,
                                 int N, int pool, int hop, int mode, int precision,
                                 float* y, int64_t* argmax, void* workspace,
                                 size_t workspace_bytes, void* stream);
/* The front end of a training batch, x = [B mixtures ; B*S sources] (adapt.py:41-48: the
 * concat the reference feeds to Adapt.front).  Same outputs as amss_filterbank_analysis_fwd
 * with Bt = B*(S+1).  On the tensor-core path with S = 2 the mixture rows are obtained by
 * linearity of the convolution (response(x_mix) = response(x_0) + response(x_1)) whenever
 * x_mix == x_0 + x_1 holds bit for bit -- how the reference's data pipeline builds its
 * mixtures (data/dataset.py:462-468); the equality is checked on the device for every call
 * and any other batch takes the stock path, so the result never depends on the assumption. */
int amss_filterbank_analysis_mix_fwd(const float* x, const float* filt, int B, int S, int L,
                                     int W, int N, int pool, int hop, int mode, int precision,
                                     float* y, int64_t* argmax, void* workspace,
                                     size_t workspace_bytes, void* stream);
size_t amss_filterbank_analysis_workspace_bytes(int Bt, int L, int W, int N, int pool,
                                                int hop, int mode, int precision);
int amss_filterbank_analysis_out_frames(int L, int W, int pool, int hop, int mode);
/* Backward of the max-pool path w.r.t. the filter (sparse through the argmax):
 * dfilt[k,n] += sum_{b,tp} dy[b,tp,n] * x[b, pos(b,tp,n)+k-pad_left]   (adapt.py:115-117) */
int amss_filterbank_analysis_bwd(const float* x, const float* dy, const int64_t* argmax,
                                 int Bt, int L, int W, int N, int Tp, int accumulate,
                                 float* dfilt, void* workspace, size_t workspace_bytes,
                                 void* stream);
size_t amss_filterbank_grad_workspace_bytes(int W, int N);

/* Sliding box sum: out[s][u] = scale * sum_{j<P} in[s][u + dir*j] (in = 0 outside [0,len_in)), dir = +1 / -1; series s starts
 * at in + s*series_stride (out + s*out_series_stride), consecutive elements are elem_stride apart.  The average-pool front
 * end (tf.layers.average_pooling2d over the stride-1 convolution, models/adapt.py:118-120) is the strided response of the
 * box-filtered signal, and UpSampling2D + conv2d_transpose (:224-243) the sparse overlap-add with box-filtered filters: with
 * this operand transform both run, with all their gradients, on the kernels of the max-pool path.                           */
int amss_box_sum(const float* in, int nser, int64_t len_in, int64_t series_stride, int64_t elem_stride, int P,
                 int dir, float scale, int64_t len_out, int64_t out_series_stride, int64_t out_elem_stride,
                 float* out, void* stream);

/* unpool (utils/ops.py:94-120) + tf.nn.conv2d_transpose(SAME)  (adapt.py:205-252), fused
 * as a sparse overlap-add: out[r,u] = sum_{tp,n} vals[r,tp,n]*filt2[u-pos+pad_left,n],
 * pos = argmax[r / S][tp][n] / N  (the mixture's argmax, tiled S times: adapt.py:212-218).
 * vals[R=B*S,Tp,N], argmax_i64[B,Tp,N], filt2[W,N] -> out[R,L]                           */
int amss_filterbank_synthesis_fwd(const float* vals, const int64_t* argmax,
                                  const float* filt2, int B, int S, int L, int W, int N,
                                  int Tp, int pool, int hop, float* out, void* workspace,
                                  size_t workspace_bytes, void* stream);
size_t amss_filterbank_synthesis_workspace_bytes(int B, int S, int L, int W, int N, int Tp);
/* dvals[R,Tp,N] and dfilt2[W,N] (either may be NULL; overwritten) from dout[R,L].        */
int amss_filterbank_synthesis_bwd(const float* dout, const float* vals,
                                  const int64_t* argmax, const float* filt2, int B, int S,
                                  int L, int W, int N, int Tp, float* dvals, float* dfilt2,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------ *
 * STFT twin  (models/network.py:480-502, 584-607)
 * ------------------------------------------------------------------------------------ */
/* tf.contrib.signal.stft(x, frame, hop, fft_length=frame), periodic hann, no padding.
 * x[R,L] -> spec[R,T,F] interleaved (re,im), mag[R,T,F] (either may be NULL),
 * T = 1+(L-frame)/hop, F = frame/2+1.  frame must be a power of two in [64, 2048].      */
int amss_stft_fwd(const float* x, int R, int L, int frame, int hop, float* spec,
                  float* mag, void* stream);
/* |stft| of the S clean sources + argmax over S (network.py:489-502):
 * non_mix[B,S,L] -> labels_u8[B,T,F] (first index wins ties), mag_non_mix[B,T,F,S]
 * (optional, NULL to skip).                                                            */
int amss_stft_labels(const float* non_mix, int B, int S, int L, int frame, int hop,
                     uint8_t* labels, float* mag_non_mix, void* stream);
/* Separator.postprocessing (network.py:584-607): (mask * |X|) * exp(j*angle(X)) ->
 * tf.contrib.signal.inverse_stft(frame, hop, window_fn=inverse_stft_window_fn(hop)).
 * spec[B,T,F] complex; exactly one of labels_i32[B,T*F] (hard one-hot masks) or
 * masks[B,T*F,S] (soft) is non-NULL -> out[B,S,(T-1)*hop+frame].                        */
int amss_istft_masked_fwd(const float* spec, const int32_t* labels, const float* masks,
                          int B, int S, int T, int frame, int hop, float* out,
                          void* stream);
/* Gradient of the soft-mask variant w.r.t. masks[B,T*F,S] (end-to-end fine-tuning through
 * postprocessing, network.py:697-723): dout[B,S,(T-1)*hop+frame] -> dmasks[B,T*F,S].      */
int amss_istft_masked_bwd(const float* spec, const float* dout, int B, int S, int T,
                          int frame, int hop, float* dmasks, void* stream);

/* ------------------------------------------------------------------------------------ *
 * BLSTM  (utils/ops.py:358-383: BasicLSTMCell gate order i,j,f,o; forget_bias 1.0)
 * ------------------------------------------------------------------------------------ */
/* One bidirectional layer, TIME-MAJOR activations: x[T,B,I]; kernel_{fw,bw}[I+H,4H] (TF
 * layout: rows 0..I-1 multiply x, rows I..I+H-1 multiply h); bias_{fw,bw}[4H] ->
 * y[T,B,2H] (fw | bw).  The reference layout is [B,T,*]: amss_transpose_01 converts.
 * `saved` (optional, NULL for inference) receives what the backward pass needs; it must
 * be handed to amss_blstm_bwd with the SAME precision (the tensor-core path keeps a bf16
 * copy of x in it).                                                                      */
size_t amss_blstm_workspace_bytes(int B, int T, int I, int H, int precision);
size_t amss_blstm_saved_bytes(int B, int T, int I, int H);
int amss_blstm_fwd(const float* x, const float* kernel_fw, const float* bias_fw,
                   const float* kernel_bw, const float* bias_bw, int B, int T, int I,
                   int H, float forget_bias, int precision, float* y, void* saved,
                   void* workspace, size_t workspace_bytes, void* stream);
/* dy[T,B,2H] -> dx[T,B,I] (may be NULL), dkernel_*[I+H,4H], dbias_*[4H] (overwritten). */
int amss_blstm_bwd(const float* x, const float* kernel_fw, const float* kernel_bw,
                   const float* y, const float* dy, const void* saved, int B, int T, int I,
                   int H, int precision, float* dx, float* dkernel_fw, float* dbias_fw,
                   float* dkernel_bw, float* dbias_bw, void* workspace,
                   size_t workspace_bytes, void* stream);

/* Diagnostics only: clock64() stamps of the tensor-core recurrence (CTA 0, steps 100..103, 12
 * slots per step; forward at [0,48), backward at [64,112)) are written to dev_buf (>= 128 int64) by later amss_blstm_fwd calls; NULL = off. */
int amss_debug_blstm_profile(long long* dev_buf);
/* Diagnostics only: clock64() stamps of the tensor-core k-means update pass (CTA 0, tiles 8..11; role 0 loader, 1 MMA
 * issuer, 2 epilogue; dev_buf[(role*4 + tile)*8 + slot], >= 96 int64) written by later amss_kmeans_fit calls; NULL = off. */
int amss_debug_kmeans_profile(long long* dev_buf);
/* Diagnostics only: schedule trace of the tensor-core recurrence kernels: per CTA {SM id, globaltimer ns at start, at end,
 * at the end of the prologue}; forward launches at [0, 4*4096), backward at [4*4096, 8*4096) of dev_buf (>= 8*4096 int64); NULL = off.    */
int amss_debug_blstm_sched(long long* dev_buf);
/* Diagnostics only: co-resident clusters of the tensor-core recurrence kernels for H hidden units
 * per direction, out4 (HOST memory) = {fwd NB=16, fwd NB=32, bwd NB=16, bwd NB=32}.          */
int amss_debug_blstm_clusters(int H, int* out4);

/* ------------------------------------------------------------------------------------ *
 * Dense / embedding head  (utils/ops.py:486-503 Conv1D k=1, :318-324 Normalize)
 * ------------------------------------------------------------------------------------ */
/* C[M,N] (ldc) = op(A) * op(B) (+ bias[N]) (+ C if accumulate); row-major with leading
 * dimensions; transa/transb as in BLAS.  out_swap_b/out_swap_t > 0: row m = t*b_count + b
 * (time-major) is written to row b*t_count + t (batch-major) -- the [T,B,*] -> [B,T,*]
 * hand-over between the BLSTM stack and the embedding reshape (models/dpcl.py:30-36).    */
int amss_gemm(const float* A, int lda, const float* B, int ldb, const float* bias, int M,
              int N, int K, int transa, int transb, int accumulate, int precision, float* C,
              int ldc, int out_swap_b, int out_swap_t, void* workspace,
              size_t workspace_bytes, void* stream);
/* bf16 operands held by the caller (no conversion pass): C[M,N] (ldc) (+)= A B (+ bias), fp32
 * accumulation and output.  Both operands are row-major bf16 matrices with leading dimensions
 * that are multiples of 8 and 16-byte aligned bases (TMA tensor maps are built over them):
 *   a_mn = 0: A is [M,K] (K contiguous);   a_mn = 1: A is stored [K,M] (dW = X^T dZ);
 *   b_mn = 1: B is [K,N] (N contiguous);   b_mn = 0: B is stored [N,K] (dX = dZ W^T).
 * norm_E > 0 (multiple of 8, <= 48, divides N, accumulate = 0): the epilogue applies
 * tf.nn.l2_normalize to every group of norm_E consecutive output columns (the embedding
 * axis of [B,T,F,E], utils/ops.py:318-324) and writes inv_norm[M, N/norm_E] exactly as
 * amss_l2norm_fwd does; the un-normalised product never reaches HBM.  norm_E = 0: plain.
 * Replaces the same reference ops as amss_gemm (utils/ops.py:501-503, :372-380).         */
int amss_gemm_bf16(const uint16_t* A, int lda, int a_mn, const uint16_t* B, int ldb, int b_mn,
                   const float* bias, int M, int N, int K, int accumulate, float* C, int ldc,
                   int out_swap_b, int out_swap_t, int norm_E, float* inv_norm, void* stream);
/* fp32 [rows,cols] (ld) -> bf16 [rows,ldd], ldd % 8 == 0, columns >= cols zero filled.    */
int amss_convert_bf16(const float* src, int rows, int cols, int ld, uint16_t* dst, int ldd,
                      void* stream);
/* in[D0,D1,C] -> out[D1,D0,C]                                                           */
int amss_transpose_01(const float* in, int D0, int D1, int C, float* out, void* stream);
/* Same, writing bf16 rows of ldd (= C padded to 8, zero filled) elements: the time-major ->
 * batch-major hand-over into the embedding head fused with its operand conversion.        */
int amss_transpose_01_bf16(const float* in, int D0, int D1, int C, uint16_t* out, int ldd,
                           void* stream);
size_t amss_gemm_workspace_bytes(int M, int N, int K, int transa, int transb, int precision);
/* dbias[N] = column sums of a bf16 [M,N] matrix (N even); workspace as amss_colsum.       */
int amss_colsum_bf16(const uint16_t* dZ, int64_t M, int N, float* dbias, void* workspace,
                     size_t workspace_bytes, void* stream);
/* v = z * rsqrt(max(sum_E z^2, 1e-12)) over groups of E consecutive values
 * (tf.nn.l2_normalize, axis=3 of [B,T,F,E]).  In place allowed.                         */
int amss_l2norm_fwd(const float* z, int64_t rows, int E, float* v, float* inv_norm,
                    void* stream);
int amss_l2norm_bwd(const float* v, const float* inv_norm, const float* dv, int64_t rows,
                    int E, float* dz, void* stream);
/* dbias[N] = column sums of dZ[M,N]                                                    */
int amss_colsum(const float* dZ, int64_t M, int N, float* dbias, void* workspace,
                size_t workspace_bytes, void* stream);
size_t amss_colsum_workspace_bytes(int64_t M, int N);

/* ------------------------------------------------------------------------------------ *
 * Losses
 * ------------------------------------------------------------------------------------ */
/* DPCL.cost (models/dpcl.py:41-86), Y = one_hot(labels) with on/off = 1/0:
 * mean_b ( ||V^T D V||_F - 2 ||V^T D Y||_F + ||Y^T D Y||_F ),  D = diag(1/sqrt(Y Y^T 1)).
 * V[B,TF,E], labels_u8[B,TF] -> loss[1]; stats (workspace) is reused by the backward.   */
size_t amss_dpcl_workspace_bytes(int B, int64_t TF, int E, int S);
int amss_dpcl_loss_fwd(const float* V, const uint8_t* labels, int B, int64_t TF, int E,
                       int S, float* loss, void* workspace, size_t workspace_bytes,
                       void* stream);
/* Same with a precision selector: AMSS_PREC_BF16 accumulates the Gram statistics on tcgen05.   */
int amss_dpcl_loss_fwd_prec(const float* V, const uint8_t* labels, int B, int64_t TF, int E,
                            int S, int precision, float* loss, void* workspace,
                            size_t workspace_bytes, void* stream);
/* dV[B,TF,E] = dloss * d(loss)/dV, using the workspace filled by the forward call.      */
int amss_dpcl_loss_bwd(const float* V, const uint8_t* labels, const float* dloss, int B,
                       int64_t TF, int E, int S, float* dV, const void* workspace,
                       void* stream);
/* Same, fused with the backward of the tf.nn.l2_normalize that produced V (utils/ops.py:323-324,
 * models/dpcl.py:36-37): V = normalised embeddings, inv_norm[B*TF] as written by amss_l2norm_fwd ->
 * dz[B,TF,E] = gradient w.r.t. the un-normalised embeddings.  dV is never materialised.
 * AMSS_PREC_BF16 runs the [points x E] x [E x E] product on tcgen05 (bf16 operands).           */
int amss_dpcl_loss_bwd_normalized(const float* V, const uint8_t* labels, const float* dloss,
                                  const float* inv_norm, int B, int64_t TF, int E, int S,
                                  int precision, float* dz, const void* workspace, void* stream);
/* Same on tcgen05 with dz written as bf16 [B,TF,E] (E % 8 == 0): the embedding-head backward
 * GEMMs (amss_gemm_bf16) read it directly and the fp32 gradient never reaches HBM.         */
int amss_dpcl_loss_bwd_normalized_bf16(const float* V, const uint8_t* labels, const float* dloss,
                                       const float* inv_norm, int B, int64_t TF, int E, int S,
                                       uint16_t* dz_bf16, const void* workspace, void* stream);
/* DPCL.cost with the weighted label matrix of --function_mask (models/network.py:381-389 feeding
 * models/dpcl.py:41-86): Y[i,:] = weights[i] * one_hot(labels[i]), so D_i = 1/sqrt(w_i * sum_{j in
 * class(i)} w_j).  fp32 kernels; weights[B,TF] are data (no gradient), a zero weight gives the
 * reference's 1/sqrt(0).  The backward writes dV, or -- with inv_norm != NULL -- dz through the
 * l2_normalize Jacobian as amss_dpcl_loss_bwd_normalized does.  Workspace: amss_dpcl_workspace_bytes. */
int amss_dpcl_loss_weighted_fwd(const float* V, const uint8_t* labels, const float* weights, int B,
                                int64_t TF, int E, int S, float* loss, void* workspace,
                                size_t workspace_bytes, void* stream);
int amss_dpcl_loss_weighted_bwd(const float* V, const uint8_t* labels, const float* weights,
                                const float* dloss, const float* inv_norm, int B, int64_t TF, int E,
                                int S, float* dV, const void* workspace, void* stream);
/* L41Model.cost, sampling=None (models/L41.py:47-63, 150-178):
 * mean_{b,tf,s} -log sigmoid(y * <spk[b,s,:], emb[b,tf,:]>), y=+1 if labels==s else -1.
 * spk[B,S,E] = (normalised) gathered speaker vectors.                                   */
/* weights[B,TF] (optional, NULL = 1): y is multiplied by it -- the label weighting of
 * --function_mask / --silence_loss (models/network.py:381-396, amss_label_weights).      */
int amss_l41_loss_fwd(const float* emb, const uint8_t* labels, const float* spk,
                      const float* weights, int B, int64_t TF, int E, int S, float* loss,
                      void* workspace, size_t workspace_bytes, void* stream);
int amss_l41_loss_bwd(const float* emb, const uint8_t* labels, const float* spk,
                      const float* weights, const float* dloss, int B, int64_t TF, int E, int S,
                      float* demb, float* dspk, void* workspace, size_t workspace_bytes,
                      void* stream);
size_t amss_l41_workspace_bytes(int B, int64_t TF, int E, int S);
/* labels_u8[B,T,F] = argmax_s |X_non_mix| from the front output (network.py:369-378):
 * front_y[B(S+1),Tp,N]: rows [0,B) mixtures, rows B+b*S+s the sources.                  */
int amss_plugged_labels(const float* front_y, int B, int S, int64_t TN, uint8_t* labels,
                        void* stream);

/* ------------------------------------------------------------------------------------ *
 * K-means  (models/Kmeans_2.py:14-188) + mask application (models/network.py:554-582)
 * ------------------------------------------------------------------------------------ */
/* X[B,L,E]; init_idx_i32[B*tries,K] = rows of X used as initial centroids (the reference
 * draws them on the host, Kmeans_2.py:61-65; the caller supplies them); notsilent_u8[B,L]
 * or NULL (Kmeans_2.py:76-82); beta: NaN => hard assignments, else softmax(-beta*d^2).
 * -> centroids[B,K,E]; labels_i32[B,L] (hard) or soft[B,L,K] (beta given);
 *    inertia[B,tries] (optional) and best_try_i32[B] (optional).                        */
size_t amss_kmeans_workspace_bytes(int B, int64_t L, int E, int K, int tries);
int amss_kmeans_fit(const float* X, const int32_t* init_idx, const uint8_t* notsilent,
                    int B, int64_t L, int E, int K, int tries, int iters, float beta,
                    int normalize_input, int assign_at_end, float* centroids,
                    int32_t* labels, float* soft, float* inertia, int32_t* best_try,
                    void* workspace, size_t workspace_bytes, void* stream);
/* notsilent = log10(max_L(latent)/latent) < threshold   (Kmeans_2.py:76-79)             */
int amss_kmeans_silence_mask(const float* latent, int B, int64_t L, float threshold,
                             uint8_t* notsilent, void* workspace, void* stream);
/* separated[B*S,TF] = X_input[B,TF] * mask  (network.py:567-580)                        */
int amss_apply_masks(const float* X_input, const int32_t* labels, const float* soft,
                     int B, int S, int64_t TF, float* separated, void* stream);

/* ------------------------------------------------------------------------------------ *
 * Input contract (models/network.py:44-88 placeholders / :65-88 pipeline; data/dataset.py:456-468)
 * ------------------------------------------------------------------------------------ */
/* Builds on the device what the reference's tf.data graph builds on the host: optional
 * per-source normalisation (x-mean)/sqrt(var) of every row of x_non_mix[B,S,L] IN PLACE
 * (--dataset_normalize, data/dataset.py:456-460; stats[B*S,2] = (mean, var), may be NULL) and the
 * mixture x_mix[B,L] = ((x_0 + x_1) + x_2 ...) (data/dataset.py:462-468).  S <= 8.         */
int amss_prepare_inputs(float* x_non_mix, int B, int S, int64_t L, int normalize, float* stats,
                        float* x_mix, void* stream);
/* Separator input options (models/network.py:409-443, 504-521), applied per mixture row of
 * X[B,TF] in this order: abs_input -> pre_func (0 none, 1 sqrt, 2 log10(x+1e-12)) ->
 * normalize (0 none, 1 '01': (x-min)/(max-min), 2 'meanstd': (x-mean)/sqrt(var)) ->
 * silence mask (silence_db > 0: x * [(max - x) < silence_db/20], max of the normalised row).
 * out may alias X.  Forward only (the input of the separator is data in every recipe).  */
int amss_separator_input_prep(const float* X, int B, int64_t TF, int abs_input, int pre_func,
                              int normalize, float silence_db, float* out, void* stream);
/* Label weights of the plugged separator (models/network.py:381-396) from the mixture rows X[B,TF]:
 * function_mask 0 none / 1 linear |X|/max / 2 sqrt / 3 square, times (silence_threshold > 0)
 * [log10(max/|X|) < silence_threshold]  ->  w[B,TF].                                      */
int amss_label_weights(const float* X, int B, int64_t TF, int function_mask,
                       float silence_threshold, float* w, void* stream);

/* ------------------------------------------------------------------------------------ *
 * Optimizer  (utils/ops.py:639-704 AMSGrad; models/network.py:181-192)
 * ------------------------------------------------------------------------------------ */
/* One fused pass over a flat parameter buffer: m,v EMA; vhat=max(vhat,v);
 * p -= lr_t * m / (sqrt(vhat)+eps); lr_t = lr*sqrt(1-beta2^t)/(1-beta1^t) computed by the
 * caller.  grad_scale multiplies g first (1/world_size and/or the global-norm clip).
 * If grad_scale_dev != NULL it is read from the device and multiplied in.               */
int amss_amsgrad_step(float* p, const float* g, float* m, float* v, float* vhat,
                      int64_t n, float lr_t, float beta1, float beta2, float eps,
                      float grad_scale, const float* grad_scale_dev, void* stream);
/* sumsq[0] += sum g^2 (caller zeroes); tf.clip_by_global_norm (models/network.py:191-192):
 * factor = clip / max(|grad_scale| * sqrt(sumsq), clip) -- the norm of grad_scale * g, so that a
 * buffer holding the SUM of G per-rank gradients is clipped like the batch-mean gradient
 * (grad_scale = 1/G).                                                                   */
int amss_sumsq(const float* g, int64_t n, float* sumsq, void* workspace, void* stream);
size_t amss_sumsq_workspace_bytes(void);
int amss_clip_factor(const float* sumsq, float clip, float grad_scale, float* factor, void* stream);
/* --optimizer SGD (models/network.py:183): tf.train.MomentumOptimizer(lr, 0.9):
 * accum = momentum*accum + g ; p -= lr*accum.  lr = the staircase-decayed rate (:175-177),
 * computed by the caller.                                                               */
int amss_momentum_step(float* p, const float* g, float* accum, int64_t n, float lr,
                       float momentum, float grad_scale, const float* grad_scale_dev, void* stream);
/* --optimizer RMSProp (models/network.py:185): tf.train.RMSPropOptimizer(lr) with the TF 1.x
 * defaults decay 0.9, momentum 0.0, epsilon 1e-10: ms = decay*ms + (1-decay)*g^2 ;
 * mom = momentum*mom + lr*g/sqrt(ms+eps) ; p -= mom.  The caller initialises ms to ONE
 * (TF's slot initialiser) and mom to zero.                                              */
int amss_rmsprop_step(float* p, const float* g, float* ms, float* mom, int64_t n, float lr,
                      float decay, float momentum, float eps, float grad_scale,
                      const float* grad_scale_dev, void* stream);

/* ------------------------------------------------------------------------------------ *
 * Adapt pre-training losses (models/adapt.py:307-402, models/network.py:196-221)
 * ------------------------------------------------------------------------------------ */
/* Per (b,s) sums over L: tt=<s,s>, aa=<a,a>, ta=<s,a>, ee=<s-a,s-a> -> stats[B*S,4]     */
int amss_wave_stats(const float* target, const float* approx, int R, int64_t L,
                    float* stats, void* stream);

/* The remaining terms of the pre-training graph over the front output y[B(S+1),TN] (TN = Tp*N; mixture rows
 * first), one fused pass (models/adapt.py:127-132 p_hat + KL sparsity, :141-160 overlap measure, :162-196 the
 * pre-training separator -- separation 0 'mask', 1 'perfect' --, :315-316 non-negativity):
 *   sep[B*S,TN] (may be NULL), p_hat[TN], terms[3] = {sparse_constraint, overlapping, mean_rows sum neg^2}.
 * bwd: dy[B(S+1),TN] from dsep (may be NULL) and dterms[3].                                        */
/* The scalar tail of Adapt.cost, pre-training branch (models/adapt.py:323-337, 374-385; models/network.py:196-221):
 * stats[B*S,4] as written by amss_wave_stats(target, synthesis), mix_stats[B*S,4] = amss_wave_stats_rows(target, mixture)
 * (may be NULL: no SDR-improvement metric), terms[3] of amss_adapt_terms_fwd, regsq[1] = |filt|^2 + |filt2|^2 ->
 * out4 = (cost, l2, sdr, sdr_improvement) and the derivatives of cost: dstats[B*S,4], dterms[3], dreg[1] (d cost / d regsq * 2:
 * the factor the filters are multiplied by).  loss_kind 0 l2, 1 sdr, 2 l2 + sdr; zero coefficients drop their term.      */
int amss_adapt_cost_fwd(const float* stats, const float* mix_stats, const float* terms, const float* regsq,
                        int B, int S, int loss_kind, float beta, float lambda, float overlap_coef,
                        float nonneg_coef, float* out4, float* dstats, float* dterms, float* dreg,
                        void* stream);
/* Gradient of amss_wave_stats w.r.t. the approximation: dapprox[R,L] from dstats[R,4] (the targets are data).     */
int amss_wave_stats_bwd(const float* target, const float* approx, const float* dstats, int R,
                        int64_t L, float* dapprox, void* stream);
/* amss_wave_stats with approx row r / approx_div (the mixture of target row r when approx_div = S).  */
int amss_wave_stats_rows(const float* target, const float* approx, int R, int64_t L, int approx_div,
                         float* stats, void* stream);
size_t amss_adapt_terms_workspace_bytes(int64_t TN);
int amss_adapt_terms_fwd(const float* y, int B, int S, int64_t TN, float rho, int separation,
                         float* sep, float* p_hat, float* terms, void* workspace,
                         size_t workspace_bytes, void* stream);
int amss_adapt_terms_bwd(const float* y, const float* p_hat, const float* dsep, const float* dterms,
                         int B, int S, int64_t TN, float rho, int separation, float* dy,
                         void* stream);

/* Tail of the enhance layer (models/network.py:640-693), fused: logits[B,S,TF] (the enhance Conv1D output) ->
 * nonlinearity over the sources (0 softmax, 1 tanh, 2 none) -> masks[B,TF,S] (optional output) * X_input[B,TF] ->
 * table[B,S,S], table[b][s][k] = sum_j (X_non_mix[b,j,s] - mask[b,j,k] X_input[b,j])^2: the pairwise distances the PIT
 * enhance cost minimises over permutations (the caller picks the permutation from S*S scalars per mixture).
 * bwd: perm[B,S] = estimate paired with source s, dcost_b[B] -> dlogits[B,S,TF].               */
size_t amss_enhance_cost_workspace_bytes(int B, int64_t TF, int S);
int amss_enhance_cost_table(const float* logits, const float* X_input, const float* X_non_mix, int B,
                            int S, int64_t TF, int nonlinearity, float* masks, float* table,
                            void* workspace, size_t workspace_bytes, void* stream);
int amss_enhance_cost_bwd(const float* logits, const float* X_input, const float* X_non_mix,
                          const int32_t* perm, const float* dcost_b, int B, int S, int64_t TF,
                          int nonlinearity, float* dlogits, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* AMSS_H_ */
