/*
 * coclr_b200 -- C ABI of the B200 (sm_100a) kernels behind the CoCLR data-parallel hot path.
 *
 * The reference (TengdaHan/CoCLR) has no FFI layer: its hot path is Python calling nn.Conv3d /
 * nn.BatchNorm3d / nn.MaxPool3d / torch.einsum.  Each entry point below
 * replaces the vendor-library call(s) named in its comment (reference file:line), takes raw device
 * pointers + sizes + a cudaStream_t, never allocates, never synchronises, and returns 0 on success
 * or a negative error code (COCLR_E_*).
 *
 * Data layout: activations are channels-last rows [B, T, H, W, ld].  A convolution OUTPUT is written as
 * fp32 (pre-BatchNorm, needed for the statistics and the backward pass); every convolution INPUT is a
 * pair of 16-bit planes (hi, lo) of the same shape with hi + lo ~= the fp32 value (fp16 pair: ~22 bits,
 * bf16 pair: ~16 bits), produced by coclr_affine_split / coclr_maxpool_fwd / coclr_pack_input /
 * coclr_bn_bwd.  Three tensor-core passes (hi*lo, lo*hi, hi*hi) then give an fp32-grade product.
 *
 * The Python host side (coclr_b200/lib.py) binds these with ctypes; INTEGRATION.md shows the stub.
 */
#ifndef COCLR_B200_H_
#define COCLR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COCLR_OK 0
#define COCLR_E_ARG (-1)    /* bad argument (shape / alignment / null) */
#define COCLR_E_LAUNCH (-2) /* CUDA launch error (cudaGetLastError != success) */
#define COCLR_E_DEVICE (-3) /* not an sm_100 device */

typedef void* coclr_stream_t; /* cudaStream_t */

/* A channels-last split-precision activation as an implicit-GEMM operand source. */
typedef struct {
  const void* hi; /* 16-bit plane [B, T, H, W, ld] (fp16 or bf16), 16-byte aligned */
  const void* lo; /* residual plane, same layout; may be NULL for single-pass kernels */
  int ld;         /* channels per pixel in memory (multiple of 8) */
  int coff;       /* first channel used (multiple of 8) */
  int C;          /* channels used (multiple of 8) */
  int T, H, W;
} coclr_src_t;

typedef struct {
  int kt, kh, kw;
  int st, sh, sw; /* each 1 or 2 */
  int pt, ph, pw;
  int transposed; /* 0: src = dst*s + k - p (forward conv / wgrad gather); 1: src = (dst + p - k)/s (dgrad) */
} coclr_geom_t;

/* ---- implicit-GEMM convolution, forward and data-gradient -----------------------------------
 * replaces nn.Conv3d forward (backbone/s3dg.py:11-13,39-42; model/pretrain.py:52,54) and the cuDNN
 * dgrad reached through loss.backward() (main_nce.py:330).  Y[m, n] (+)= sum_k A[m, k] * Wp[n, k],
 * m over the B*Td*Hd*Wd destination pixels, k = tap * C + channel.  Optionally accumulates the
 * per-channel sum / sum of squares of Y (train-mode BatchNorm3d statistics, s3dg.py:16,46-47). */
typedef struct {
  coclr_src_t src;
  coclr_geom_t g;
  int B, Td, Hd, Wd; /* destination pixel grid */
  int Kreal;         /* taps * src.C */
  const void* wpk;   /* packed weights from coclr_pack_weights */
  const float* wunscale; /* [n_tiles*BN] per-column 1/scale, or NULL */
  int N, BN, n_tiles;    /* real output channels, tile width (multiple of 32, <= 256), tiles */
  float* dst;            /* fp32 [B,Td,Hd,Wd,dst_ld] */
  int dst_ld, dst_coff;
  int accumulate;    /* dst += result */
  double* stats_sum; /* [N] per-output-channel sum of Y, accumulated (zero it first); NULL to skip */
  double* stats_sq;  /* [N] per-output-channel sum of Y*Y */
  int npass;         /* 1 = single 16-bit pass, 3 = hi/lo split (fp32-grade) */
  int a_bf16;        /* format of the src planes: 0 = fp16, 1 = bf16 */
  int b_bf16;        /* format of the packed weights */
  const float* out_scale; /* device scalar multiplied into the result (e.g. 1/s of gradient planes stored scaled by s, see
                           * coclr_bn_bwd_t.dy_scale), or NULL */
} coclr_conv_t;
int coclr_conv_igemm(const coclr_conv_t* p, int num_sms, coclr_stream_t stream);
size_t coclr_conv_packed_bytes(int N, int Kreal, int* BN_out, int* n_tiles_out);
/* coclr_conv_igemm runs stride-1 (1,k,k) / (k,1,1) / 1x1x1 shapes (and the temporally strided (k,1,1) stem conv) whose
 * K chunks do not straddle taps on the TMA-staged kernel (csrc/conv_tma.cu: cp.async.bulk.tensor halo slabs shared by
 * the taps of one dimension, resident or streamed weight tiles, bulk-tensor store / add-reduce epilogue) and everything
 * else on the cp.async gather kernel.  coclr_conv_tma_plan reports (without a GPU) whether a launch takes the TMA
 * kernel: returns 1 and fills info[8] = {A slots, B slots, weights resident, staging buffers, shared-memory bytes,
 * tiles, slab bytes per plane, slab types}, else 0.  coclr_set_conv_tma(0) forces the gather kernel (tests, A/B
 * timing; also COCLR_TMA=0 in the environment). */
int coclr_conv_tma_plan(const coclr_conv_t* p, int* info);
void coclr_set_conv_tma(int enabled);

/* ---- weight-gradient -----------------------------------------------------------------------
 * replaces cuDNN wgrad.  dW[n, c, tap] += sum_m dY[m, n] * A[m, (tap, c)] into the PyTorch weight
 * layout [Cout, Cin_real, kt, kh, kw] (fp32 atomics; zero dW first). */
typedef struct {
  coclr_src_t src;   /* conv input planes (as the forward pass consumed them) */
  coclr_geom_t g;    /* forward geometry (transposed = 0) */
  coclr_src_t dy;    /* output-gradient planes [B,Td,Hd,Wd,ld], C = Cout rounded up to 8 */
  int B, Td, Hd, Wd;
  int Cout, Cin_real; /* real sizes of dW (src.C may be padded, e.g. 8 for the RGB stem) */
  float* dw;
  int npass;
  int dy_bf16, src_bf16;
  int splits;        /* pixel-range splits (>= 1) */
  const float* out_scale; /* device scalar multiplied into the result before it is added to dw, or NULL */
  float* ws;         /* optional workspace (16-byte aligned) of ws_floats floats, see coclr_wgrad_ws_floats; NULL: none */
  long ws_floats;
} coclr_wgrad_t;
int coclr_conv_wgrad(const coclr_wgrad_t* p, coclr_stream_t stream);
/* coclr_conv_wgrad runs stride-1 "same" (1,k,k) / (k,1,1) / 1x1x1 shapes on the TMA-staged kernel (csrc/wgrad_tma.cu:
 * dY tiles and X halo slabs by cp.async.bulk.tensor, the taps of the reuse dimension as the N blocks of one MMA) and
 * everything else (strided convs, the space-to-depth stem) on the cp.async gather kernel.  Tests / tuning can force the
 * gather kernel with coclr_set_wgrad_tma(0) or COCLR_WGRAD_TMA=0; `splits` only steers the gather kernel. */
void coclr_set_wgrad_tma(int enabled);
/* 1 when the shape runs on the TMA-staged kernel; info[8] (may be NULL) = {tile extent dims 0..2, columns per MMA,
 * 64-cout blocks per work item, pipeline stages, pixel splits, work items} */
int coclr_wgrad_tma_plan(const coclr_wgrad_t* p, int* info);
/* floats of workspace with which the TMA-staged kernel writes its per-CTA partial sums with plain stores and adds them
 * to dw in one reduction launch (no fp32 atomics: deterministic, and ~2x faster on the many small layers); 0 for shapes
 * on the gather kernel.  The workspace may be shared by launches on the same stream. */
long coclr_wgrad_ws_floats(const coclr_wgrad_t* p);

/* ---- weight packing -------------------------------------------------------------------------
 * PyTorch conv weight [Cout, Cin, kt, kh, kw] -> swizzled 16-bit hi/lo tile images.
 * mode 0 (forward): rows n = cout, k = tap*Cpad + cin.   mode 1 (dgrad): rows n = cin, k = tap*Cout_pad + cout.
 * Each row is scaled by a power of two (max |w| -> [0.5,1)) before the split; 1/scale goes to unscale. */
typedef struct {
  const float* w;
  int Cout, Cin, taps;
  int Cpad;      /* padded channel count of the K index (>= Cin for mode 0, >= Cout for mode 1) */
  int mode;
  int bf16;
  void* wpk;
  float* unscale; /* [n_tiles*BN] */
} coclr_pack_t;
int coclr_pack_weights(const coclr_pack_t* p, coclr_stream_t stream);
/* every weight of an encoder in one launch: table_dev = device array of n coclr_pack_t, row_start_dev = device int
 * [n + 1], first packed row (of the n_tiles*BN rows coclr_conv_packed_bytes implies) of each entry, ascending */
int coclr_pack_weights_batch(const coclr_pack_t* table_dev, const int* row_start_dev, int n, int total_rows,
                             coclr_stream_t stream);

/* ---- BatchNorm3d statistics -> affine (train mode; backbone/s3dg.py:16,46-47) ------------------
 * per-channel sums written by coclr_conv_igemm -> (scale, shift), saved (mean, rstd) for backward, running-stat
 * momentum update (unbiased variance).  training == 0: scale/shift from the running statistics (eval-mode BN,
 * main_coclr.py:363).  Stand-alone launch (coclr_bn_finalize) or fused into coclr_affine_split. */
typedef struct {
  const double* sum;
  const double* sumsq;
  long count; /* B*T*H*W */
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float momentum, eps;
  int training;
  float* scale;
  float* shift;
  float* save_mean; /* may be NULL */
  float* save_rstd;
  int C;
} coclr_bn_finalize_t;
int coclr_bn_finalize(const coclr_bn_finalize_t* p, coclr_stream_t stream);

/* ---- BatchNorm-apply + ReLU + precision split --------------------------------------------------
 * hi/lo[m, out_coff + c] = split( relu?( scale[c] * x[m, coff + c] + shift[c] ) ); scale == NULL: identity.
 * The elementwise half of nn.BatchNorm3d + nn.ReLU (backbone/s3dg.py:16-17,24-27), one pass, producing the
 * operand planes every consumer (conv / pool / avg-pool) reads. */
typedef struct {
  const float* x;
  int ld, coff, C; /* C multiple of 4 */
  long M;
  const float* scale;
  const float* shift;
  int relu;
  void* hi;
  void* lo; /* may be NULL */
  int out_ld, out_coff;
  int bf16;
  void* hi2; /* optional bf16 twin of the planes (same out_ld / out_coff) for the weight-gradient GEMM, whose */
  void* lo2; /* two operands must share one 16-bit format (tcgen05 kind::f16); NULL to skip */
  /* optional fused BatchNorm finalize (bn.scale != NULL): scale/shift are derived in-kernel from the statistics
   * (bn.training) or the running stats, written to bn.scale/bn.shift/bn.save_* and the running stats are updated
   * by one block -- saves the separate coclr_bn_finalize launch; `scale`/`shift` above are then ignored */
  coclr_bn_finalize_t bn;
  /* optional residual branch of a ResNet bottleneck (`out += residual` then ReLU, backbone/resnet_2d3d.py:75-80,
   * 118-123): operand planes of the same 16-bit format as hi/lo, added after the affine and before the ReLU */
  const void* res_hi; /* NULL: no residual */
  const void* res_lo; /* may be NULL */
  int res_ld, res_coff;
} coclr_split_t;
int coclr_affine_split(const coclr_split_t* p, int num_sms, coclr_stream_t stream);


/* backward of BatchNorm+ReLU: dA (grad w.r.t. relu(bn(y)), fp32) -> dY (grad w.r.t. y) written as bf16
 * hi/lo planes for the dgrad / wgrad kernels; also dgamma / dbeta (accumulated).  Replaces
 * NativeBatchNormBackward0 + ReluBackward0. */
typedef struct {
  const float* y;  /* raw conv output [M, ld] */
  const float* dA; /* same layout */
  int ld, coff, C;
  long M;
  const float* scale; /* [C] gamma*rstd (as written by finalize) */
  const float* shift;
  const float* mean;
  const float* rstd;
  int relu;
  double* sums;   /* workspace [2*C] */
  float* dgamma;  /* [C] or NULL, += */
  float* dbeta;
  void* dy_hi;    /* bf16 planes [M, ld] (same ld / coff as y) */
  void* dy_lo;    /* may be NULL */
  /* residual block output (relu(bn(y) + r), backbone/resnet_2d3d.py:75-80): r's forward planes enter the ReLU mask,
   * and dz = dA*[bn(y)+r > 0] is also the gradient of the residual branch, written (=) or added (+=) to dres */
  const void* res_hi; /* NULL: plain BN(+ReLU) */
  const void* res_lo;
  int res_ld, res_coff, res_bf16;
  float* dres; /* fp32 [M, dres_ld] at dres_coff, or NULL */
  int dres_ld, dres_coff, dres_accumulate;
  /* dy_fp16 != 0: the planes are written as fp16 hi/lo of dY * s, s a power of two chosen per call so that the largest
   * |dY| the data allows (a per-channel bound from max |dz|, max |xhat| and the two means, reduced over the channels)
   * lands at 2^14: ~22 significant bits near the top of the range instead of bf16 hi/lo's 16, and the weight-gradient
   * GEMM can pair them with the forward pass's fp16 activation planes (tcgen05 kind::f16 needs one format for both
   * operands) -- no bf16 twin of every activation.  dy_scale[0] = s, dy_scale[1] = 1/s (device, written by the call;
   * consumers pass &dy_scale[1] as out_scale).  amax: workspace [2*C] floats. */
  int dy_fp16;
  float* amax;
  float* dy_scale;
} coclr_bn_bwd_t;
int coclr_bn_bwd(const coclr_bn_bwd_t* p, int num_sms, coclr_stream_t stream);

/* bias+ReLU backward of the projection head (model/pretrain.py:52-53): dA <- dA*[h+b>0] in place, dbias += */
int coclr_bias_relu_bwd(const float* h, const float* bias, float* dA, float* dbias, int M, int C,
                        coclr_stream_t stream);

/* ---- nn.MaxPool3d (backbone/s3dg.py:105,151,162,173,190) ---------------------------------------- */
typedef struct {
  const void* x_hi; /* input planes [B,Ti,Hi,Wi,ldx] fp16 */
  const void* x_lo;
  int ldx, x_coff;
  void* y_hi;       /* output planes [B,To,Ho,Wo,ldy] fp16 */
  void* y_lo;
  int ldy, y_coff;
  void* y2_hi;      /* optional bf16 twin of the output planes (see coclr_split_t), or NULL */
  void* y2_lo;
  unsigned char* idx; /* [B*To*Ho*Wo, C] arg-max tap, may be NULL in forward-only use */
  int B, C, Ti, Hi, Wi, To, Ho, Wo;
  coclr_geom_t g;
  /* backward only: fp32 gradient buffers */
  const float* dy; /* [B,To,Ho,Wo,C] */
  float* dx;       /* [B,Ti,Hi,Wi,ldx] at x_coff */
  int accumulate;
} coclr_pool_t;
int coclr_maxpool_fwd(const coclr_pool_t* p, coclr_stream_t stream);
int coclr_maxpool_bwd(const coclr_pool_t* p, coclr_stream_t stream);

/* ---- nn.AdaptiveAvgPool3d((1,1,1)) (model/pretrain.py:51) ---------------------------------------- */
int coclr_avgpool_fwd(const void* x_hi, const void* x_lo, int bf16 /* plane format */, int ld, int coff, float* out, int B, int Pn, int C,
                      coclr_stream_t stream);
int coclr_avgpool_bwd(const float* dfeat, float* dA, int ld, int coff, int B, int Pn, int C, coclr_stream_t stream);

/* ---- block[:, i].contiguous() + NCDHW -> channels-last fp16 hi/lo planes with C padded to 8
 * (model/pretrain.py:149-150); batch_index (device int64[B], or NULL) gathers source clips
 * out[b] = x[batch_index[b]] = the shuffle-BN pick x_gather[idx_this] (pretrain.py:124).
 * peer_x != NULL (device array of one clip-buffer base per rank, NVLink peer-mapped, clips_per_peer clips each; x is
 * then ignored): clip g = batch_index[b] is read from peer_x[g / clips_per_peer] + (g % clips_per_peer) * batch_stride,
 * i.e. the all-gather of every rank's clips (pretrain.py:105-106) is replaced by P2P loads of only the B clips this
 * rank encodes.
 * norm_mean / norm_std != NULL: the per-channel (x - mean) / std of the reference's GPU-side input transform `tr`
 * (T.Normalize, main_nce.py:207-209,299-302) is applied on the fly; together with the free batch / channel strides
 * (the loader's [B, C, num_seq*seq_len, H, W] tensor is addressed in place) this folds `tr` -- normalise, view,
 * transpose(1,2), contiguous -- into the packing pass ------------------------------------------------------- */
int coclr_pack_input(const float* x, long batch_stride, long chan_stride, int Cin, void* out_hi, void* out_lo,
                     void* out2_hi, void* out2_lo /* optional bf16 twin */, int B, long thw, const long* batch_index,
                     const void* const* peer_x, int clips_per_peer,
                     const float* norm_mean, const float* norm_std /* device [Cin] or both NULL */,
                     coclr_stream_t stream);

/* space-to-depth variant for the stride-2 7x7 RGB stem (backbone/s3dg.py:145): planes [B, T, H/2, W/2 + 2*pad_x, 16]
 * with channel (dy*2+dx)*Cin + c = x[b, c, t, 2Y+dy, 2X+dx] (4*Cin real channels, rest zero) at pixel X + pad_x of a
 * row; H, W even.  The pad_x pixels on either side of a row are NOT written: the caller zeroes the planes once, which
 * materialises the conv's horizontal zero padding (pad_x = 2 for the 4x4 stem) so that the four taps of a kernel row
 * are one contiguous 128-byte run for every output pixel (the TMA window operand of csrc/conv_tma.cu). */
int coclr_pack_input_s2d(const float* x, long batch_stride, long chan_stride, int Cin, void* out_hi, void* out_lo,
                         void* out2_hi, void* out2_lo, int B, int T, int H, int W, int pad_x, const long* batch_index,
                         const void* const* peer_x, int clips_per_peer, const float* norm_mean, const float* norm_std,
                         coclr_stream_t stream);

/* ---- F.normalize(z + bias, dim=1) (model/pretrain.py:154,167) ------------------------------------- */
int coclr_l2norm_fwd(const float* z, const float* bias, float* q, float* inv_norm, int B, int D, coclr_stream_t stream);
int coclr_l2norm_bwd(const float* q, const float* dq, const float* inv_norm, float* dz, float* dbias, int B, int D,
                     coclr_stream_t stream);

/* ---- _momentum_update_key_encoder (model/pretrain.py:76-80), one launch over the flat parameters -- */
int coclr_ema_update(float* k, const float* q, float m, float one_minus_m, long n, int num_sms, coclr_stream_t stream);

/* ---- _dequeue_and_enqueue column write (model/pretrain.py:93): queue[:, ptr:ptr+n] = keys^T -------- */
int coclr_queue_enqueue(float* queue, const float* keys, int dim, int K, int ptr, int n, coclr_stream_t stream);

/* ---- torch.optim.Adam, coupled L2 (main_nce.py:190-200), flat buffers ----------------------------- */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  long n;
  float grad_scale; /* e.g. 1/world_size after a sum all-reduce */
  float beta1, beta2, eps, weight_decay;
  float step_size; /* lr / (1 - beta1^t) */
  float bc2_sqrt;  /* sqrt(1 - beta2^t) */
} coclr_adam_t;
int coclr_adam_step(const coclr_adam_t* p, int num_sms, coclr_stream_t stream);

/* ---- InfoNCE logits + temperature + cross-entropy (model/pretrain.py:175-182; main_nce.py:201,314) ----
 * logits[b, 0] = q_b.k_b / T, logits[b, 1+j] = q_b.queue[:, j] / T; loss_rows[b] = logsumexp - logits[b,0];
 * dlogits = d(mean_b loss_rows)/d(logits).  loss_rows / dlogits may be NULL.
 * ws: optional scratch of 2 * B * ceil(K / 1024) floats; with it, queues longer than 2048 are cut into 1024-column
 * slices on separate CTAs (config 3: K = 16384 would otherwise run on B of the 148 SMs). */
int coclr_nce_logits_ce(const float* q, const float* k, const float* queue, float T, int B, int D, int K,
                        float* logits, float* loss_rows, float* dlogits, float* ws, coclr_stream_t stream);
/* dq = d(logits)^T contraction with [k | queue] / T (no gradient to k or the queue, pretrain.py:160,176) */
int coclr_nce_logits_bwd(const float* dlogits, const float* k, const float* queue, float T, int B, int D, int K,
                         float* dq, coclr_stream_t stream);

/* ---- CoCLR positive mining (model/pretrain.py:392-413): same-source mask OR top-k of the second view's similarity ----
 * mask[b, 0] = 1; mask[b, 1+j] = (k_vsource[b] == queue_vname[j]) OR j in top-`topk` of kf[b] . queue_second[:, j] taken
 * over the columns that are not same-source (the reference fills those with -inf before torch.topk, :406-407).
 * topk == 0 (or a queue that is not full yet, :404): same-source mask only; kf / queue_second may then be NULL.
 * mask: [B, 1+K] bytes (torch.bool storage). */
int coclr_mask_topk(const float* kf, const float* queue_second, const long* k_vsource, const long* queue_vname, int B, int D,
                    int K, int topk, unsigned char* mask, coclr_stream_t stream);

/* ---- S3D-G feature gating, SelfGating (backbone/s3dg.py:68-78) applied to the branch outputs of a SepInception
 * (:125-129): out[b, c, thw] = sigmoid(fc(mean_thw(a[b])))[c] * a[b, c, thw].  The four branches write channel slices
 * of ONE channels-last concat buffer [B, P = T*H*W, C] (16-bit hi/lo planes, row stride ld); the four nn.Linear layers
 * are the diagonal blocks (coff, n) of the fc step: call coclr_gate_fc / coclr_gate_fc_bwd once per member.
 *   forward : coclr_gate_mean -> coclr_gate_fc (x members) -> coclr_gate_apply (planes scaled in place)
 *   backward: coclr_gate_bwd_reduce (dgate[b, c] = sum_thw dout * a, a = act(y * scale + shift) recomputed from the raw
 *             conv output y and the BatchNorm affine) -> coclr_gate_fc_bwd (x members; dW [n, n] row-major like
 *             nn.Linear.weight, dbias [n], dmean) -> coclr_gate_bwd_apply (dout <- gate * dout + dmean / P, in place:
 *             the gradient w.r.t. the un-gated activation, which coclr_bn_bwd then consumes unchanged).
 * mean / gate / dgate / dmean: fp32 [B, C]. */
int coclr_gate_mean(const void* x_hi, const void* x_lo, int bf16 /* plane format */, int ld, int B, int P, int C, float* mean,
                    coclr_stream_t stream);
int coclr_gate_fc(const float* mean, const float* W, const float* bias, float* gate, int B, int C, int coff, int n,
                  coclr_stream_t stream);
int coclr_gate_apply(void* x_hi, void* x_lo, int bf16, int ld, int B, int P, int C, const float* gate, coclr_stream_t stream);
int coclr_gate_bwd_reduce(const float* y, int ldy, const float* scale, const float* shift, int relu, const float* dout, int ldd,
                          int B, int P, int C, float* dgate, coclr_stream_t stream);
int coclr_gate_fc_bwd(const float* dgate, const float* gate, const float* mean, const float* W, float* dW, float* dbias,
                      float* dmean, int B /* <= 256 */, int C, int coff, int n, coclr_stream_t stream);
int coclr_gate_bwd_apply(float* dout, int ldd, const float* gate, const float* dmean, int B, int P, int C, coclr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* COCLR_B200_H_ */
