/*
 * erfnet_b200.h -- C ABI of liberfnet_b200.so: hand-written sm_100a kernels for the ERFNet hot path.
 *
 * The reference (lorenwel/erfnet_pytorch) is pure Python on torch and has NO native interface of its own
 * (SURVEY.md 2b); every entry point below replaces a run-time torch/cuDNN library call made by the reference
 * at the cited file:line.  The Python host side (erfnet_pytorch_b200/) mirrors the reference's nn.Module API
 * and reaches these symbols through ctypes (INTEGRATION.md shows the binding).
 *
 * Contract:
 *  - plain pointers and sizes only; all buffers are caller-owned DEVICE memory (torch tensors); the library
 *    allocates nothing persistent and never frees;
 *  - every call only ENQUEUES work on `stream` (a cudaStream_t) and never synchronises;
 *  - re-entrant, no global mutable state except a thread-local last-error string;
 *  - returns ERF_OK or a negative code (never throws/aborts); erf_last_error_string() describes the failure.
 *
 * Activations are NHWC ("pixel-major"): element (n,y,x,c) at ((n*H + y)*W + x)*ld + c with ld >= C the pixel
 * pitch in elements.  dtype selects fp32 (check mode) or bf16 (fast path) storage; accumulation is fp32.
 */
#ifndef ERFNET_B200_H_
#define ERFNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ERF_OK 0
#define ERF_ERR_INVALID_ARGUMENT (-1)
#define ERF_ERR_CUDA (-2)
#define ERF_ERR_UNSUPPORTED (-3)

#define ERF_F32 0
#define ERF_BF16 1

typedef void* erf_stream_t; /* cudaStream_t */

const char* erf_last_error_string(void);
int erf_abi_version(void);

/* ---- generic gather convolution (check path for every conv of the net, fwd and dgrad) --------------------
 * transposed == 0 (nn.Conv2d, eval/erfnet.py:16,30-38,85):
 *   out[n,oy,ox,co] = sum_{ky,kx,ci} in[n, oy*stride - pad_h + ky*dil_h, ox*stride - pad_w + kx*dil_w, ci] * w[ky,kx,ci,co]
 * transposed == 1 (nn.ConvTranspose2d, eval/erfnet.py:102,124):
 *   out[n,oy,ox,co] = sum over taps with (oy + pad_h - ky*dil_h) % stride == 0 (same in x) of
 *                     in[n,(oy + pad_h - ky*dil_h)/stride,(ox + pad_w - kx*dil_w)/stride,ci] * w[ky,kx,ci,co]
 * Weights are pre-packed by the host as w[(ky*KW + kx)*Cin + ci][w_ld], w_ld = Cout rounded up to 8, zero padded,
 * in the activation dtype.  A dgrad is the same call with a differently packed weight tensor. */
typedef struct {
  int N, H, W, Cin, in_ld;
  int Ho, Wo, Cout, out_ld, out_coff;
  int KH, KW, stride, pad_h, pad_w, dil_h, dil_w, transposed;
  int w_ld;
} erf_conv_geom;

/* epilogue, applied in this order: + bias -> statistics -> + add -> relu -> * (mask > 0) */
typedef struct {
  const float* bias;   /* [Cout] or NULL */
  int relu;            /* out = max(out, 0) */
  const void* mask;    /* NHWC, same dtype: out *= (mask > 0)   (ReLU backward), or NULL */
  int mask_ld;
  const void* add;     /* NHWC, same dtype: out += add (residual / gradient accumulation), or NULL */
  int add_ld;
  double* stats;       /* [2*stats_c]: stats[out_coff+co] += sum, stats[stats_c+out_coff+co] += sum of squares of
                          (conv + bias), atomically accumulated (BN batch statistics, eval/erfnet.py:22,50,56), or NULL */
  int stats_c;         /* channel count of the tensor the statistics are taken over (>= out_coff + Cout) */
} erf_conv_epilogue;

int erf_conv_fwd(const void* in, const void* w, void* out, const erf_conv_geom* g, const erf_conv_epilogue* e,
                 int dtype, erf_stream_t stream);

/* Weight / bias gradient of the convolution described by g (autograd of the above).
 * dw: fp32 [(KH*KW)*Cin][w_ld], db: fp32 [Cout] or NULL; both are ACCUMULATED into (caller zeroes). */
int erf_conv_wgrad(const void* in, const void* dout, float* dw, float* db, const erf_conv_geom* g, int dtype,
                   erf_stream_t stream);
/* Same, but dw is ACCUMULATED in the torch parameter layout: nn.Conv2d [Cout][cin_real][KH][KW] (g->transposed == 0) or
 * nn.ConvTranspose2d [cin_real][Cout][KH][KW] (g->transposed == 1); input channels >= cin_real (padding) are dropped. */
int erf_conv_wgrad_torch(const void* in, const void* dout, float* dw, float* db, const erf_conv_geom* g, int cin_real,
                         int dtype, erf_stream_t stream);

/* ---- the 3-channel stem: DownsamplerBlock(3, 16) = cat([Conv2d(3, 13, 3, stride 2, padding 1), MaxPool2d(2, 2)]) (eval/erfnet.py:67, 12-23)
 * x NHWC [N][H][W][4] (channel 3 = padding), y NHWC [N][H/2][W/2][16]: channels 0..12 = conv + bias, 13..15 = pooled input.
 * y = relu?(raw * scale[c] + shift[c]) (scale/shift [16] or NULL = identity; eval mode passes the BN affine of the pooled channels and
 * folds the conv channels' BN into w / bias); stats (fp64 [32], optional) += per-channel sum and sum of squares of the stored y.
 * w: torch layout [13][3][3][3] fp32.  CUDA-core kernels: K = 27 fits no tensor-core shape and the block is bandwidth-bound. */
int erf_stem_fwd(const void* x, int N, int H, int W, const float* w, const float* bias, const float* scale, const float* shift,
                 int relu, void* y, double* stats, int dtype, erf_stream_t stream);
/* The same block reading the caller's NCHW fp32 image [N][3][H][W] directly (inference: no separate layout pass); y has element type dtype. */
int erf_stem_fwd_nchw(const float* x, int N, int H, int W, const float* w, const float* bias, const float* scale, const float* shift,
                      int relu, void* y, double* stats, int dtype, erf_stream_t stream);
/* dw[13][3][3][3] += conv weight gradient from x and dy (the 16-channel gradient of y; channels 13..15 are ignored);
 * db[13] += per-channel sum of dy (NULL to skip). */
int erf_stem_wgrad(const void* x, const void* dy, int N, int H, int W, float* dw, float* db, int dtype, erf_stream_t stream);

/* ---- MaxPool2d(2, stride 2) half of DownsamplerBlock (eval/erfnet.py:17,21) ------------------------------ */
/* out = relu?(max2x2(in)*scale[c] + shift[c]); scale/shift NULL = identity; stats (of the raw max) as in erf_conv_epilogue */
int erf_maxpool2x2_fwd(const void* in, int N, int H, int W, int C, int in_ld, void* out, int out_ld, int out_coff,
                       const float* scale, const float* shift, int relu, double* stats, int stats_c, int dtype,
                       erf_stream_t stream);
/* dx[n,y,x,c] += dout[n,y/2,x/2,out_coff+c] at the first maximum of each window (torch tie rule). */
int erf_maxpool2x2_bwd(const void* in, int N, int H, int W, int C, int in_ld, const void* dout, int out_ld,
                       int out_coff, void* dx, int dx_ld, int dtype, erf_stream_t stream);

/* ---- layout conversion at the API boundary (reference tensors are NCHW fp32) ---------------------------- */
int erf_nchw_f32_to_nhwc(const float* in, int N, int C, int H, int W, void* out, int out_ld, int dtype,
                         erf_stream_t stream); /* channels C..out_ld-1 are zero filled */
int erf_nhwc_to_nchw_f32(const void* in, int in_ld, int N, int C, int H, int W, float* out, int dtype,
                         erf_stream_t stream);

/* ---- elementwise: out = relu?( x*scale[i] + shift[i] + res ), i = c (sc_per_sample==0) or n*C+c ---------
 * BatchNorm apply (+Dropout2d scale, + residual) + ReLU of eval/erfnet.py:22-23,50-51,56-61,107-108.
 * scale/shift may be NULL (identity), res may be NULL. */
int erf_affine_act(const void* x, int x_ld, const float* scale, const float* shift, int sc_per_sample,
                   const void* res, int res_ld, void* out, int out_ld, int N, int HW, int C, int relu, int dtype,
                   erf_stream_t stream);

/* out = g * (mask > 0): backward of the F.relu calls of eval/erfnet.py:23,48,51,54,61,108 when not fused */
int erf_relu_bwd(const void* g, int g_ld, const void* mask, int mask_ld, void* out, int out_ld, long long n_px, int C,
                 int dtype, erf_stream_t stream);

/* stats[c] += sum_p x[p,c]; stats[C+c] += sum_p x[p,c]^2 over a stored NHWC activation (BN batch statistics when the
 * producing kernel does not accumulate them itself) */
int erf_channel_stats(const void* x, int ld, int C, long long n_px, double* stats, int dtype, erf_stream_t stream);

/* ---- BatchNorm2d(eps=1e-3) training statistics --------------------------------------- */
/* stats[2C] (sum, sumsq over `count` values/channel) -> mean, invstd (biased var), running stats update
 * (momentum, unbiased var), and the fused apply vectors scale/shift: [C], or [N,C] when drop ([N,C] Dropout2d
 * multipliers) is given.  stats is reset to zero. */
int erf_bn_finalize(double* stats, double count, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float momentum, float eps, const float* drop, int N, int C, float* mean,
                    float* invstd, float* scale, float* shift, erf_stream_t stream);
/* eval mode: scale = gamma/sqrt(rv+eps), shift = beta - rm*scale */
int erf_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                       float eps, int C, float* scale, float* shift, erf_stream_t stream);
/* backward: gb = g * (mask>0) * drop[n,c];  sums[c] += gb, sums[C+c] += gb*xhat,  xhat = (y-mean)*invstd */
int erf_bn_bwd_reduce(const void* g, int g_ld, const void* y, int y_ld, const void* mask, int mask_ld,
                      const float* drop, const float* mean, const float* invstd, int N, int HW, int C,
                      double* sums, int dtype, erf_stream_t stream);
/* dy = gamma*invstd*(gb - sums[c]/M - xhat*sums[C+c]/M); also writes dgamma = sums[C+c], dbeta = sums[c] and, when gmasked is
 * given, gmasked = g * (mask>0) (the ReLU backward of the residual path, eval/erfnet.py:61). */
int erf_bn_bwd_apply(const void* g, int g_ld, const void* y, int y_ld, const void* mask, int mask_ld,
                     const float* drop, const float* gamma, const float* mean, const float* invstd,
                     const double* sums, int N, int HW, int C, void* dy, int dy_ld, float* dgamma, float* dbeta,
                     void* gmasked, int gmasked_ld, int dtype, erf_stream_t stream);

/* ---- train-mode nn.BatchNorm2d in ONE launch per direction (csrc/bn_fused.cu; eval/erfnet.py:19,37,55,105 in training mode) ----------
 * erf_bn_fused_supported: 1 when C takes the single-launch kernels (C % 8 == 0, C/8 divides 32, C <= 256). */
int erf_bn_fused_supported(int C);
/* Forward: stats = fp64 [2C] batch sums (sum x, sum x^2) of x accumulated by its producer (read only; the caller re-zeroes them for the next
 * step).  mean = stats[c]/count, var (biased) = stats[C+c]/count - mean^2, invstd = 1/sqrt(var + eps);
 * out = relu?(((x - mean) * invstd * gamma + beta) * drop[n,c] (+ res)); mean / invstd (fp32 [C]) are written for the backward pass and the
 * running statistics updated with `momentum` and the UNBIASED variance (torch semantics) -- replaces erf_bn_finalize + erf_affine_act. */
int erf_bn_fwd_apply(const void* x, int x_ld, const double* stats, double count, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, float momentum, float eps, const float* drop, const void* res,
                     int res_ld, void* out, int out_ld, int N, int HW, int C, int relu, float* mean, float* invstd, int dtype,
                     erf_stream_t stream);
/* Backward: erf_bn_bwd_reduce + erf_bn_bwd_apply as one kernel with a grid barrier between the two phases; what a thread read in phase 1
 * stays in registers / shared memory for phase 2.  sums (fp64 [2C]) and counter (one 32-bit word) must be ZERO on entry.  Falls back to the
 * two launches when the grid would not be co-resident or the layout is unsupported (same arithmetic).
 * beta_for_mask (optional, bf16 path): the BN's beta, given when `mask` is exactly relu(BN(y)) of THIS BatchNorm with no dropout and no
 * residual (BN1 of a block, the samplers' BN): the kernel then recomputes the mask as fma(y, gamma*invstd, beta - mean*gamma*invstd) > 0, the
 * forward pass's own expression, instead of reading the mask tensor (one tensor read less). */
int erf_bn_bwd_fused(const void* g, int g_ld, const void* y, int y_ld, const void* mask, int mask_ld, const float* beta_for_mask,
                     const float* drop, const float* gamma, const float* mean, const float* invstd, double* sums, unsigned* counter,
                     int N, int HW, int C, void* dy, int dy_ld, float* dgamma, float* dbeta, void* gmasked, int gmasked_ld,
                     int dtype, erf_stream_t stream);
/* The same with `sum_copies` >= 1 zeroed copies of the sums (fp64 [sum_copies][2C]): CTA b accumulates into copy b % sum_copies and phase 2 adds the
 * copies up -- 148 fp64 atomics on one address serialise in the L2 (~3 us at the end of phase 1, tools/timeline_bn.py); 8 copies hide that. */
int erf_bn_bwd_fused_spread(const void* g, int g_ld, const void* y, int y_ld, const void* mask, int mask_ld, const float* beta_for_mask,
                            const float* drop, const float* gamma, const float* mean, const float* invstd, double* sums, int sum_copies,
                            unsigned* counter, int N, int HW, int C, void* dy, int dy_ld, float* dgamma, float* dbeta, void* gmasked,
                            int gmasked_ld, int dtype, erf_stream_t stream);

/* SMs a 3-tap conv / data-gradient launch (main_sms) and a 3-tap weight-gradient launch (side_sms) may take from now on; 0 = all.  Host-side
 * state read at launch / graph-capture time: parallel.FlatAdamTrainer sets it around a backward pass whose weight gradients run on a second
 * stream (ERF_WGRAD_SIDE=1), so that the two kinds of persistent kernels really run side by side. */
int erf_set_launch_cap(int main_sms, int side_sms);

/* ---- CrossEntropyLoss2d (train/evaluation_functions.py:6-14) ------------------------------------------- */
/* logits element (p, c): NHWC -> logits[p*ld + c]; NCHW (nchw_hw > 0) -> logits[(p/hw)*C*hw + c*hw + p%hw].
 * pass 1: wsum += sum_valid w[y];  pass 2: loss_sum += sum_valid w[y]*nll, dlogits = w[y]*(softmax-onehot)/wsum. */
int erf_ce_weight_sum(const int64_t* target, long long n_px, const float* weight, int C, int ignore_index,
                      float* wsum, erf_stream_t stream);
int erf_ce_fwd_bwd(const void* logits, int ld, int nchw_hw, const int64_t* target, long long n_px,
                   const float* weight, int C, int ignore_index, const float* wsum, float* loss_sum,
                   void* dlogits, int dtype, erf_stream_t stream);

/* Same for NHWC bf16 logits of pitch 24 or 32, which also accumulates colsum[c] += sum over pixels of the STORED (bf16) dlogits[p,c]: the bias
 * gradient of the convolution that produced the logits (decoder.output_conv, eval/erfnet.py:124), without a separate pass over the 151 MB of
 * dlogits. */
int erf_ce_fwd_bwd_colsum_supported(int ld, int nchw_hw, int dtype);
int erf_ce_fwd_bwd_colsum(const void* logits, int ld, const int64_t* target, long long n_px, const float* weight, int C,
                          int ignore_index, const float* wsum, float* loss_sum, void* dlogits, float* colsum, erf_stream_t stream);

/* ---- iouEval (eval/iouEval.py:21-69) and argmax (eval/eval_iou.py:92) ----------------------------------- */
/* cm[g*nc + p] += #pixels with label g predicted p (u64, bit exact); labels outside [0,nc) are skipped. */
int erf_confusion_hist(const void* pred, const void* gt, int elem_bytes, long long n, int nc,
                       unsigned long long* cm, erf_stream_t stream);
/* tp/fp/fn[c] for c < kept, float64 like the reference accumulators; ignore < 0 means none. */
int erf_iou_counts(const unsigned long long* cm, int nc, int ignore, int kept, double* tp, double* fp, double* fn,
                   erf_stream_t stream);
int erf_argmax(const void* logits, int ld, int nchw_hw, long long n_px, int C, void* out, int out_elem_bytes,
               int dtype, erf_stream_t stream);

/* ---- masked regression losses of the fork (train/evaluation_functions.py:32-131) -------------------------- */
/* kind                      | a             | b            | prob        | per-pixel loss over labelled pixels (target != -1)
 * ERF_LOSS_L1               | out [n_px]    | -            | -           | |a - t|                                  (L1LossMasked l.89-97)
 * ERF_LOSS_LOGLIK           | mean [n_px]   | var [n_px]   | -           | log(2 pi (b+eps)) + (a-t)^2/(b+eps)      (LogLikelihoodLossMasked l.71-87)
 * ERF_LOSS_L1_CLASSPROB     | cost [C]      | -            | [B,C,H,W]   | sum_c |a_c - t| prob_c                   (L1LossClassProbMasked l.32-47)
 * ERF_LOSS_LOGLIK_CLASSPROB | cost [C]      | var [C]      | [B,C,H,W]   | sum_c (log(2 pi v_c) + (a_c-t)^2/v_c) prob_c, v_c = b_c+eps   (l.49-69)
 * ERF_LOSS_MSE_WEIGHTED     | out [n_px]    | -            | -           | (a - t)^2 over ALL elements              (MSELossWeighted l.122-131)
 * All operands fp32; target [n_px] float; hw = H*W (class kinds).  fwd: acc[0] += sum, acc[1] += count (fp64, caller-zeroed);
 * loss = scale * acc[0] / acc[1].  bwd: gradients of that loss times go[0] (go may be NULL = 1): d_a/d_b [n_px] for the per-pixel
 * kinds, d_prob [B,C,H,W] for the class kinds, whose per-class sums are ACCUMULATED UNSCALED into dclass[0..C) (cost) and
 * dclass[C..2C) (variance) in fp64 (caller-zeroed; multiply by go*scale/acc[1]). */
#define ERF_LOSS_L1 0
#define ERF_LOSS_LOGLIK 1
#define ERF_LOSS_L1_CLASSPROB 2
#define ERF_LOSS_LOGLIK_CLASSPROB 3
#define ERF_LOSS_MSE_WEIGHTED 4
int erf_masked_loss_fwd(int kind, const float* a, const float* b, const float* prob, const float* target,
                        long long n_px, long long hw, int C, float eps, double* acc, erf_stream_t stream);
int erf_masked_loss_bwd(int kind, const float* a, const float* b, const float* prob, const float* target,
                        long long n_px, long long hw, int C, float eps, float scale, const double* acc,
                        const float* go, float* d_a, float* d_b, float* d_prob, double* dclass, erf_stream_t stream);

/* ---- SoftMaxConv tail (train/erfnet_blocks.py:104-110): Dropout2d multiplier + softmax over channels -------------------------- */
/* prob (NCHW fp32) = softmax_c(logits[n,p,c] * drop[n,c]); logits NHWC with row pitch ld >= C (C <= 32); drop [N,C] or NULL. */
int erf_softmax_nhwc_fwd(const void* logits, int ld, const float* drop, int N, long long hw, int C, float* prob, int dtype,
                         erf_stream_t stream);
/* dlogits[n,p,c] = drop[n,c] * prob_c * (dprob_c - sum_k dprob_k prob_k); channels [C, ld) are written as zeros. */
int erf_softmax_nhwc_bwd(const float* prob, const float* dprob, const float* drop, int N, long long hw, int C, void* dlogits,
                         int ld, int dtype, erf_stream_t stream);

/* ---- Adam as configured by train/main.py:43-46,213 (L2 in the gradient), over one flat fp32 buffer ------- */
int erf_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, float grad_scale, erf_stream_t stream);

/* Same, with the step counter on the device (*step_dev is incremented, then used for the bias corrections) so that the
 * call can be captured in a CUDA graph and replayed. */
int erf_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                      float eps, float weight_decay, int* step_dev, float grad_scale, erf_stream_t stream);

/* Same, with the learning rate and the weight decay ALSO in device memory (hyper_dev = fp32 [2] = {lr, weight_decay}): a launch captured
 * in a CUDA graph bakes by-value arguments in, whereas the reference steps a LambdaLR every epoch (train/main.py:238-239) -- the host
 * updates hyper_dev between replays (FlatAdamTrainer.set_lr) and the replayed kernel sees the new values. */
int erf_adam_step_graph(float* p, const float* g, float* m, float* v, long long n, const float* hyper_dev, float beta1,
                        float beta2, float eps, int* step_dev, float grad_scale, erf_stream_t stream);

/* ---- tcgen05 implicit-GEMM 3x1 / 1x3 (dilated) convolution, C = 64 or 128, NHWC bf16 ----------
 * out[p,co] = epilogue(sum_t sum_ci in[p + (t-1)*dil along H (horizontal=0) or W (horizontal=1), ci] * Wt[t][co][ci]),
 * epilogue as erf_conv_fwd (bias -> + res -> relu -> * (mask>0)).  TMA-staged 128B-swizzled activation tiles (zero-filled
 * borders = padding), tcgen05.mma with TMEM accumulators.  w_img = host-packed shared-memory image [3][C/64][C][64] bf16
 * with the 16-byte chunks of row n XOR-swizzled by (n & 7) (erfnet_pytorch_b200.ops.pack_conv1d_tc).
 * stats  (fp64 [2C]): += per-channel sum and sum of squares of the stored (bf16) output -- the BN batch statistics;
 * colsum (fp32 [C]): += per-channel sum of the stored output -- the bias gradient when this call is a data gradient.
 * relu: bit 0 = ReLU in the epilogue; bit 1 = "static weights": no kernel in flight on the stream writes w_img (cached inference images), so
 * the kernel may copy it to shared memory before the programmatic-dependent-launch wait, under the previous kernel's tail (same for the pair kernel).
 * Returns ERF_ERR_UNSUPPORTED unless erf_conv1d_tc_supported(C, W, dil). */
int erf_conv1d_tc_supported(int C, int W, int dil);
int erf_conv1d_tc(const void* in, const void* w_img, const float* bias, void* out, const void* res, const void* mask,
                  int N, int H, int W, int C, int horizontal, int dil, int relu, double* stats, float* colsum,
                  erf_stream_t stream);

/* Fused pair out = act(conv1x3_d(relu(conv3x1_d(in) + bias1)) + bias2 [+ res]) -- one half of a non_bottleneck_1d block in eval mode
 * (eval/erfnet.py:47-51 / 53-61 with BN folded into w2/bias2) -- in one kernel: the intermediate row stays in shared memory.
 * C = 16 or 64, W % 128 == 0, (W/128)*C <= 128; images as for erf_conv1d_tc; results are bit-identical to two erf_conv1d_tc calls. */
int erf_conv1d_pair_tc_supported(int C, int W, int dil);
int erf_conv1d_pair_tc(const void* in, const void* w1_img, const float* bias1, const void* w2_img, const float* bias2, void* out,
                       const void* res, int N, int H, int W, int C, int dil, int relu, erf_stream_t stream);

/* Build both operand images (forward and data-gradient) of a 3-tap conv weight w[Co][Ci][3] fp32 (= nn.Conv2d weight
 * [Co,Ci,3,1] or [Co,Ci,1,3], contiguous) for erf_conv1d_tc; each image is 3*C*C bf16. */
int erf_pack_conv1d_tc(const float* w, int C, void* img_fwd, void* img_dgrad, erf_stream_t stream);
/* All 3-tap convs of a net in one launch: table (DEVICE memory) of n_entries {fp32 weight, forward image, data-gradient
 * image, C}; max_C = largest C in the table.  Same images as erf_pack_conv1d_tc. */
typedef struct erf_pack1d_entry {
  const float* w;
  void* img_fwd;
  void* img_dgrad;
  int C;
  int reserved;
} erf_pack1d_entry;
int erf_pack_conv1d_tc_batched(const erf_pack1d_entry* table_dev, int n_entries, int max_C, erf_stream_t stream);

/* tcgen05 weight gradient of the same convolutions: dw[co][ci][t] += sum_p x[p + (t-1)*dil][ci] * dy[p][co]  (fp32, in
 * the nn.Conv2d weight layout [Co,Ci,3,1] == [Co,Ci,1,3]; ACCUMULATED, e.g. straight into the parameter's .grad).  Both NHWC operands are consumed MN-major; per-CTA TMEM accumulators are dumped to
 * `workspace` (>= erf_conv1d_wgrad_tc_workspace(C) bytes, caller-owned) and reduced by a second kernel.  W % 64 == 0. */
size_t erf_conv1d_wgrad_tc_workspace(int C);
int erf_conv1d_wgrad_tc(const void* x, const void* dy, float* dw, void* workspace, size_t ws_bytes, int N, int H, int W,
                        int C, int horizontal, int dil, erf_stream_t stream);
/* The weight gradients of up to four 3-tap convolutions that share (N, H, W, C) -- the four convs of one non_bottleneck_1d block -- in ONE
 * launch plus one reduce launch: persistent CTAs walk the jobs in order (csrc/conv1d_wgrad_group_tc.cu).  dw_j += as erf_conv1d_wgrad_tc. */
typedef struct { const void* x; const void* dy; float* dw; int horizontal, dil; } erf_wgrad_job;
size_t erf_conv1d_wgrad_group_tc_workspace(int C);
int erf_conv1d_wgrad_group_tc(const erf_wgrad_job* jobs, int njobs, void* workspace, size_t ws_bytes, int N, int H, int W, int C,
                              erf_stream_t stream);
/* Same with the reduction of the per-CTA partial sums INSIDE the kernel (no second launch): after a grid barrier every CTA sums its slice of
 * the slab over all CTAs' partials and adds it to dw.  counter = one zeroed 32-bit word per call (grid barrier); falls back to the two
 * launches when the grid would not be co-resident. */
int erf_conv1d_wgrad_tc_fused(const void* x, const void* dy, float* dw, void* workspace, size_t ws_bytes, unsigned* counter, int N,
                              int H, int W, int C, int horizontal, int dil, erf_stream_t stream);
/* out[c] += sum_p x[p*ld + coff + c]   (bias gradients) */
int erf_channel_sum(const void* x, int ld, int coff, int C, long long n_px, float* out, int dtype, erf_stream_t stream);

/* ---- generic tcgen05 implicit-GEMM convolution with software im2col ------------------------------------------------------
 * Serves the 3x3 stride-2 convs (eval/erfnet.py:16), the stride-2 transposed convs (l.102, 124; one launch per output
 * parity class, sub-pixel form), the 1x1 head (l.85) and all their data gradients.  One launch = one class of output
 * pixels sharing a tap list: pixel (n,y,x) of the enumeration grid (EN,EH,EW) reads input pixel
 * (in_step*y + tap_dy[t], in_step*x + tap_dx[t]) per tap (zero outside the image) and writes output pixel
 * (out_step*y + out_oy, out_step*x + out_ox), channels [out_c0, out_c0 + n_valid).  NHWC bf16 in/out. */
typedef struct {
  const void* in; int H, W, in_ld, Cin;            /* Cin: multiple of 16, <= 128 */
  int EN, EH, EW, in_step;
  int ntaps; int tap_dy[9]; int tap_dx[9];
  const void* w_img;                                /* erf_pack_im2col_weights image [ntaps][Cin/CH][NP][CH] */
  const float* bias;                                /* [n_valid] or NULL */
  void* out; int Ho, Wo, out_ld, out_c0, out_step, out_oy, out_ox;
  int NP, n_valid, relu;                            /* NP = padded N (16/32/64/128); n_valid % 8 == 0 */
  int units;                                        /* filled in by the library */
  int in_step_x, out_step_x;                        /* 0: same as in_step / out_step; else the column step differs from the row step */
  int bias_period;                                  /* 0: bias[n]; else bias[(out_c0 + n) % bias_period] (column-folded outputs) */
  float* nchw_out; int nchw_C;                      /* non-NULL (row-parity form with n_split = NP/2 <= 32 only): the epilogue writes the reference's
                                                     * NCHW fp32 tensor [EN][nchw_C][Ho][2*Wo] directly -- value pairs (b = 0, 1) of channel c as one
                                                     * float2 at columns 2x, 2x+1 -- instead of the NHWC tensor `out` (eval/erfnet.py:132 returns NCHW) */
} erf_im2col_desc;
/* Row-parity form of a stride-2 transposed conv: the two column parities b of an output row are folded into N (n = b*n_split + c),
 * the output is addressed as [N][Ho][Wo/2][2*ld] with out_step = 2 (rows) and out_step_x = 1, so each launch writes whole rows.
 * In the weight image tap t then holds w[..off[t]] for b = 0 and w[..off2[t]] for b = 1 (offset -1 = this tap does not feed that
 * parity: zeros). */
typedef struct { int ntaps; long long off[9]; int n_split; long long off2[9]; } erf_im2col_taps;
int erf_conv_im2col_tc(const erf_im2col_desc* d, erf_stream_t stream);
/* n_split == 0: img[t][kc][n][..] = w[(n0+n)*stride_n + k*stride_k + taps->off[t]] (fp32 torch weight, any of the Conv2d / ConvTranspose2d
 * forward / data-gradient orientations), zero for n0+n >= n_src or k >= k_src; bf16, swizzled for the kernel above. */
int erf_pack_im2col_weights(const float* w, long long stride_n, long long stride_k, int n0, int n_src, int k_src,
                            const erf_im2col_taps* taps, int Cin, int NP, void* img, erf_stream_t stream);

/* Weight gradient of the same layers on the tensor cores: dw[((cd)*Cg + cg)*KHW + tap_off[t]] += sum over the dense-side
 * pixels q of G[gather(q,t)][cg] * D[q][cd] (torch parameter layout for both nn.Conv2d -- G = x, D = dy -- and
 * nn.ConvTranspose2d -- G = dy, D = x).  gather(q,t) = (g_step*y + tap_dy[t], g_step*x + tap_dx[t]) on G's (GH, GW) grid.
 * The launch covers dense channels [d_c0, d_c0 + NP) of the pixel row (d_c_base = channel of the conv's channel 0). */
typedef struct {
  const void* g; int GH, GW, g_ld, Cg, CgP, g_copy_bytes, g_step;     /* gathered operand (M side) */
  const void* d; int EN, EH, EW, d_ld, d_c0, d_c_base, Cd, NP;        /* dense operand (N side), its pixel grid is the loop */
  int ntaps; int tap_dy[9]; int tap_dx[9]; int tap_off[9]; int KHW;
  int tail_bytes;                                                     /* filled in by the library */
  int g_step_x;                                                       /* column step of the gather (0: = g_step) */
  int fold, fold_c;   /* fold > 0: G is a column-folded view [..][GW][fold * 2]: gathered channel cg = b * fold + c holds channel c of the column
                       * parity b; only c < fold_c exists in the parameter, which is indexed ((cd)*fold_c + c)*KHW + tap_off[t] + b */
} erf_wgrad_im2col_desc;
size_t erf_conv_wgrad_im2col_tc_workspace(void);
int erf_conv_wgrad_im2col_tc(const erf_wgrad_im2col_desc* d, float* dw, void* workspace, size_t ws_bytes, erf_stream_t stream);

/* ---- ImageNet classification tail (imagenet/erfnet_imagenet.py:89-110, imagenet/main.py:80-85, 168-189, 288-301) ------------------- */
/* nn.AvgPool2d(k, stride 1, padding 0) over NHWC activations: out fp32 [N, H-k+1, W-k+1, C]; bwd writes every channel of dx (pitch ld,
 * channels >= C as zeros). */
int erf_avgpool_fwd(const void* in, int N, int H, int W, int ld, int C, int k, float* out, int dtype, erf_stream_t stream);
int erf_avgpool_bwd(const float* dout, int N, int H, int W, int ld, int C, int k, void* dx, int dtype, erf_stream_t stream);
/* nn.Linear: out[M,N] = x[M,K] w[N,K]^T + bias[N] (fp32); bwd: dx[M,K] = dout w (NULL: skipped), dw[N,K] += dout^T x, db[N] += column sums. */
int erf_linear_fwd(const float* x, const float* w, const float* bias, float* out, int M, int N, int K, erf_stream_t stream);
int erf_linear_bwd(const float* x, const float* w, const float* dout, float* dx, float* dw, float* db, int M, int N, int K,
                   erf_stream_t stream);
/* nn.CrossEntropyLoss() over logits [N,C] fp32 and int64 targets: acc[0] += sum of -log_softmax[target], acc[1] += #rows whose target is
 * not ignore_index (caller-zeroed; loss = acc[0] / acc[1]); lse[N] = log-sum-exp per row; rank[N] = position of the target logit in the
 * descending order of its row (ties: lower class first; -1 for ignored rows): top-k precision = #(rank < k) / N (imagenet/main.py:288-301). */
int erf_ce_rows_fwd(const float* logits, const long long* target, int N, int C, long long ignore_index, float* lse, int* rank, float* acc,
                    erf_stream_t stream);
/* dlogits = go[0] * (softmax - onehot) / acc[1] for valid rows, 0 for ignored rows (go NULL: 1). */
int erf_ce_rows_bwd(const float* logits, const long long* target, int N, int C, long long ignore_index, const float* lse,
                    const float* acc, const float* go, float* dlogits, erf_stream_t stream);
/* torch.optim.SGD(lr, momentum, weight_decay) over one flat fp32 buffer: d = grad_scale * g + wd * p; buf = first_step ? d : momentum * buf + d;
 * p -= lr * buf (momentum == 0: p -= lr * d, buf unused). */
int erf_sgd_step(float* p, const float* g, float* buf, long long n, float lr, float momentum, float weight_decay, int first_step,
                 float grad_scale, erf_stream_t stream);
/* Same for CUDA-graph replay: hyper_dev = fp32 [2] = {lr, weight_decay} in device memory (imagenet/main.py:276-286 changes lr every 30
 * epochs); buf must start zeroed (momentum * 0 + d = torch's first-step initialisation). */
int erf_sgd_step_graph(float* p, const float* g, float* buf, long long n, const float* hyper_dev, float momentum, float grad_scale,
                       erf_stream_t stream);

/* ---- mean teacher (train/main.py:140-142, 384-390): teacher = teacher * (1 - rate) + rate * student over flat fp32 buffers ----------- */
int erf_ema_update(float* teacher, const float* student, long long n, float rate, erf_stream_t stream);

/* ---- I/O side ----------------------------------------------------------------------------------------------------------------- */
/* uint8 HWC pixels [n_px, 3] -> NHWC activations [n_px, ld] (channels >= 3 zero): v = u8 / 255 (torchvision ToTensor, train/dataset.py:113-114),
 * then (v - mean[c]) / std[c] when mean / std (device fp32 [3]) are given.  Replaces ToTensor + the NCHW fp32 -> NHWC conversion. */
int erf_ingest_u8(const void* in, long long n_px, int ld, const float* mean, const float* stdv, void* out, int dtype, erf_stream_t stream);
/* out[i] = lut[in[i]] for 0 <= in[i] < lut_n, else in[i]; in uint8 or int64, out int64: ToLabel + Relabel (eval/transform.py:55-70), and the
 * trainId -> id table of eval/eval_cityscapes_server.py:39-63. */
int erf_label_lut(const void* in, int elem_bytes, long long n, const long long* lut, int lut_n, long long* out, erf_stream_t stream);
/* Colorize (eval/transform.py:73-93): out uint8 [n_img, 3, hw], out[:, ch, p] = cmap[label][ch] for label < n_colors (cmap uint8 [n_colors,3]), else 0. */
int erf_colorize(const void* labels, int elem_bytes, long long n_img, long long hw, const unsigned char* cmap, int n_colors,
                 unsigned char* out, erf_stream_t stream);

/* ---- tcgen05 / TMEM / TMA building blocks (sm_100a) ------------------------------------------------------ */
/* Probe used by the GPU tests to pin the UMMA shared-memory descriptor semantics the fused kernels rely on:
 * D[128 x N] = A[rows row_shift .. row_shift+127] (K-major, K = 16*ksteps) * B[N x K]^T, layout: 0 = no swizzle
 * (core-matrix interleave), 2 = 128-byte swizzle.  A_rows x K and N x K bf16 row-major in, D fp32 row-major out. */
int erf_umma_probe(const void* a, int a_rows, const void* b, int n, int ksteps, int row_shift, int layout,
                   float* d, int iters, long long* cycles, erf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ERFNET_B200_H_ */
