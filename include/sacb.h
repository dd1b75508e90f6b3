/* libsac_b200 -- C ABI of the B200-native SAC target step (sm_100a).
 *
 * The reference (visinf/da-sac) has no FFI layer: its hot path is the Python
 * module contract models.get_model(...) -> SAC.forward(...) consumed by
 * train.py:88-104,211-250.  These entry points are what a binding for that
 * path binds to (INTEGRATION.md shows the ctypes stub); each one cites the
 * reference code it replaces (paths relative to /root/reference).
 *
 * Conventions: every function returns 0 on success, <0 on error
 * (sacb_last_error() gives the message); nothing throws across the ABI; the
 * caller owns every buffer; all work is asynchronous on `stream`
 * (a cudaStream_t passed as void*); device pointers must be 16-byte aligned
 * (activation planes 128-byte).  No torch types appear here.
 *
 * Activation layout ("split planes"): an fp32 activation tensor x[N,H,W,C]
 * (NHWC) is held as two bf16 planes hi = bf16(x), lo = bf16(x - hi), so that
 * x = hi + lo to ~2^-17 relative; convolutions evaluate
 * hi*hi + lo*hi + hi*lo on the tcgen05 tensor cores with fp32 accumulation in
 * TMEM ("bf16x3", fp32-equivalent accuracy; SURVEY.md section 7 precision).
 */
#ifndef SACB_H_
#define SACB_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SACB_ABI_VERSION 6

const char* sacb_last_error(void);
int sacb_abi_version(void);
/* number of kernels launched by this library since process start (bench.py "gpu_launches") */
int64_t sacb_launch_count(void);

/* ---------------------------------------------------------------- convolution as implicit GEMM
 * Replaces nn.Conv2d forward (models/deeplabv2.py:59-70,107,122,147) fused with the
 * eval-mode SyncBatchNorm affine, residual add and ReLU that follow it
 * (deeplabv2.py:77-99), and -- with flipped/transposed weights -- the autograd
 * data-gradient of the same conv (train.py:232).
 *
 *   acc[m, k]  = sum_{r,s,c} X[n, p*stride - pad + r*dil, q*stride - pad + s*dil, c] * Wt[r*S+s][k][c]
 *   v          = acc * scale[k] + shift[k]            (if scale != NULL)
 *   v         += add_f32[m*K + k]                      (if add_f32)
 *   v         += add_hi[m*K+k] + add_lo[m*K+k]         (if add_hi)
 *   v          = max(v, 0)                             (if relu)
 *   v          = mask_hi[m*K+k] > 0 ? v : 0            (if mask_hi; ReLU backward)
 *   out_hi/out_lo[m*K+k] = split(v); out_f32[m*K+k] = v; out_nchw[((n*k_valid+k)*P+p)*Q+q] = v
 *   colsum[k] += sum_m v                               (if colsum; caller zero-fills; fp32 atomics, one per CTA and k)
 * with m = (n*P + p)*Q + q.                                                          */
#define SACB_PRECISION_BF16X3 0   /* lo*hi + hi*lo + hi*hi: the parity mode every golden test runs in */
#define SACB_PRECISION_BF16 1     /* hi*hi only: SURVEY.md's "fast mode", AMP-class accuracy (~1e-2 on logits), reported separately */
typedef struct SacbConvGemm {
  uint32_t size;              /* sizeof(SacbConvGemm), ABI versioning */
  int32_t N, H, W, C;         /* input NHWC; C % 64 == 0 */
  int32_t K;                  /* output channels as stored in wt (multiple of the N tile: 32, 64 or 128) */
  int32_t k_valid;            /* channels actually written (<= K) */
  int32_t R, S, stride, dil, pad;
  int32_t P, Q;               /* output spatial size */
  const void* x_hi; const void* x_lo;     /* bf16 [N,H,W,C] */
  const void* wt_hi; const void* wt_lo;   /* bf16 [R*S][K][C] */
  const float* scale; const float* shift; /* [K] or NULL */
  const float* add_f32;                   /* [M,K] or NULL */
  const void* add_hi; const void* add_lo; /* bf16 [M,K] or NULL */
  const void* mask_hi;                    /* bf16 [M,K] or NULL */
  int32_t relu;
  void* out_hi; void* out_lo;             /* bf16 [M,K] or NULL */
  float* out_f32;                         /* [M,K] or NULL */
  float* out_nchw;                        /* [N,k_valid,P,Q] or NULL */
  float* colsum;                          /* [K] or NULL: BN d(beta) of the unit that consumes this gradient */
  int32_t precision;                      /* SACB_PRECISION_*: 0 = bf16x3 split (parity mode, fp32-equivalent), 1 = single-pass
                                             bf16 on the hi planes only (fast mode; the lo planes are still written) */
  int32_t unit_scale;                     /* caller's promise: scale == NULL or scale[c] == 1 for every c (the BN scale lives in
                                             the weights, SacbPrepItem.fold_wf).  The residual add_hi/add_lo may then be added
                                             BEFORE the affine, which lets the CTA-pair kernel feed it to the tensor core through
                                             TMA (residual x identity into the same TMEM accumulator) instead of loading it in
                                             the epilogue.  0 = no promise (always correct). */
} SacbConvGemm;
int sacb_conv_gemm(const SacbConvGemm* d, void* stream);

/* Filter gradient (autograd of nn.Conv2d w.r.t. weight, train.py:232):
 *   dw[split][k][r*S+s][c] = sum_{(n,p,q) in split} G[n,p,q,k] * X[n, p*stride-pad+r*dil, q*stride-pad+s*dil, c]
 * Deterministic split-K: every split writes its own partial plane; sacb_wgrad_finalize sums them in order.
 * sacb_conv_wgrad_splits(d) returns the number of planes (>= 1) so the caller can size dw. G, X: split planes. */
typedef struct SacbConvWgrad {
  uint32_t size;
  int32_t N, H, W, C;         /* input X NHWC, C % 64 == 0 */
  int32_t K;                  /* channels of G as stored (multiple of 64) */
  int32_t k_valid;            /* rows of dw written */
  int32_t R, S, stride, dil, pad;
  int32_t P, Q;
  const void* x_hi; const void* x_lo;   /* bf16 [N,H,W,C] */
  const void* g_hi; const void* g_lo;   /* bf16 [N,P,Q,K] */
  float* dw;                            /* fp32 [splits][k_valid][R*S][C] */
  int32_t splits;                       /* requested split-K factor over pixels; 0 = auto */
  int32_t precision;                    /* as in SacbConvGemm */
} SacbConvWgrad;
int sacb_conv_wgrad_splits(const SacbConvWgrad* d);
int sacb_conv_wgrad(const SacbConvWgrad* d, void* stream);

/* ---------------------------------------------------------------- elementwise / layout kernels
 * (declared in sacb_elem.cu; see DESIGN.md for the reference lines each replaces) */

/* fp32 NCHW image -> stem conv 7x7 s2 p3 (3->64) + BN affine + ReLU -> split planes NHWC
 * (deeplabv2.py:160-163). w: fp32 [64][3][7][7]. */
int sacb_stem_fwd(const float* x_nchw, const float* w, const float* scale, const float* shift,
                  void* out_hi, void* out_lo, int N, int H, int W, int P, int Q, void* stream);
/* dW of the stem conv: g split planes [N,P,Q,64] (already masked by ReLU), x fp32 NCHW -> dw[64][3][7][7] (+=) */
int sacb_stem_wgrad(const float* x_nchw, const void* g_hi, const void* g_lo, float* dw,
                    int N, int H, int W, int P, int Q, void* stream);
/* Tensor-core first conv (3 input channels): A[n,p,q, (c*R+r)*R+s] = x[n,c,stride*p-pad+r,stride*q-pad+s] (3*R*R taps
 * zero-padded to KP columns) turns the conv into a 1x1 sacb_conv_gemm with C = KP (weights from sacb_stem_pack_weight:
 * [K][KP]) and its filter gradient into a 1x1 sacb_conv_wgrad whose partial planes sacb_stem_unpack_wgrad sums into
 * dwraw[K][3*R*R]. ResNet stem: R=7, stride 2, pad 3, KP=192; VGG features.0: R=3, stride 1, pad 1, KP=64. */
int sacb_stem_im2col(const float* x_nchw, void* a_hi, void* a_lo, int N, int H, int W, int P, int Q, int R, int stride,
                     int pad, int KP, void* stream);
int sacb_stem_pack_weight(const float* w, void* hi, void* lo, int K, int taps, int KP, void* stream);
int sacb_stem_unpack_wgrad(const float* parts, int splits, float* dwraw, int K, int taps, int KP, void* stream);
/* MaxPool2d(2, 2) of torchvision vgg16 (models/deeplabv2.py:238-260) on split planes, and its backward (+ReLU mask) */
int sacb_maxpool2_fwd(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, uint8_t* idx,
                      int N, int H, int W, int C, int P, int Q, void* stream);
int sacb_maxpool2_bwd(const float* g_out, const uint8_t* idx, const void* in_hi, void* gin_hi, void* gin_lo,
                      int N, int H, int W, int C, int P, int Q, void* stream);
/* MaxPool2d(3, 2, 1, ceil_mode=True) on split planes (deeplabv2.py:126); idx = argmax tap (uint8) for backward */
int sacb_maxpool_fwd(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, uint8_t* idx,
                     int N, int H, int W, int C, int P, int Q, void* stream);
/* g_in[n,h,w,c] = (sum of g_out over windows whose argmax is (h,w)) masked by in_hi > 0 -> split planes */
int sacb_maxpool_bwd(const float* g_out, const uint8_t* idx, const void* in_hi, void* gin_hi, void* gin_lo,
                     int N, int H, int W, int C, int P, int Q, void* stream);
/* out = split( mask_hi>0 ? (a + b) : 0 ), a/b fp32 [n] (b may be NULL); mask_hi may be NULL */
int sacb_add_mask_split(const float* a, const float* b, const void* mask_hi, void* out_hi, void* out_lo,
                        int64_t n, void* stream);
/* scatter a compact stride-2 1x1 data gradient back to the full map:
 * out[n,h,w,c] = (h,w even ? a[n,h/2,w/2,c] (+ b) : 0) masked by mask_hi > 0 -> split planes */
int sacb_scatter2_mask_split(const float* a, const float* b, const void* mask_hi, void* out_hi, void* out_lo,
                             int N, int H, int W, int C, int P, int Q, void* stream);
/* colsum[c] = sum_m (hi[m,c] + lo[m,c])  (BN d(beta), conv d(bias)); fp32 atomics into zeroed colsum */
int sacb_colsum(const void* hi, const void* lo, float* colsum, int64_t M, int C, void* stream);
/* weight preparation (per optimiser step): OIHW fp32 -> fprop planes wf[R*S][Kf][C] (rows >= K zero) and,
 * if wt_* != NULL, dgrad planes wt[R*S][C][Kt] = w[k][c][R-1-r][S-1-s] * scale[k] (scale NULL = 1; k >= K zero) */
int sacb_prep_weight(const float* w_oihw, const float* scale, int K, int C, int R, int S, int Kf, int Kt,
                     void* wf_hi, void* wf_lo, void* wt_hi, void* wt_lo, void* stream);
/* BN fold (basenet.py:97-100 eval-mode statistics, trainable affine), optionally with the conv bias of VGG convs:
 * scale = gamma * rsqrt(var + eps), shift = beta + (conv_bias - mean) * scale; gamma == NULL: scale = 1, shift = conv_bias */
int sacb_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                 const float* conv_bias, float* scale, float* shift, int C, void* stream);
/* finalize a conv+BN unit's parameter gradients from the raw filter gradient (DESIGN.md "BN backward"):
 * dwraw[k][rs][c] = sum over `splits` partial planes (plane stride K*R*S*C);
 * dw_oihw[k][c][r][s] = scale[k] * dwraw[k][rs][c];
 * dgamma[k] = (sum_{rs,c} w[k][c][r][s] * dwraw[k][rs][c] + (conv_bias[k] - mean[k]) * dbeta[k]) * rsqrt(var[k]+eps);
 * dbias[k] = scale[k] * dbeta[k] (if dbias; scale NULL = 1) */
int sacb_wgrad_finalize(const float* dwraw, const float* w_oihw, const float* scale, const float* mean,
                        const float* var, float eps, const float* dbeta, float* dw_oihw, float* dgamma,
                        const float* conv_bias, float* dbias, int K, int C, int R, int S, int splits, void* stream);

/* Multi-layer variants: ONE launch for a whole network.  `items_dev` is a DEVICE array of per-layer descriptors (the
 * arguments of sacb_bn_fold + sacb_prep_weight, resp. sacb_wgrad_finalize, with the same meaning); `block_begin_dev[i]` is
 * the first thread block of item i (prefix sums of sacb_prep_item_blocks(...), resp. of K), `total_blocks` their sum.
 * sacb_prepare_batched: scale == NULL skips the fold, wf_hi == NULL skips the planes (first conv: packed separately);
 * the dgrad planes are pre-multiplied by gamma * rsqrt(var + eps) when gamma != NULL.
 * sacb_wgrad_finalize_batched additionally copies d(beta) to dbeta_out (the BN bias gradient in the flat buffer). */
typedef struct SacbPrepItem {
  const float* w;                                   /* OIHW fp32 */
  const float* gamma; const float* beta; const float* mean; const float* var; const float* conv_bias;
  float* scale; float* shift;                       /* [K] folded affine out */
  void* wf_hi; void* wf_lo; void* wt_hi; void* wt_lo;
  int32_t K, C, R, S, Kf, Kt;
  int32_t fold_wf;                                  /* 1: the fprop planes carry the folded BN scale too (wf = w * gamma / sigma), the
                                                       caller then runs the conv with a unit scale (SacbConvGemm.unit_scale) */
  int32_t reserved;
} SacbPrepItem;
int sacb_prep_item_blocks(int K, int C, int R, int S, int Kf, int Kt, int with_wf, int with_wt);
int sacb_prepare_batched(const SacbPrepItem* items_dev, const int32_t* block_begin_dev, int n_items, int total_blocks,
                         float eps, void* stream);
typedef struct SacbFinalizeItem {
  const float* dwraw; const float* w; const float* scale; const float* mean; const float* var; const float* dbeta;
  float* dw; float* dgamma; const float* conv_bias; float* dbias; float* dbeta_out;
  int32_t K, C, RS, splits;
} SacbFinalizeItem;
int sacb_wgrad_finalize_batched(const SacbFinalizeItem* items_dev, const int32_t* block_begin_dev, int n_items,
                                int total_blocks, float eps, void* stream);

/* ---------------------------------------------------------------- ASPP head as a tap-unrolled 1x1 GEMM
 * Classifier_Module (deeplabv2.py:101-116): sum of four 3x3 dilated convs 2048 -> 19 (+ biases).
 *   Z[pix, (i*9 + r*3+s)*19 + k] = sum_c X[pix,c] * W_i[k,c,r,s]   via sacb_conv_gemm (1x1, K = sacb_aspp_jpad() = 768)
 *   logits[n,k,p,q] = sum_i b_i[k] + sum_{i,r,s} Z[(n, p+(r-1)d_i, q+(s-1)d_i), (i*9+r*3+s)*19 + k]   (sacb_aspp_gather)
 * backward: Gcol[pix, (i*9+rs)*19+k] = g[n,k,p-(r-1)d_i,q-(s-1)d_i] (sacb_aspp_gcol) feeds sacb_conv_gemm (data
 * gradient, weights wt[c][j]) and sacb_conv_wgrad (filter gradient), unpacked by sacb_aspp_unpack_wgrad. */
int sacb_aspp_jpad(void);
int sacb_aspp_pack_weights(const float* w0, const float* w1, const float* w2, const float* w3, int C,
                           void* wf_hi, void* wf_lo, void* wt_hi, void* wt_lo, void* stream);
int sacb_aspp_gather(const float* Z, const float* b0, const float* b1, const float* b2, const float* b3,
                     const int32_t* dil4 /* host */, float* out_nchw, int N, int P, int Q, void* stream);
int sacb_aspp_gcol(const float* g_nchw, const int32_t* dil4 /* host */, void* hi, void* lo, int N, int P, int Q,
                   void* stream);
int sacb_aspp_unpack_wgrad(const float* parts, int splits, int C, float* dw0, float* dw1, float* dw2, float* dw3,
                           void* stream);

/* ---------------------------------------------------------------- SAC tail (models/sac.py)
 * teacher logits -> pseudo labels; replaces SAC._refine + _update_running_conf + _avg_pool +
 * _pseudo_labels_probs (sac.py:104-117,151-187,238-313). */
typedef struct SacbTail {
  uint32_t size;
  int32_t BT, T, C, h, w, H, W;
  const float* teacher_logits;   /* [BT,C,h,w] fp32 NCHW */
  const int64_t* y;              /* [BT,H,W]; -1 = augmentation padding (sac.py:337) */
  const float* affine;           /* [BT,2,3] */
  const float* affine_inv;       /* [BT,2,3] */
  float* running_conf;           /* [C] in/out (updated if training) */
  int32_t training, discount;
  float beta, stat_momentum, conf_upper, conf_lower;
  /* workspace (element counts from sacb_tail_*_elems) */
  float* probs;                  /* [BT,H*W,CP] masked teacher probabilities, pixel-major */
  float* pooled;                 /* [BT/T,H*W,CP2] averaged probs (0..C-1) + valid mask (C) */
  float* part_sums;              /* [BT*ceil(H*W/256), C] partial class sums */
  float* peaks;                  /* [BT,C] */
  /* outputs */
  float* conf;                   /* [BT,H,W] */
  uint8_t* idx;                  /* [BT,H,W] */
  uint8_t* labels;               /* [BT,H,W]  (255 = ignore) */
  float* conf_mean;              /* [H,W]  batch-mean confidence (the [B,B,H,W] broadcast of sac.py:148) */
  float* thresholds;             /* [BT,C] */
  float* refined;                /* optional [BT,C,H,W] teacher_refined, NULL to skip */
  /* fractional groups (train.py:185-209, sac.py:198-216: a group's T views spread over several ranks, here T = the
   * views THIS rank holds): 0 = whole tail; 1 = stop after writing the un-normalised reference-frame sums to `pooled`
   * (the caller sum-reduces `pooled` over the ranks sharing the group); 2 = resume: normalise `pooled`, labels;
   * 3 = diagnostics after 0 or 2 of the same forward: only warp `pooled` back once more into `refined` (net_outs
   * "teacher_refined"); running_conf, thresholds, labels and conf_mean are left alone and no exchange is needed. */
  int32_t phase;
  /* MODEL.CONF_POOL / CONF_POOL_ON (core/config.py:150-151): 0 = avg_pool (sac.py:238-269, default), 1 = minentropy_pool
   * (sac.py:218-236: every view receives the distribution of the group's lowest-entropy view), 2 = pooling off
   * (_refine(pool=False), sac.py:284-285: labels straight from each view's masked teacher probabilities) */
  int32_t pool_mode;
} SacbTail;
int sacb_teacher_tail(const SacbTail* d, void* stream);
size_t sacb_tail_part_sums_elems(int BT, int C, int H, int W);
size_t sacb_tail_probs_elems(int BT, int C, int H, int W);
size_t sacb_tail_pooled_elems(int G, int C, int H, int W);

/* student loss (deeplabv2.py:217-224 + sac.py:134-149): fused upsample + log-softmax + weighted NLL */
typedef struct SacbLoss {
  uint32_t size;
  int32_t BT, C, h, w, H, W;
  const float* logits;           /* [BT,C,h,w] student */
  const int64_t* y;              /* [BT,H,W] ground truth as given (-1 and 255 ignored) */
  const uint8_t* labels;         /* pseudo labels */
  const float* conf_mean;        /* [H,W] */
  const float* running_conf;     /* [C] */
  float focal_p;
  float* losses;                 /* [2] : loss_ce, self_ce (zeroed by the call) */
  double* scratch;               /* [2] */
  /* backward */
  float grad_scale;              /* d(total)/d(self_ce), e.g. LR_TARGET */
  float* dlogits;                /* [BT,C,h,w] or NULL */
  /* optional backward workspace.  grad_rows alone (W <= 1216): the per-pixel gradient of an up-sampled row stays in shared
   * memory and only its row adjoint is written, then the column adjoint -- the full-resolution gradient is never
   * materialised.  Both non-NULL (any W): two-stage form through grad_px.  Neither: the gather kernel. */
  float* grad_px;                /* [BT,C,H,W] */
  float* grad_rows;              /* [BT,C,H,w] */
} SacbLoss;
int sacb_student_loss_fwd(const SacbLoss* d, void* stream);
/* dlogits = grad_scale * d(self_ce)/d(logits); with d->labels == NULL: grad_scale * d(loss_ce)/d(logits) (CE against y) */
int sacb_student_loss_bwd(const SacbLoss* d, void* stream);
/* bilinear align_corners=True upsample [B,C,h,w] -> [B,C,H,W] (F.interpolate, deeplabv2.py:217) */
int sacb_upsample(const float* in, float* out, int B, int C, int h, int w, int H, int W, void* stream);

/* FCN-8s score fusion (fcn.py:107-134): out = bilinear_up(in, align_corners=True) (+ addend); its adjoint; the
 * NCHW fp32 -> NHWC split-plane conversion that feeds the 19-class score convs' gradients; Dropout2d on split planes */
int sacb_upsample_add(const float* in, const float* addend, float* out, int B, int C, int h, int w, int H, int W, void* stream);
int sacb_upsample_bwd(const float* g_out, float* g_in, int B, int C, int h, int w, int H, int W, void* stream);
int sacb_nchw_to_planes(const float* g_nchw, void* hi, void* lo, int N, int C, int P, int Q, int Cpad, void* stream);
int sacb_channel_scale(void* hi, void* lo, const float* m /* [N,C] */, int N, int P, int Q, int C, void* stream);

/* ---------------------------------------------------------------- optimiser-side multi-tensor kernels
 * flat fp32 buffers with a segment table: seg_ranges[2*i], seg_ranges[2*i+1] = [begin, end) of tensor i. */
/* SAC._momentum_update (sac.py:83-102): out[0] = sum_seg ||slow-fast||_2 ; if update: slow = m*slow+(1-m)*fast */
int sacb_ema_norm(float* slow, const float* fast, const int64_t* seg_ranges, int nseg, float momentum,
                  int update, float* seg_sq, float* out, void* stream);
/* torch.optim.SGD step (base_trainer.py:61-66): per-segment lr / weight decay, momentum buffer in place */
int sacb_sgd(float* p, const float* g, float* mom, const int64_t* seg_ranges, const float* seg_lr,
             const float* seg_wd, int nseg, float momentum, int first_step, void* stream);

/* ---------------------------------------------------------------- target-view augmentation on the device
 * Replaces DataTarget.__getitem__'s PIL pipeline (datasets/dataloader_target.py:264-306; datasets/tf_target.py:140-156
 * GuidedRandHFlip, :158-239 MaskRandScaleCrop, :331-390 blur / jitter / greyscale, :32-98 to-tensor / normalise / mask):
 * from G base crops (uint8 RGB, already at crop size) to G*T views.  Per view v the host supplies view_params[v][16]:
 *   [0] flip (+1 / -1)   [1] crop top  [2] crop left  [3] crop height  [4] crop width   (window in the flipped base crop;
 *                                       it may extend outside for zoom-out, then the outside is black and masked)
 *   [5] Gaussian blur sigma (0 = off)   [6] colour jitter on/off   [7..10] op order (0 brightness, 1 contrast, 2 saturation,
 *   3 hue)   [11] brightness  [12] contrast  [13] saturation factors  [14] hue shift in [-0.5, 0.5]   [15] greyscale on/off
 * Outputs follow the reference's batch_target: frames1 (photometric copy), gt (-1 inside padding, else the label or 255),
 * frames2 (clean copy); both normalised with mean/std and zeroed inside padding.  affine / affine_inv are computed on the
 * host from the same parameters (da_sac_b200/augment.py, dataloader_target.py:220-262). */
#define SACB_AUG_NPARAM 16
typedef struct SacbAug {
  uint32_t size;
  int32_t G, T, H, W;
  const uint8_t* base;          /* [G,H,W,3] RGB */
  const uint8_t* base_mask;     /* [G,H,W] > 0 = padding (MaskRandCrop), or NULL */
  const uint8_t* base_label;    /* [G,H,W] or NULL (= 255: unlabelled target data) */
  const float* view_params;     /* device [G*T][SACB_AUG_NPARAM] */
  float mean[3], std[3];
  /* workspace */
  uint8_t* raw;                 /* [G*T,H,W,3] resized 8-bit views */
  uint8_t* mask;                /* [G*T,H,W] */
  float* tmp;                   /* [G*T,H,W,3] */
  float* levels;                /* [G*T,H,W,3] */
  uint64_t* grey_sum;           /* [G*T] */
  /* outputs */
  float* frames1;               /* [G*T,3,H,W] */
  int64_t* gt;                  /* [G*T,H,W] */
  float* frames2;               /* [G*T,3,H,W] */
} SacbAug;
int sacb_target_augment(const SacbAug* d, void* stream);

/* ---------------------------------------------------------------- gradient all-reduce fused with SGD over NVLink peer memory
 * Replaces DistributedDataParallel's gradient all-reduce (train.py:104,232; sum over ranks / world) + torch.optim.SGD.step()
 * (train.py:233) for the data-parallel target step with ONE kernel per rank and no NCCL call:
 *   rank r: for its 1/world slice of the flat buffers:  g = (sum_{p=0..world-1} grads[p][i]) / world  (peer loads, fixed order)
 *           SGD with momentum / weight decay exactly as sacb_sgd (momentum buffer touched on the slice only)
 *           params[p][i] = new value for EVERY rank p (peer stores)
 * Ranks synchronise through epoch flags in peer-mapped memory (flags[p]: sacb_p2p_flag_words() zero-initialised uint32 per
 * rank; the epoch counter is device-resident, so the call is CUDA-graph capturable).  Waits are bounded (trap on timeout).
 * All ranks must call it once per step with the same n / segment table; world == 1 degenerates to sacb_sgd.
 * Buffers shared between processes come from sacb_symm_alloc (cudaMalloc, zero-filled) and travel as CUDA IPC handles. */
#define SACB_P2P_MAX_WORLD 8
#define SACB_IPC_HANDLE_BYTES 64
int sacb_symm_alloc(size_t bytes, void** dptr);
int sacb_symm_free(void* dptr);
int sacb_ipc_export(const void* dptr, void* handle64 /* out: SACB_IPC_HANDLE_BYTES */);
int sacb_ipc_import(const void* handle64, void** peer_ptr);
int sacb_ipc_close(void* peer_ptr);
int sacb_p2p_flag_words(void);
typedef struct SacbAllreduceSgd {
  uint32_t size;
  int32_t world, rank;
  float* const* grads;          /* host array [world]: rank p's flat gradient buffer as mapped in THIS process */
  float* const* params;         /* host array [world]: rank p's flat parameter buffer */
  uint32_t* const* flags;       /* host array [world]: rank p's flag words */
  float* mom;                   /* local momentum buffer [n] */
  const int64_t* seg_ranges; const float* seg_lr; const float* seg_wd;   /* device, as sacb_sgd; segments sorted by begin */
  int32_t nseg;
  int64_t n;                    /* elements of the flat buffers (multiple of 4) */
  float momentum;
  int32_t first_step;
  /* optional NVLS path (both or neither): multicast mappings of the SAME gradient / parameter allocations (one address that
   * reaches every rank's replica, e.g. torch.distributed._symmetric_memory's multicast_ptr).  The owner rank then reads the sum
   * of its slice with multimem.ld_reduce (the NVSwitch adds) and writes the new weights with multimem.st.  The summation order
   * inside the switch is not specified: replicas still agree bit for bit (one owner per element), runs may differ in the last bit. */
  float* mc_grads;
  float* mc_params;
} SacbAllreduceSgd;
int sacb_allreduce_sgd(const SacbAllreduceSgd* d, void* stream);

/* ---------------------------------------------------------------- training-mode batch norm (ABN baseline, cfg.MODEL.BASELINE)
 * Replaces torch.nn.SyncBatchNorm in TRAINING mode (models/deeplabv2.py:15,28-31,60-71,124,149,183; active when
 * models/__init__.py:29 sets freeze_bn = False) around the same tcgen05 convolutions, run with a raw epilogue:
 *   forward : z = conv(x);  (sum z, sum z^2) -> [all-reduce over ranks] -> mean, invstd, running-stat update;
 *             y = relu?((z - mean) * gamma * invstd + beta (+ residual))
 *   backward: g = dL/dy_pre;  (sum g, sum g * xhat) -> [all-reduce] -> d gamma, d beta (local sums),
 *             dz = gamma * invstd * (g - mean(g) - xhat * mean(g * xhat)),   xhat = (z - mean) * invstd
 * Planes are bf16 split planes [M][C], C % 8 == 0.  Moments are deterministic: per-block partial sums in double, added in
 * block order.  `partials` needs sacb_bn_moments_partial_elems(M, C) doubles; `sums` is [2][C] doubles.              */
size_t sacb_bn_moments_partial_elems(int64_t M, int C);
/* mode 0: sums = (sum a, sum a^2) with a = z;  mode 1: sums = (sum a, sum a * xhat(z)) with a = g */
int sacb_bn_moments(const void* a_hi, const void* a_lo, const void* z_hi, const void* z_lo, const float* mean,
                    const float* invstd, int mode, int64_t M, int C, double* partials, double* sums, void* stream);
/* count = elements per channel over ALL ranks.  Writes mean, invstd, scale = gamma * invstd and, if running_mean != NULL,
 * running = (1 - momentum) * running + momentum * batch (unbiased variance), as nn.BatchNorm2d does in training mode. */
int sacb_bn_train_finalize(const double* sums, double count, const float* gamma, float eps, float momentum,
                           float* running_mean, float* running_var, float* mean, float* invstd, float* scale, int C,
                           void* stream);
int sacb_bn_apply(const void* z_hi, const void* z_lo, const float* mean, const float* scale, const float* beta,
                  const void* res_hi, const void* res_lo, int relu, void* y_hi, void* y_lo, int64_t M, int C, void* stream);
/* d gamma / d beta from this rank's sums; coef [3][C] = (gamma * invstd, mean(g), mean(g * xhat)) from the global sums */
int sacb_bn_bwd_finalize(const double* sums_local, const double* sums_global, double count_global, const float* gamma,
                         const float* invstd, float* dgamma, float* dbeta, float* coef, int C, void* stream);
/* may run in place (dz == g) */
int sacb_bn_bwd_apply(const void* g_hi, const void* g_lo, const void* z_hi, const void* z_lo, const float* mean,
                      const float* invstd, const float* coef, void* dz_hi, void* dz_lo, int64_t M, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SACB_H_ */
