/*
 * libstreamyolo_sm100.so -- C ABI of the B200-native StreamYOLO hot path.
 *
 * This is the drop-in boundary of SURVEY.md section 8(b): plain `extern "C"` entry
 * points, raw device pointers + sizes + a cudaStream_t, no torch types.  The host
 * side (streamyolo_b200/model/*.py, a mirror of the reference's exps/model API)
 * binds them with ctypes; INTEGRATION.md shows the stub a reference maintainer adds.
 *
 * Conventions
 *   - every function is asynchronous on `stream`, re-entrant, allocates nothing,
 *     never synchronises the device and never throws; it returns 0 or an SY_E*
 *     status and `sy_last_error_string()` describes the last failure of the
 *     calling thread.  The caller owns every buffer (PyTorch caching allocator).
 *   - activations are NHWC bf16 "views" (SyTensor): pixel (n,y,x) channel c lives
 *     at ptr + (((n*h + y)*w + x)*pitch + c) elements; pitch >= c lets a view be
 *     a channel slice of a wider concat buffer (pitch, slice offsets and c are
 *     multiples of 8 so that every pixel row is 16-byte aligned).
 *   - conv weights are bf16, packed [Cout][kh*kw][Cin] (K-major GEMM B operand).
 *   - the library requires an sm_100a device; there is no other code path.
 *
 * Each entry point cites the reference interface it replaces (paths relative to
 * /root/reference; "[yolox]" = the un-vendored yolox==0.3.0 dependency).
 */
#ifndef STREAMYOLO_SM100_H_
#define STREAMYOLO_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* sy_stream_t; /* == cudaStream_t */

enum {
  SY_OK = 0,
  SY_EINVAL = 1,   /* bad shape / alignment / null pointer */
  SY_EARCH = 2,    /* device is not sm_100 */
  SY_ELAUNCH = 3,  /* CUDA launch or driver error (see sy_last_error_string) */
  SY_EWORKSPACE = 4
};

typedef struct {
  void* ptr;      /* bf16 */
  int32_t n, h, w, c;
  int64_t pitch;  /* elements between consecutive pixels */
} SyTensor;

/* -------- runtime ---------------------------------------------------------- */
const char* sy_last_error_string(void);
int sy_version(void);
/* 0 when the current device is sm_100 and the driver exposes cuTensorMapEncodeTiled. */
int sy_check_device(void);

/* Mark [ptr, ptr + bytes) as a persisting-L2 access window for the kernels subsequently launched on `stream` (inherited by
 * the kernel nodes of a stream capture); bytes = 0 clears it.  *granted = the window the device allows (0: none).  The host
 * side (engine.py) draws the raw conv outputs of all train-mode BaseConvs that fit from one arena inside this window: the
 * normalise pass then reads them from L2 and the next layer overwrites them before they reach HBM. */
int sy_l2_persist_window(void* ptr, size_t bytes, float hit_ratio, size_t* granted, sy_stream_t stream);

/* -------- convolution (replaces [yolox] BaseConv.conv / nn.Conv2d, e.g.
 * exps/model/darknet.py:115-165, exps/model/dfp_pafpn.py:33-105,
 * exps/model/tal_head.py:55-104) ------------------------------------------------- */
enum { SY_CONV_RAW = 0, SY_CONV_FUSED = 1 };

/* One BatchNorm parameter set covering output channels [c_begin, next segment's c_begin or Cout):
 * lets ONE conv launch serve two BaseConv modules that read the same input (CSPLayer conv1 | conv2). */
typedef struct {
  const float* gamma; const float* beta;          /* [c] */
  float* running_mean; float* running_var;        /* [c], updated in place (may be NULL) */
  int64_t* num_batches_tracked;                   /* += number of statistic groups (may be NULL) */
  int32_t c_begin;
} SyBnSegment;

typedef struct {
  SyTensor x;            /* input  */
  SyTensor y;            /* output: RAW -> conv result; FUSED -> silu(acc*scale+shift)(+res) */
  const void* w;         /* bf16 [Cout][kh*kw][Cin] */
  int32_t kh, kw;        /* 1 or 3 each, padding (k-1)/2 */
  int32_t stride;        /* 1 or 2 */
  int32_t mode;          /* SY_CONV_RAW / SY_CONV_FUSED */
  int32_t act;           /* FUSED: 1 = SiLU, 0 = identity */
  const float* scale;    /* FUSED: [Cout] (folded BatchNorm), may be NULL = 1 */
  const float* shift;    /* FUSED: [Cout], may be NULL = 0 */
  SyTensor res;          /* FUSED: optional residual added after the activation (ptr NULL = none) */
  /* ---- RAW mode: per-channel batch statistics of the stored output (sy_conv2d_tc only) ---- */
  int32_t split_n;       /* images >= split_n form statistics group 1 (0 or >= n: one group) */
  float* stat_partials;  /* [n_partials][Cout][2 groups][2 (sum, sumsq)] floats, one row per CTA (16B aligned), or NULL */
  int32_t n_partials;    /* >= sy_conv_stat_rows(); rows of CTAs that did not run are NOT written */
  int32_t* rows_written; /* out (host int, may be NULL): number of partial rows this launch writes */
  SyBnSegment bn[2];     /* bn[0].gamma != NULL: finalize BatchNorm in the kernel tail (1-2 parameter segments) */
  float momentum, eps;
  float* scale_shift;    /* [2 (scale|shift)][2 groups][Cout]: y = x*scale + shift, ready when the kernel ends */
  float* mean_invstd;    /* optional [2 (mean|invstd)][2 groups][Cout]: the batch statistics themselves, saved for
                          * sy_bn_act_backward (NULL: not written) */
  uint32_t* sync;        /* four zero-initialised counters (grid barriers); the kernel leaves them at zero */
  /* With bn[]: optional normalise + act (+ residual) pass INSIDE the same launch (after a second grid barrier
   * every CTA re-reads the raw tiles it stored -- L2 resident -- and writes apply_y = act(y*scale+shift) (+ apply_res)).
   * apply_y.ptr == NULL: leave it to sy_bn_act_apply.  The *_group1_offset are element offsets added to the
   * addresses of statistics-group-1 images (DFP fusion, see sy_bn_act_apply). */
  SyTensor apply_y, apply_res;
  int64_t apply_y_group1_offset, apply_res_group1_offset;
  /* ---- debugging only: CTA 0 records (event, clock64) int64 pairs of its three pipeline roles ---- */
  void* debug_timeline;  /* device buffer of 2*debug_timeline_events int64, or NULL */
  int32_t debug_timeline_events;
  int32_t debug_flags;   /* 0 in production; 1 = skip MMAs, 2 = skip TMA loads (pipeline dissection, results invalid) */
  /* ---- validation only: the fp32 accumulators themselves, before the bf16 rounding of the stored result:
   * debug_f32[pixel][Cout] (pixel = flattened (n, oh, ow)), written next to the normal output.  This is where
   * north_star's "within 1e-3 of the reference" is literal (tests/test_gpu_ops.py::test_conv_fp32_accumulators). ---- */
  float* debug_f32;
} SyConvDesc;

/* Rows of the statistics workspace (= SM count: one row per persistent CTA). */
int sy_conv_stat_rows(void);
/* tcgen05 implicit-GEMM kernel (TMA -> smem -> UMMA -> TMEM -> epilogue).  In RAW mode with
 * stat_partials it also accumulates per-channel (sum, sum of squares) of the stored values per
 * statistics group, one partial row per CTA; with bn[] the persistent grid (all CTAs co-resident)
 * ends with a grid barrier and finalizes BatchNorm in parallel: batch statistics -> scale/shift,
 * running statistics (group 0 then group 1, unbiased variance, momentum).  Deterministic.  Do not
 * run two such launches concurrently on one GPU (the barrier needs every SM). */
int sy_conv2d_tc(const SyConvDesc* d, sy_stream_t stream);
/* Host-only query (no launch, no GPU needed): the tiling sy_conv2d_tc chooses for a layer shape -- A-operand mode
 * (0 patch tiles, 1 linear tiles / im2col-mode TMA, 2 halo), tile width BN, tiles and rounds of the persistent grid. */
typedef struct SyConvPlan {
  int32_t mode, bn, m_tiles, n_tiles, rounds, kblocks, patch_h, patch_w;
} SyConvPlan;
int sy_conv2d_plan(int32_t n, int32_t h, int32_t w, int32_t cin, int32_t cout, int32_t kh, int32_t kw, int32_t stride,
                   SyConvPlan* out);
/* plain CUDA-core direct convolution with the same x/y/w/FUSED contract (no statistics):
 * device-side cross-check of the tensor-core kernel. */
int sy_conv2d_simt(const SyConvDesc* d, sy_stream_t stream);

/* Depthwise k x k convolution (groups = channels, k in {1, 3, 5}, stride 1 / 2): the first half of [yolox] DWConv, selected
 * by depthwise=True at exps/model/darknet.py:109, dfp_pafpn.py:31, tal_head.py:53.  Same descriptor and RAW / FUSED contract
 * as sy_conv2d_tc with x.c == y.c and w = bf16 [kh*kw][C]; the statistics fields are ignored (train-mode BatchNorm runs
 * through sy_channel_stats / sy_bn_finalize / sy_bn_act_apply).  Coalesced, vectorised CUDA-core kernel (HBM-bound). */
int sy_dwconv2d(const SyConvDesc* d, sy_stream_t stream);

/* Focus stem, part 1: [yolox] Focus space-to-depth (TL/BL/TR/BR channel order), used at
 * exps/model/darknet.py:115.  x is the NCHW float32 frame-pair batch [b, in_ch, h, w]
 * (exps/model/dfp_pafpn.py:120,145 split it); image n of y takes frame n / b (0 = current,
 * 1 = support; channels 3*frame .. 3*frame+2) of batch element n % b.  y = [frames*b, h/2, w/2, 64]
 * bf16: for each focus pixel its three horizontal taps (x-1, x, x+1; zero outside the image), each
 * 12 focus channels + 4 zero channels, then 16 zero channels (one aligned 128-byte row per pixel).
 * The stem's 3x3 conv then runs on sy_conv2d_tc as a 3x1 conv over 64 channels with weights packed
 * [cout][3][64]. */
int sy_focus_pack(const float* x, int32_t b, int32_t in_ch, int32_t h, int32_t w_px, int32_t frames,
                  SyTensor y, sy_stream_t stream);

/* -------- BatchNorm (train mode) + SiLU (replaces nn.BatchNorm2d + nn.SiLU inside
 * [yolox] BaseConv; eps/momentum from cfgs/s_s50_onex_dfp_tal_flip.py:40-44) ---------- */
/* Per-(row-chunk, channel) partial sums of a stored tensor: partials [P][2][c],
 * P = sy_stats_num_partials(n, h*w) image-major. */
int sy_stats_num_partials(int32_t n, int32_t hw);
int sy_channel_stats(SyTensor x, float* partials, int32_t n_partials, sy_stream_t stream);
/* Reduce partials per group (group 0 = partial rows [0,p_split), group 1 = the rest;
 * the two frames of a pair are normalised separately, exps/model/dfp_pafpn.py:120,145),
 * update running statistics sequentially group 0 then group 1 (unbiased variance,
 * momentum) and emit scale/shift [groups][c] with y = x*scale + shift. */
int sy_bn_finalize(const float* partials, int32_t n_partials, int32_t p_split, int32_t groups,
                   int64_t count_per_group, int32_t c,
                   const float* gamma, const float* beta, float* running_mean, float* running_var,
                   int64_t* num_batches_tracked, float momentum, float eps,
                   float* scale_out, float* shift_out, sy_stream_t stream);
/* y = act(x*scale[g]+shift[g]) (+ res), g = (image >= split_n); single bf16 rounding.
 * y_group1_offset / res_group1_offset: element offsets added to the y / res addresses of the images of
 * statistics group 1 (0 = plain views).  The DFP fusion (exps/model/dfp_pafpn.py:168-170) uses them to
 * write jian(support frame n) into channels [c, 2c) of output image n - split_n, next to jian(current). */
int sy_bn_act_apply(SyTensor x, const float* scale, const float* shift, int32_t split_n, int32_t act,
                    SyTensor res, SyTensor y, int64_t y_group1_offset, int64_t res_group1_offset,
                    sy_stream_t stream);

/* -------- glue ------------------------------------------------------------- */
/* F.interpolate(mode="nearest", size=) of exps/model/dfp_pafpn.py:125,130 into a channel
 * slice; index rule src = min(floor(dst * (float)in/out), in-1) evaluated in float32. */
int sy_upsample_nearest(SyTensor x, SyTensor y, sy_stream_t stream);
/* [yolox] SPPBottleneck pooling (exps/model/darknet.py:156): y5,y9,y13 = stride-1
 * same-padded max pools (k = 5, 9, 13; -inf padding) of x. */
int sy_spp_maxpool(SyTensor x, SyTensor y5, SyTensor y9, SyTensor y13, sy_stream_t stream);
/* strided copy of a view (used for 3-channel duplicates and buffers). */
int sy_copy(SyTensor x, SyTensor y, sy_stream_t stream);

/* -------- head: prediction convs + decode (exps/model/tal_head.py:105-131,167-171,
 * 174,197-199,225-260) ------------------------------------------------------------ */
typedef struct {
  SyTensor cls_feat, reg_feat;   /* [b, h, w, c] */
  const float* w_reg; const float* b_reg;   /* [4][c], [4]  */
  const float* w_obj; const float* b_obj;   /* [1][c], [1]  */
  const float* w_cls; const float* b_cls;   /* [ncls][c], [ncls] */
  int32_t num_classes;
  int32_t stride;          /* 8 / 16 / 32 */
  int32_t anchor_offset;   /* first anchor row of this level in the [b, a_total, 5+ncls] output */
  int32_t a_total;
  int32_t sigmoid;         /* eval: sigmoid on obj / cls */
  int32_t decode;          /* xy = (xy + grid) * stride, wh = exp(wh) * stride */
  float* out;              /* [b, a_total, 5+ncls] float32 */
  float* origin;           /* [b, a_total, 4] raw reg (tal_head.py:185-194) or NULL */
} SyHeadPredDesc;
int sy_head_pred_decode(const SyHeadPredDesc* d, sy_stream_t stream);

/* -------- SimOTA + Trend-Aware loss (exps/model/tal_head.py:262-712) --------- */
typedef struct {
  int32_t b, a_total, max_labels, num_classes;
  int32_t n_levels;
  int32_t level_h[4], level_w[4], level_stride[4];
  const float* outputs;    /* [b, a_total, 5+ncls] decoded, logits for obj/cls */
  const float* origin;     /* [b, a_total, 4] raw reg */
  const float* labels_fut; /* [b, max_labels, 5] (cls, cx, cy, w, h) */
  const float* labels_cur; /* [b, max_labels, 5] */
  float gamma, ignore_thr, ignore_value;
  int32_t use_l1;
  void* workspace; size_t workspace_bytes;   /* >= sy_tal_loss_workspace_bytes() */
  float* loss_out;         /* [6]: total, 5*iou, obj(conf), cls, l1, num_fg / max(num_gt, 1) */
  /* optional dumps for tests / backward: */
  int32_t* fg_out;         /* [b, a_total] 0/1 or NULL */
  int32_t* matched_out;    /* [b, a_total] gt index or -1, or NULL */
  float* pred_iou_out;     /* [b, a_total] or NULL */
} SyTalLossDesc;
size_t sy_tal_loss_workspace_bytes(int32_t b, int32_t a_total, int32_t max_labels, int32_t num_classes);
int sy_tal_loss(const SyTalLossDesc* d, sy_stream_t stream);

/* Weight gradient of a convolution (the cuDNN backward-filter call behind loss.backward(),
 * exps/train_utils/double_trainer.py:114, for every [yolox] BaseConv: exps/model/darknet.py:115-165,
 * dfp_pafpn.py:33-105, tal_head.py:55-104):  dw[co][ci][r][s] (+)= sum_p dy[p][co] * x[p @ (r, s)][ci].
 * x = the layer's input [n, h, w, cin] bf16, dy = gradient w.r.t. the conv output [n, ho, wo, cout] bf16 (same kernel
 * size / stride / padding (k-1)/2 as the forward), dw = fp32 in PyTorch's OIHW parameter layout.  Split-K over the
 * pixels on the tensor cores, then a fixed-order reduction (deterministic).  The workspace holds the fp32 partials. */
typedef struct SyConvWgradDesc {
  SyTensor x;
  SyTensor dy;
  int32_t kh, kw, stride;
  float* dw;               /* [cout, cin, kh, kw] fp32 */
  int32_t accumulate;      /* 0: dw = result, 1: dw += result */
  void* workspace;         /* sy_conv2d_wgrad_workspace_bytes(desc) bytes, 16-byte aligned */
  size_t workspace_bytes;
} SyConvWgradDesc;
size_t sy_conv2d_wgrad_workspace_bytes(const SyConvWgradDesc* d);
int sy_conv2d_wgrad_tc(const SyConvWgradDesc* d, sy_stream_t stream);

/* Backward of BatchNorm(train) + SiLU behind a [yolox] BaseConv (autograd of nn.BatchNorm2d + nn.SiLU under
 * loss.backward(), exps/train_utils/double_trainer.py:114).  raw = the conv output the forward stored, dy = gradient w.r.t.
 * the BaseConv output; scale / shift / mean / invstd = [2 groups][c] as published by the forward (scale = gamma * invstd,
 * shift = beta - mean * scale; images >= split_n form statistics group 1).  Writes draw (bf16, gradient w.r.t. the conv
 * output, input of the conv data / weight gradient kernels), dgamma / dbeta (fp32, (+)=).  partials: sy_bn_act_bwd_rows(n,
 * h*w) rows of 2*c floats; coef: 8*c floats of scratch. */
typedef struct SyBnActBwdDesc {
  SyTensor raw, dy, draw;
  const float* scale; const float* shift; const float* mean; const float* invstd;
  int32_t split_n;
  int32_t act;             /* 1: SiLU, 0: identity */
  float* dgamma; float* dbeta;
  int32_t accumulate;
  float* partials; int32_t n_partials;
  float* coef;
} SyBnActBwdDesc;
int sy_bn_act_bwd_rows(int32_t n, int32_t hw);
int sy_bn_act_backward(const SyBnActBwdDesc* d, sy_stream_t stream);

/* Zero-insertion D[n, 2i, 2j, :] = g[n, i, j, :] (D = [n, H, W, c], H in {2h-1, 2h}): the data gradient of a stride-2
 * 3x3 conv (the first convs of dark2..dark5, bu_conv1/2: exps/model/darknet.py:118-160, dfp_pafpn.py:57-69) is then
 * sy_conv2d_tc on D with the flipped, channel-transposed filter, stride 1. */
int sy_dilate2(SyTensor g, SyTensor D, sy_stream_t stream);

/* Backward of F.interpolate(mode="nearest") (exps/model/dfp_pafpn.py:126,131): dx[n, iy, ix] = sum of dy over the
 * destination pixels whose source index (the forward's fp32 expression) is (iy, ix). */
int sy_upsample_nearest_backward(SyTensor dy, SyTensor dx, sy_stream_t stream);

/* y += x (bf16): gradient accumulation where a tensor feeds several consumers (Bottleneck shortcuts,
 * exps/model/dfp_pafpn.py:168-170 "+ cur", FPN features read by two branches). */
int sy_add(SyTensor x, SyTensor y, sy_stream_t stream);

/* Backward of the three SPP max pools ([yolox] SPPBottleneck, exps/model/darknet.py:156): dx = gradient reaching x
 * through MaxPool2d(5), (9), (13) (stride 1, padding k/2), each window's gradient going to its first maximum in
 * row-major order like PyTorch.  The identity branch of the concat is not included.  Deterministic (gather). */
size_t sy_spp_maxpool_backward_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t c);
int sy_spp_maxpool_backward(SyTensor x, SyTensor d5, SyTensor d9, SyTensor d13, SyTensor dx, void* workspace,
                            size_t workspace_bytes, sy_stream_t stream);

/* Backward of the three 1x1 prediction convs of one head level (exps/model/tal_head.py:101-131, 163-171):
 * grad_raw [b, a_total, 5 + nc] (d loss / d raw head outputs, sy_tal_loss_backward) -> gradients w.r.t. the cls / reg
 * tower outputs (bf16 views), the conv weights ([4][c], [1][c], [nc][c]) and biases (fp32, (+)=).
 * partials: sy_head_pred_bwd_rows(b, h, w) rows of (5 + nc) * (c + 1) floats. */
typedef struct SyHeadPredBwdDesc {
  const float* grad_raw;
  SyTensor cls_feat, reg_feat;         /* the forward's inputs */
  SyTensor d_cls_feat, d_reg_feat;     /* outputs */
  const float* w_reg; const float* w_obj; const float* w_cls;
  int32_t num_classes, a_total, anchor_offset;
  float* dw_reg; float* dw_obj; float* dw_cls; float* db_reg; float* db_obj; float* db_cls;
  int32_t accumulate;
  float* partials; int32_t n_partials;
} SyHeadPredBwdDesc;
int sy_head_pred_bwd_rows(int32_t b, int32_t h, int32_t w);
int sy_head_pred_backward(const SyHeadPredBwdDesc* d, sy_stream_t stream);

/* Backward of the loss: what autograd computes for loss.backward() (exps/train_utils/double_trainer.py:114)
 * through TALHead.get_losses (exps/model/tal_head.py:426-461): the SimOTA assignment, the class targets and the
 * normalised TAL weights are constants (tal_head.py:479 @torch.no_grad, weights detached), so the gradient is
 * per anchor.  Must run after sy_tal_loss on the SAME workspace (assignment and loss sums are read from it).
 * grad_outputs: d/d outputs[b, a, :] (decoded boxes, obj / cls logits); grad_origin: d/d origin_preds;
 * grad_raw: d/d the raw head-conv outputs, i.e. the decode of tal_head.py:237-241 folded in and the L1 path added
 * (what the prediction convs' backward consumes).  Any of the three may be NULL. */
typedef struct SyTalLossBwdDesc {
  const float* outputs;    /* [b, a_total, 5 + num_classes] as given to sy_tal_loss */
  const float* origin;     /* [b, a_total, 4] or NULL when !use_l1 */
  const float* labels_fut; /* [b, max_labels, 5] */
  int32_t b, a_total, max_labels, num_classes, n_levels;
  int32_t level_h[4], level_w[4], level_stride[4];
  float gamma;
  int32_t use_l1;
  void* workspace;         /* the workspace sy_tal_loss ran on */
  size_t workspace_bytes;
  float grad_scale;        /* d objective / d total_loss (1, or the AMP loss scale) */
  float* grad_outputs;     /* [b, a_total, 5 + num_classes] or NULL */
  float* grad_origin;      /* [b, a_total, 4] or NULL */
  float* grad_raw;         /* [b, a_total, 5 + num_classes] or NULL */
} SyTalLossBwdDesc;
int sy_tal_loss_backward(const SyTalLossBwdDesc* d, sy_stream_t stream);

/* Detection post-processing = [yolox 0.3.0] yolox.utils.postprocess as called by the evaluators and the streaming
 * driver (exps/evaluators/onex_stream_evaluator.py:148, sAP/streamyolo/streamyolo_det.py:62-83): cxcywh -> xyxy,
 * class_conf / class_pred = max over the class scores, keep obj * class_conf >= conf_thre, class-aware greedy NMS
 * (torchvision.ops.batched_nms semantics), rows [x1, y1, x2, y2, obj, class_conf, class_pred] in decreasing score order.
 * pred = the eval-mode head output [b, a_total, 5 + num_classes] (decoded boxes, sigmoid scores).  One CTA per image;
 * a_total <= 16384.  det_out [b, max_det, 7] (rows past count_out[i] are not written), count_out [b]. */
typedef struct SyNmsDesc {
  const float* pred;
  int32_t b, a_total, num_classes, max_det;
  float conf_thre, nms_thre;
  int32_t class_agnostic;
  void* workspace;         /* sy_postprocess_nms_workspace_bytes(b, a_total) bytes, 16-byte aligned */
  size_t workspace_bytes;
  float* det_out;
  int32_t* count_out;
} SyNmsDesc;
size_t sy_postprocess_nms_workspace_bytes(int32_t b, int32_t a_total);
int sy_postprocess_nms(const SyNmsDesc* d, sy_stream_t stream);

/* -------- training step glue (streamyolo_b200/csrc/train_glue.cu) ---------------------------------------------- */
/* fp32 OIHW conv parameter -> bf16 GEMM operand, on the device (one launch per parameter per optimiser step):
 *   mode 0  out[o][r*kw+s][i] = w[o][i][r][s]                          forward B operand of sy_conv2d_tc
 *   mode 1  out[i][taps-1-(r*kw+s)][co_offset + o] = w[o][i][r][s]     data-gradient operand (flipped taps, transposed
 *           channels; rows of out_pitch elements so that the conv1 | conv2 pair of a CSPLayer packs into one operand)
 *   mode 2  out[o][r][s*16 + i] = w[o][i][r][s], 64 columns per (o, r) Focus stem (see sy_focus_pack)
 * Replaces the weight.to(bf16).permute chain a PyTorch host would run after every optimizer.step()
 * (exps/train_utils/double_trainer.py:119-121). */
int sy_pack_conv_weight(const float* w, int32_t cout, int32_t cin, int32_t kh, int32_t kw, int32_t mode, void* out,
                        int64_t out_pitch, int32_t co_offset, sy_stream_t stream);

/* The same for MANY parameters in one launch: items (a DEVICE array) lists (parameter, layout) pairs with the arguments of
 * sy_pack_conv_weight.  The launch works in TILES (64 output x 32 input channels of one item; 64 output channels of a stem
 * item): `begin` = index of the item's first tile (prefix sum of sy_pack_item_tiles over the items), `total` = their sum.
 * The trainer re-packs every conv operand of the model (forward and data-gradient layouts) with it after each optimiser step. */
typedef struct SyPackItem {
  const float* w;
  void* out;
  int32_t cout, cin, kh, taps, mode, co_offset;
  int64_t out_pitch;
  int64_t begin;
} SyPackItem;
int64_t sy_pack_item_tiles(int32_t cout, int32_t cin, int32_t mode);
int sy_pack_conv_weights_batch(const SyPackItem* items_dev, int32_t n_items, int64_t total, sy_stream_t stream);

/* The optimiser step of the reference trainer as one launch over flat fp32 state:
 * GradScaler.unscale_ + [yolox] Exp.get_optimizer's SGD(momentum, nesterov) with weight decay on the conv / linear weights
 * only + [yolox] ModelEMA.update (exps/train_utils/double_trainer.py:113-123, 173-175).  Elements [0, n_param) are
 * parameters in the order (BatchNorm weights, biases | decayed weights from decay_begin); [n_param, n_total) are the
 * floating-point buffers (BatchNorm running statistics) that only the EMA follows.  Step-by-step the same roundings as
 * torch.optim.SGD / ModelEMA in fp32.  found_inf (device float, may be NULL): non-zero skips the whole update. */
typedef struct SySgdEmaDesc {
  float* param;              /* [n_total] model state (parameters then float buffers) */
  const float* grad;         /* [n_param] (the all-reduced flat gradient buffer) */
  float* momentum_buf;       /* [n_param] */
  float* ema;                /* [n_total] or NULL (no EMA) */
  int64_t n_param, n_total, decay_begin;
  float lr, momentum, weight_decay, inv_scale;
  int32_t nesterov;
  float ema_decay, ema_one_minus_decay;
  const float* found_inf;
  /* optional device array [lr, momentum, weight_decay, inv_scale, ema_decay, 1 - ema_decay]: when non-NULL it REPLACES the
   * six scalars above, so that a step captured in a CUDA graph follows the LR schedule / EMA ramp / loss scale of the
   * iteration it is replayed in (the host rewrites the 24 bytes before each replay). */
  const float* hyper;
} SySgdEmaDesc;
int sy_sgd_nesterov_ema_step(const SySgdEmaDesc* d, sy_stream_t stream);

/* Input pipeline on the device: Exp.preprocess (cfgs/s_s50_onex_dfp_tal_flip.py:160-171) =
 * F.interpolate(inputs, size=tsize, mode="bilinear", align_corners=False) on the NCHW fp32 frame-pair batch
 * (x: [nc = B*6][hi][wi] -> y: [nc][ho][wo]) and the label rescale targets[..., 1::2] *= sx, [..., 2::2] *= sy
 * (labels: rows x cols floats, column 0 = class, in place). */
int sy_resize_bilinear(const float* x, int32_t nc, int32_t hi, int32_t wi, float* y, int32_t ho, int32_t wo,
                       sy_stream_t stream);
int sy_scale_labels(float* labels, int64_t rows, int32_t cols, float sx, float sy, sy_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* STREAMYOLO_SM100_H_ */
