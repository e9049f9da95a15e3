/*
 * din_sm100.h — C ABI of libdin_sm100.so: the DIN stage-2 forward hot path on NVIDIA B200 (sm_100a).
 *
 * The reference (JacobYuan7/DIN-Group-Activity-Recognition-Benchmark) is 100 % Python; its only
 * native boundary is the external RoIAlign torch extension
 *   CropAndResizeFunction.apply(featuremap, boxes, box_ind, crop_h, crop_w, extrapolation)
 * (imported at infer_model.py:3, called at infer_model.py:178-180).  Every other op on the path is a
 * torch library call.  Each entry point below names the reference call site(s) it replaces.
 *
 * Conventions (all entry points)
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - the CALLER owns every buffer (inputs, outputs, weights, workspaces); the library never
 *     allocates or frees device memory and keeps no pointer after the call returns;
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*), no internal
 *     synchronisation, re-entrant across threads that use distinct streams; device = current device;
 *   - return value: DIN_OK (0) or a negative DIN_ERR_* code; the message is available from
 *     din_last_error_string() (thread-local).  Shape / alignment violations are reported before any
 *     launch.  The library never throws and never exits;
 *   - activations between backbone layers are NHWC fp16; accumulation is fp32; the person-level head
 *     (LayerNorms, Dynamic Relation / Dynamic Walk, read-out) is fp32 end to end.
 */
#ifndef DIN_SM100_H_
#define DIN_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DIN_API __attribute__((visibility("default")))
#else
#define DIN_API
#endif

#define DIN_OK 0
#define DIN_ERR_INVALID_ARG (-1)
#define DIN_ERR_CUDA (-2)
#define DIN_ERR_UNSUPPORTED (-3)

/* ---- library ------------------------------------------------------------------------------- */

/* ABI version (bumped on any signature change). */
DIN_API int din_abi_version(void);
/* Last error message of the calling thread ("" if none). Never NULL. */
DIN_API const char* din_last_error_string(void);
/* Number of SMs of the current device (148 on B200), or a negative error. */
DIN_API int din_device_sm_count(void);
/* Encoded CUtensorMaps are cached per (base pointer, extents, strides, box, swizzle): SURVEY.md section 8b.  Returns the
 * number of cached maps and, through the optional out-pointers, the hit / miss counts since load (diagnostics). */
DIN_API int din_tmap_cache_stats(unsigned long long* hits, unsigned long long* misses);
/* Diagnostics of din_conv3x3_stem_pair_nhwc_f16 with DIN_FUSED_DEBUG=1 in the environment: every bounded mbarrier wait
 * of that kernel records (wait id << 24 | block << 12 | thread) in a host-mapped word before it traps, so that a
 * mis-programmed pipeline names the wait that starved even though the context is lost.  Word 0 = first timeout. */
DIN_API unsigned int din_debug_word(int i);

/* ---- backbone ------------------------------------------------------------------------------ */

/*
 * Stem convolution fused with prep_images.
 * Replaces: utils.py:8-19 prep_images ((x/255 - 0.5) * 2) followed by the first backbone conv
 *   (VGG16 features.0 3x3 s1 p1 + ReLU, backbone.py:88-99; ResNet-18 conv1 7x7 s2 p3 + folded BN
 *   + ReLU, backbone.py:115-132; Inception-v3 Conv2d_1a_3x3 s2 p0 + folded BN + ReLU, backbone.py:44).
 * x   : [n, 3, h, w] fp32 NCHW, raw pixel values 0..255 (volleyball.py:243,270)
 * w   : [c_out, 3, kh, kw] fp32 (OIHW, BN already folded by the caller), bias: [c_out] fp32
 * y   : [n, oh, ow, c_out] fp16 NHWC, oh = (h + 2*pad - kh)/stride + 1 (same for ow)
 * c_out must be a multiple of 8 and <= 64. kh*kw <= 49.  prep != 0 applies prep_images to x first.
 */
DIN_API int din_stem_conv_nchw_f32(const float* x, const float* w, const float* bias, void* y, int n, int h,
                           int w_in, int c_out, int kh, int kw, int stride, int pad, int relu, int prep,
                           void* stream);

/*
 * The same stem with uint8 NHWC ingest: the decoded frame exactly as the reference's loader holds it before
 * `img.transpose(2,0,1)` and `.float()` (volleyball.py:239-243,270; collective.py:189-193) -- 4x fewer bytes
 * over PCIe and HBM than the fp32 CHW tensor; uint8 -> fp32 is exact, so the result is bit-identical to
 * din_stem_conv_nchw_f32 on the same pixel values.  Replaces the loader's transpose + float conversion in
 * addition to what din_stem_conv_nchw_f32 replaces.
 * x : [n, h, w, 3] uint8.  Only the three backbone stems are instantiated (c_out/kh/kw/stride = 64/3/3/1,
 * 64/7/7/2, 32/3/3/2); anything else returns DIN_ERR_UNSUPPORTED.
 */
DIN_API int din_stem_conv_nhwc_u8(const uint8_t* x, const float* w, const float* bias, void* y, int n, int h,
                          int w_in, int c_out, int kh, int kw, int stride, int pad, int relu, int prep,
                          void* stream);

/*
 * Implicit-GEMM convolution on the tcgen05 tensor cores (TMA-staged fp16 operands, fp32 TMEM
 * accumulators), fused bias (+ residual) (+ ReLU) epilogue.  Also serves as the dense GEMM of the
 * path (kh = kw = 1, n = h = 1, w = rows).
 * Replaces: every Conv2d(+BN)(+ReLU) of the truncated torchvision backbones (backbone.py:10-132) and
 *   nn.Linear fc_emb_1 (infer_model.py:50,184).
 */
typedef struct DinConvDesc {
  int32_t n, h, w;          /* input spatial extent (NHWC) */
  int32_t c_in;             /* input channels used by this conv; multiple of 8 (K is processed in blocks of 64:
                               a partial last block is zero-filled by the TMA unit) */
  int32_t x_c_stride;       /* channel count of the buffer x lives in (>= c_in; x may be a channel slice) */
  int32_t c_out;            /* output channels; multiple of 8 */
  int32_t y_c_stride;       /* channel count of the buffer y lives in (>= c_out; concat-by-offset writes) */
  int32_t kh, kw;           /* filter taps */
  int32_t stride;           /* 1 or 2 */
  int32_t pad_h, pad_w;     /* zero padding */
  int32_t relu;             /* fuse ReLU */
  int32_t out_f32;          /* 0: y is fp16, 1: y is fp32 */
  int32_t pool2;            /* fuse MaxPool2d(2,2) after (bias, ReLU): y is [n, oh/2, ow/2, ...] (fp16, no residual) */
  int32_t w_split;          /* 0/1: w_packed holds one fp16 part; 2: hi and lo fp16 parts ([c_out][2][kh][kw][c_in_padded]),
                               both accumulated in the same TMEM tile (weights effectively exact, 2x tensor work) */
} DinConvDesc;

/*
 * x        : fp16, element (img, yy, xx, c) at x[((img*h + yy)*w + xx)*x_c_stride + c]; 16-byte aligned
 * w_packed : fp16 [c_out][kh][kw][c_in_padded], c_in_padded = c_in rounded up to 64, zero columns beyond c_in
 *            (K-major rows; see din_pack_conv_weight_f16); 16-byte aligned
 * bias     : fp32 [c_out] or NULL
 * residual : fp16, same indexing as y with y_c_stride, added before ReLU; or NULL
 * y        : fp16 or fp32 (out_f32), element (img, oy, ox, co) at y[((img*oh + oy)*ow + ox)*y_c_stride + co]
 *            (with pool2: oh, ow are the pooled extents floor(oh/2), floor(ow/2), as nn.MaxPool2d(2,2))
 */
DIN_API int din_conv2d_nhwc_f16(const DinConvDesc* desc, const void* x, const void* w_packed, const float* bias,
                        const void* residual, void* y, void* stream);

/* Several convolutions that read the SAME input (the 1x1 branch heads of an Inception block: branch1x1, branch5x5_1 /
 * branch7x7_1, branch*dbl_1 and branch_pool's 1x1, torchvision Inception{A,C}.forward under backbone.py:62-80) as ONE
 * implicit GEMM whose weight rows are the branches' filters stacked: the input is read once, N is wide enough for the
 * 256-column MMA, one launch instead of four.  Output columns [0, split_col) are written to y (channel stride
 * desc.y_c_stride: the block's concat buffer), columns [split_col, c_out) to y2 (channel stride y2_c_stride: scratch that
 * the branches' next convolutions read as channel slices).  Columns [norelu_lo, norelu_hi) skip desc.relu: the pool
 * branch's 1x1 runs BEFORE its 3x3 average pool here (both are linear and zero padding with count_include_pad commutes
 * with a bias-free 1x1), so the pool streams c_out instead of c_in channels; bias + ReLU follow the pool
 * (din_avgpool3_bias_relu_nhwc_f16).  split_col, norelu_lo, norelu_hi: multiples of 32.  No pool2 / out_f32 / residual. */
typedef struct DinConvBranchOut {
  int32_t split_col;
  int32_t y2_c_stride;
  int32_t norelu_lo, norelu_hi;
} DinConvBranchOut;
DIN_API int din_conv2d_branches_nhwc_f16(const DinConvDesc* desc, const DinConvBranchOut* branches, const void* x,
                                         const void* w_packed, const float* bias, void* y, void* y2, void* stream);

/* VGG-16's first two layers in ONE launch (inference): conv1_1 = 3x3 pad 1 over the 3-channel image (prep_images fused,
 * bias, ReLU) computed tile by tile INSIDE the CTA-pair kernel of conv1_2 = 3x3 pad 1, 64 -> 64 channels, bias, ReLU,
 * optional fused 2x2 max-pool: conv1_1's 64-channel output (118 MB per 720p frame) never reaches HBM
 * (csrc/conv_tcgen05.cu: conv1_fused_2cta_kernel).  Same values as din_stem_conv_* followed by din_conv2d_nhwc_f16.
 * Replaces: prep_images (utils.py:8-19) + vgg16.features[0:4] (+ features[4] with pool2) of backbone.py:88-99.
 * x        : fp32 NCHW [n,3,h,w] raw 0..255, w % 4 == 0   |   (x_is_u8) uint8 NHWC [n,h,w,3], w % 16 == 0
 *            (TMA row pitch; other widths: use the two separate calls)
 * w1, b1   : conv1_1 weight fp32 OIHW [64,3,3,3] and bias [64] (or NULL)
 * w2_packed: conv1_2 weight packed by din_pack_conv_weight_f16 (fp16 [64][3][3][64]); b2: fp32 [64] or NULL
 * y        : fp16 NHWC [n, h, w, y_c_stride] (pool2: [n, h/2, w/2, y_c_stride]), channels [0, 64) written */
DIN_API int din_conv3x3_stem_pair_nhwc_f16(const void* x, int x_is_u8, const float* w1, const float* b1,
                                           const void* w2_packed, const float* b2, void* y, int n, int h, int w,
                                           int y_c_stride, int relu2, int pool2, int prep, void* stream);

/* Data gradient of a convolution fused with the backward of the ReLU that FOLLOWS the layer below:
 *   dx = conv(dz, w_packed) * [y_saved > 0]
 * (w_packed = the data-gradient filter, DinPackJob.transposed = 1; y_saved = that layer's saved ReLU output, fp16, indexed
 * as dx with y_c_stride).  Same kernels as din_conv2d_nhwc_f16 -- the mask replaces the residual add in the epilogue -- so
 * the gradient is written once instead of written, re-read, masked and written again.  desc.relu / pool2 / out_f32 = 0.
 * Replaces: cudnn_convolution_backward_input + threshold_backward under train_net_dynamic.py:186 loss.backward(). */
DIN_API int din_conv2d_relu_bwd_nhwc_f16(const DinConvDesc* desc, const void* dz, const void* w_packed,
                                 const void* y_saved, void* dx, void* stream);

/* OIHW fp32 [c_out, c_in, kh, kw] (optionally scaled per output channel by `scale`, for BN folding;
 * NULL = 1) -> fp16 [c_out][kh][kw][c_in_padded], zero-filled for c_in <= c < c_in_padded.
 * Rounding is error-feedback along each output row (every weight within 1 ulp(fp16) of its fp32 value, the
 * row's summed rounding error ~1 ulp): weight rounding errors do not average out over pixels, see
 * csrc/conv_tcgen05.cu.  split = 2 instead writes [c_out][2][kh][kw][c_in_padded]: hi = RN(w), lo = RN(w - hi)
 * (for DinConvDesc.w_split = 2).  Device-to-device, one-time (not on the per-step path). */
DIN_API int din_pack_conv_weight_f16(const float* w_oihw, const float* scale, void* w_packed, int c_out, int c_in,
                             int c_in_padded, int kh, int kw, int split, void* stream);

/* Several weights packed by ONE launch (the packing walk is latency-bound on a few CTAs, so a backbone's ~25 filters are
 * packed side by side: a training step re-packs every weight after the optimizer step).
 * transposed = 0: as din_pack_conv_weight_f16 (w is [rows][cols][kh][kw], scale indexes rows).
 * transposed = 1: the DATA-GRADIENT filter of the convolution whose forward weight w is [cols][rows][kh][kw]
 *                 (torch: w.permute(1,0,2,3).flip(2,3), times scale[col]): packed row = input channel, column = output
 *                 channel, taps rotated by 180 degrees -- what din_conv2d_nhwc_f16 needs to compute dX = conv(dZ, .).
 * Replaces: the per-layer weight preparation that torch.nn.Conv2d's cuDNN backward does internally
 * (backbone/backbone.py:88-132 layers under train_net_dynamic.py:186 loss.backward()). */
typedef struct DinPackJob {
  const float* w;           /* fp32 source weight (device) */
  const float* scale;       /* fp32 per-output-channel scale (BatchNorm folding) or NULL */
  void* out;                /* fp16 [rows][split][kh][kw][cols_padded] (device, 16-byte aligned) */
  int32_t rows, cols;       /* packed rows (GEMM M) and columns per tap */
  int32_t cols_padded;      /* cols rounded up to a multiple of 64, zero-filled */
  int32_t kh, kw;
  int32_t split;            /* 1, or 2 (hi / lo parts) */
  int32_t transposed;       /* see above */
  int32_t reserved;
} DinPackJob;
DIN_API int din_pack_conv_weights_f16(const DinPackJob* jobs, int n_jobs, void* stream);

/* ResNet-18's stem in ONE launch (inference): conv1 7x7 stride 2 pad 3 over the 3-channel image (prep_images fused, eval
 * BatchNorm folded into w / bias by the caller) + ReLU + MaxPool2d(3, 2, 1) (resnet18.conv1 / bn1 / relu / maxpool,
 * backbone.py:115-132).  The stand-alone stem is bound by its output writes and the pool discards three quarters of them:
 * here the 64-channel half-resolution map never reaches HBM (a CTA keeps the last three convolution rows of its strip in
 * shared memory and emits a pooled row for every second one).  Bit-identical to din_stem_conv_* followed by
 * din_maxpool2d_nhwc_f16(k = 3, stride = 2, pad = 1).
 * x: fp32 NCHW [n,3,h,w] raw 0..255 (w % 4 == 0) | (x_is_u8) uint8 NHWC [n,h,w,3] (w % 16 == 0); w fp32 OIHW [64,3,7,7];
 * bias fp32 [64] or NULL; y fp16 NHWC [n, ph, pw, 64], ph = ((h + 6 - 7) / 2 + 1 - 1) / 2 + 1 (pw alike).
 * DIN_ERR_UNSUPPORTED for other widths / unaligned x: use the two separate calls. */
DIN_API int din_stem7x7_pool_nhwc_f16(const void* x, int x_is_u8, const float* w, const float* bias, void* y, int n, int h,
                                      int w_in, int prep, void* stream);

/*
 * Max / average pooling, NHWC fp16, over channels [0, c) of buffers with x_c_stride / y_c_stride channels
 * (so a pool can read a channel slice and write straight into a concat buffer).
 * Replaces nn.MaxPool2d(2,2) in vgg16.features (normally fused into the conv epilogue, see pool2),
 * resnet18.maxpool (3,2,1), F.max_pool2d(3,2) at backbone.py:50,56 and in Inception's Mixed_6a, and
 * F.avg_pool2d(3,1,1) of the Inception blocks' pool branches (count_include_pad=True: divides by k*k).
 * Max pooling: padding elements never win, as in torch.
 */
DIN_API int din_maxpool2d_nhwc_f16(const void* x, void* y, int n, int h, int w, int c, int x_c_stride,
                                   int y_c_stride, int k, int stride, int pad, void* stream);
DIN_API int din_avgpool2d_nhwc_f16(const void* x, void* y, int n, int h, int w, int c, int x_c_stride,
                                   int y_c_stride, int k, int stride, int pad, void* stream);

/* y = relu?(avg_pool3x3(x, stride 1, pad 1, count_include_pad) + bias[c]): the tail of an Inception pool branch whose 1x1
 * convolution ran before the pool (din_conv2d_branches_nhwc_f16).  Channel-strided like the pools above; bias fp32 [c]. */
DIN_API int din_avgpool3_bias_relu_nhwc_f16(const void* x, void* y, const float* bias, int n, int h, int w, int c,
                                            int x_c_stride, int y_c_stride, int relu, void* stream);

/*
 * Bilinear resize with align_corners=True, NHWC fp16, [n,h,w,c] -> [n,oh,ow,c] (channel-strided buffers).
 * Replaces F.interpolate(features, size=(OH,OW), mode='bilinear', align_corners=True) (infer_model.py:169)
 * and, by writing at a channel offset of the multiscale map, the torch.cat at infer_model.py:172.
 */
DIN_API int din_upsample_bilinear_nhwc_f16(const void* x, void* y, int n, int h, int w, int c, int x_c_stride,
                                           int y_c_stride, int oh, int ow, void* stream);

/* ---- person-level head --------------------------------------------------------------------- */

/*
 * RoIAlign = TF crop_and_resize with longcw/RoIAlign.pytorch's box transform (transform_fpcoor=True,
 * extrapolation_value=0): one bilinear sample per output bin, zero where the sample point leaves
 * [0,h-1] x [0,w-1].
 * Replaces: roi_align.roi_align.RoIAlign(crop_h, crop_w)(featuremap, boxes, box_ind), the reference's one
 *   native extension (infer_model.py:3,48,178-181; Dockerfile:6-8).
 * fm      : fp16 NHWC [n_img, h, w, fm_c_stride], channels [0, d) are read
 * boxes   : fp32 [m, 4] = (x1, y1, x2, y2) in feature-map units (volleyball.py:249-251)
 * box_ind : int32 [m] frame index of each box (boxes with an index outside [0, n_img) produce zeros)
 * out     : fp16 [m][crop_h*crop_w][d]  -- bin-major, channel-minor: the K order of the packed fc_emb_1
 *           weight, so `out` is the A operand of the embedding GEMM without a transpose
 *           (the reference flattens (d, ky, kx), infer_model.py:181; the caller permutes fc_emb_1's
 *           columns once at load time instead).
 */
DIN_API int din_roi_align_nhwc_f16(const void* fm, const float* boxes, const int32_t* box_ind, void* out,
                                   int n_img, int h, int w, int d, int fm_c_stride, int m, int crop_h,
                                   int crop_w, void* stream);
/* Same sampling, un-rounded fp32 output [m][crop_h*crop_w][d]: the A operand of the fp32 embedding path
 * (din_linear_f32) that the plan takes when there are fewer actor rows than half an MMA tile -- a launch that small is
 * latency-bound, so skipping the fp16 rounding of crops and fc_emb_1 weights costs nothing there. */
DIN_API int din_roi_align_nhwc_f16_f32out(const void* fm, const float* boxes, const int32_t* box_ind, float* out,
                                          int n_img, int h, int w, int d, int fm_c_stride, int m, int crop_h,
                                          int crop_w, void* stream);

/*
 * LayerNorm over strided groups with optional pre-add, ReLU, post-add:
 *     y = [relu]( LN(x (+ pre)) * gamma + beta ) (+ post)           (fp32, biased variance, eps inside sqrt)
 * group g in [0, n_outer*n_inner): base = (g / n_inner)*outer_stride + (g % n_inner)*inner_stride;
 * element (r, c), r < rows, c < cols, lives at base + r*row_stride + c; gamma/beta at r*cols + c.
 * n_valid (int32 [n_outer], may be NULL): groups with (g % n_inner) >= n_valid[g / n_inner] are skipped.
 * Replaces: nl_emb_1 (+ReLU) infer_model.py:185-186; point_ln (+ReLU) :191-192; dpi_nl fusion :203-216;
 *   hier_LN dynamic_infer_module.py:493-494; Collective LayerNorm([T, C]) infer_model.py:1298-1301.
 */
DIN_API int din_group_layernorm_f32(const float* x, const float* pre, const float* post, const float* gamma,
                                    const float* beta, float* y, int n_outer, int n_inner,
                                    long long outer_stride, long long inner_stride, int rows,
                                    long long row_stride, int cols, float eps, int relu,
                                    const int32_t* n_valid, void* stream);

/*
 * y[m, n] (+)= x[m, k] . w[n, k]^T (+ bias[n]) (ReLU), fp32 (k a multiple of 16).  accumulate != 0 adds
 * to the existing y (the sum over parallel DIMs, dynamic_infer_module.py:441); ReLU applies to the new term.
 * Replaces: point_conv 1x1 (infer_model.py:189-190), hidden_weight (dynamic_infer_module.py:149).
 */
DIN_API int din_linear_f32(const float* x, const float* w, const float* bias, float* y, int m, int n, int k,
                           int relu, int accumulate, void* stream);

/*
 * Dynamic Relation + Dynamic Walk for one sampling ratio, fused (affinity convs -> softmax over the
 * kt x kn neighbourhood -> dynamic-walk bilinear sampling -> relation-weighted aggregation).
 * Replaces: Dynamic_Person_Inference.dynamic_infer_ratio and helpers
 *   (infer_module/dynamic_infer_module.py:184-282, 344-404) plus the ratio mean / beta sum (:142-147).
 * x      : fp32 [b, t, n, c]
 * w_tap  : fp32 [kt*kn][n_out][c], n_out = 2*kt*kn (+ kt*kn if scale_factor): p_conv rows first (T-axis
 *          offsets, then N-axis offsets), then scale_conv rows; tap-major repack of the OIHW weights
 * b_cat  : fp32 [n_out]
 * y      : fp32 [b, t, n, c];  y = coef * out   or, if accumulate, y += coef * out
 * coef   : *coef_ptr if coef_ptr != NULL (e.g. &beta[r], device memory) else coef_scalar (1/len(ratios))
 * n_valid: int32 [b] real actors per clip (Collective: infer_model.py:1289), NULL = n everywhere; actors
 *          >= n_valid[b] are neither read nor written and the zero padding starts right after them.
 * scale_factor == 0 -> plain mean over the taps (:280).  kt, kn odd, kt*kn <= 9, n <= 16.
 */
DIN_API int din_dynamic_infer_f32(const float* x, const float* w_tap, const float* b_cat, float* y, int b, int t,
                                  int n, int c, int kt, int kn, int ratio, int scale_factor,
                                  const float* coef_ptr, float coef_scalar, int accumulate,
                                  const int32_t* n_valid, void* stream);

/*
 * Group read-out: max over actors -> fc_activities -> mean over frames.
 * Replaces: infer_model.py:224-232 (Volleyball) and :1311-1313 (Collective, with n_valid).
 * s [b, t, n, c] fp32, w [a, c], bias [a] -> logits [b, a].   a <= 64.
 * n_valid[i] < 1 is a device-side trap (message on stdout): the reference's torch.max over an empty actor axis raises.
 */
DIN_API int din_readout_f32(const float* s, const float* w, const float* bias, float* logits, int b, int t,
                            int n, int c, int a, const int32_t* n_valid, void* stream);

/* ---- after the path: loss / metrics on the device, stage-1 helper ------- */

/*
 * Cross-entropy (mean reduction, optional class weights), argmax, correct count, confusion matrix and
 * epoch meters in ONE launch, with d(loss)/d(logits) for the backward pass; nothing is copied to the host.
 * Replaces: F.cross_entropy / torch.argmax / torch.eq().sum() / .item() / ConfusionMeter.add /
 *   AverageMeter.update per step (train_net_dynamic.py:191-199, 201-210, 217, 258-292; utils.py:193-264).
 * logits       : fp32 [b, a];  labels: int64 [b] (labels outside [0, a) are ignored, like ignore_index)
 * class_weight : fp32 [a] or NULL (cfg.actions_weights, train_net_dynamic.py:203-204)
 * loss         : fp32 [1] = loss_scale * sum_i w[y_i] nll_i / sum_i w[y_i]             (or NULL)
 * correct      : int32 [1] number of rows whose first arg-max equals the label (overwritten; or NULL)
 * conf         : int32 [a, a], conf[target][predicted] += 1 (ACCUMULATES across calls; or NULL)
 * meters       : fp64 [4] ACCUMULATING: sum(loss * b), sum(b), sum(correct), steps        (or NULL)
 * dlogits      : fp32 [b, a] = loss_scale * w[y_i] / sum w * (softmax(logits_i) - onehot(y_i))   (or NULL)
 * a <= 64.  labels[i] == -100 (torch's ignore_index) contributes nothing; any other label outside [0, a) is a device-side
 * trap (message on stdout), as torch's device-side assert.
 */
DIN_API int din_ce_metrics_f32(const float* logits, const int64_t* labels, const float* class_weight,
                               float loss_scale, float* loss, int32_t* correct, int32_t* conf, double* meters,
                               float* dlogits, int b, int a, void* stream);

/*
 * y[o, i] = mean over a of x[o, a, i]   (fp32; x is [outer, len, inner]).
 * Replaces: actions_scores.reshape(B,T,N,-1).mean(dim=1) of the stage-1 base model (base_model.py:138-139).
 */
DIN_API int din_mean_axis_f32(const float* x, float* y, int outer, int len, int inner, void* stream);

/* ---- backward of the person-level head (first slice) ------------------------ *
 * What `total_loss.backward()` (train_net_dynamic.py:220-224) computes for everything after the feature map,
 * i.e. the stage-2 training step with the backbone frozen (config.py:39 train_backbone = False).  All fp32.
 * Parameter gradients are OVERWRITTEN (fixed summation order, deterministic); activation gradients follow the
 * per-call `accumulate` / "+=" rules stated below.                                                            */

/*
 * General strided GEMM: c[m, n] (+)= alpha * sum_k A(m, k) * B(k, n),  A(m,k) = a[m*a_stride_m + k*a_stride_k]
 * (fp32), B(k,n) = b[k*b_stride_k + n*b_stride_n] (fp32, or fp16 when b_is_f16 -- the RoIAlign crops).
 * Replaces autograd's mm / addmm backward of nn.Linear / 1x1 conv: dX = dY.W, dW = dY^T.X
 *   (fc_emb_1 infer_model.py:184, point_conv :189-190, hidden_weight dynamic_infer_module.py:149).
 */
DIN_API int din_gemm_f32(const float* a, long long a_stride_m, long long a_stride_k, const void* b, int b_is_f16,
                         long long b_stride_k, long long b_stride_n, float* c, long long ldc, int m, int n, int k,
                         float alpha, int accumulate, void* stream);

/* y[n] = sum_m x[m*ld + n]  (bias gradients of the linears above). */
DIN_API int din_colsum_f32(const float* x, float* y, int m, int n, long long ld, void* stream);

/* y = x * mask * scale, mask = 0/1 bytes (NULL = ones).  Dropout forward and backward with a caller-supplied
 * mask (nn.Dropout / F.dropout: infer_model.py:209,216; dynamic_infer_module.py:495). */
DIN_API int din_scale_mask_f32(const float* x, const uint8_t* mask, float scale, float* y, long long count,
                               void* stream);

/* dz = dy * [y > 0] (fp32): backward of F.relu from its saved output (base_model.py:120, the stage-1 embedding). */
DIN_API int din_relu_bwd_f32(const float* y, const float* dy, float* dz, long long count, void* stream);

/*
 * Backward of din_readout_f32 (max over actors -> fc_activities -> mean over frames; infer_model.py:224-232).
 * s [b,t,n,c], w [a,c], dlogits [b,a]  ->  ds [b,t,n,c] (gradient at each channel's FIRST arg-max actor, zero
 * elsewhere), dw [a,c], dbias [a].  pooled_ws: workspace [b,t,c].
 */
DIN_API int din_readout_bwd_f32(const float* s, const float* w, const float* dlogits, float* ds, float* pooled_ws,
                                float* dw, float* dbias, int b, int t, int n, int c, int a, const int32_t* n_valid,
                                void* stream);

/*
 * Backward of din_group_layernorm_f32 (same group geometry; `post` needs no kernel: its gradient is dy).
 * dx (+)= d(loss)/d(x) (which is also the gradient of `pre`); dgamma / dbeta [rows*cols] overwritten, or both
 * NULL.  stats_ws: workspace [2 * n_outer * n_inner].
 * Replaces autograd through nn.LayerNorm + F.relu (+ residual adds): infer_model.py:185-186, 191-192, 203-216,
 *   1298-1301; dynamic_infer_module.py:493-494.
 */
DIN_API int din_group_layernorm_bwd_f32(const float* x, const float* pre, const float* gamma, const float* beta,
                                        const float* dy, float* dx, float* dgamma, float* dbeta, float* stats_ws,
                                        int n_outer, int n_inner, long long outer_stride, long long inner_stride,
                                        int rows, long long row_stride, int cols, float eps, int relu,
                                        int accumulate_dx, const int32_t* n_valid, void* stream);

/*
 * Backward of din_dynamic_infer_f32 for one sampling ratio (dynamic_infer_module.py:184-282 under autograd):
 * given dy = d(loss)/d(y) where y = coef * DIN_ratio(x),
 *   dx     += gradient through the four corner gathers (fp32 atomics) and through p_conv / scale_conv
 *             (the caller zero-fills dx or passes the gradient x already received from other consumers);
 *   dw_tap  = [kt*kn][n_out][c], db_cat = [n_out]: gradients in the packed layout of the forward;
 *   dcoef   = [1] d(loss)/d(coef) (the gradient of beta[r] when beta_factor), or NULL.
 * The offsets receive gradient only through the bilinear weights (floor is detached, :208); clamped positions
 * pass gradient where 0 <= p <= max (torch.clamp); |.|' = sign with sign(0) = 0.
 * ws: workspace of din_dynamic_infer_bwd_ws_floats(...) floats.  Same shape limits as the forward.  Four launches.
 */
DIN_API long long din_dynamic_infer_bwd_ws_floats(int b, int t, int n, int c, int kt, int kn, int scale_factor);
DIN_API int din_dynamic_infer_bwd_f32(const float* x, const float* w_tap, const float* b_cat, const float* dy,
                                      float* dx, float* dw_tap, float* db_cat, float* dcoef, float* ws, int b, int t,
                                      int n, int c, int kt, int kn, int ratio, int scale_factor,
                                      const float* coef_ptr, float coef_scalar, const int32_t* n_valid,
                                      void* stream);

/* ---- the loader's per-frame work on the device: JPEG decode + resize -------------------- */

/*
 * PIL-exact bilinear resize of interleaved uint8 RGB frames: src [n, h, w, 3] -> dst [n, oh, ow, 3], bit-identical to
 * Pillow's Image.resize((ow, oh), Image.BILINEAR) (antialiased when shrinking; ImagingResample, 8 bits per channel:
 * 22-bit fixed-point coefficients, horizontal pass into a uint8 intermediate, then the vertical pass).
 * Replaces transforms.functional.resize(img, image_size) at volleyball.py:239 / collective.py:183.
 * tmp: n * h * ow * 3 bytes of device scratch, needed only when both axes change (NULL otherwise).  h == oh && w == ow: a copy.
 */
DIN_API int din_resize_bilinear_u8(const uint8_t* src, int n, int h, int w, uint8_t* dst, int oh, int ow, uint8_t* tmp,
                                   void* stream);

/* The nvJPEG backend in use (nvjpegBackend_t: 3 = hardware engine, 2 = GPU-assisted Huffman, 0 / 1 = hybrid with CPU
 * Huffman), -1 before the first decode.  DIN_NVJPEG_BACKEND forces one; unset: the first of 3, 2, 0 that initialises. */
DIN_API int din_jpeg_backend(void);

/* Height / width of a JPEG stream in host memory (nvjpegGetImageInfo). */
DIN_API int din_jpeg_image_info(const unsigned char* jpeg, size_t nbytes, int* h, int* w);

/*
 * n JPEG files held in HOST memory -> frames [n, out_h, out_w, 3] uint8 on the device: nvJPEG batched decode (library
 * code, resolved with dlopen at the first call; DIN_ERR_UNSUPPORTED when libnvjpeg.so.12 is absent) followed by
 * din_resize_bilinear_u8 for the frames whose size differs from the target; frames already at the target size are decoded
 * straight into `frames`.  The result is what the uint8 stems ingest (din_stem_conv_nhwc_u8, x_is_u8 of
 * din_conv3x3_stem_pair_nhwc_f16).  Replaces Image.open + resize + np.array of volleyball.py:237-240 / collective.py:181-184.
 * workspace: device scratch for the frames that need resizing: sum over those frames of
 *            align256(h * w * 3) + align256(h * out_w * 3) bytes (din_jpeg_image_info gives h, w); may be NULL if none does.
 * cpu_threads: host threads nvJPEG may use for the Huffman stage of its hybrid backend.
 */
DIN_API int din_jpeg_decode_resize_u8(const unsigned char* const* jpeg, const size_t* nbytes, int n, uint8_t* frames,
                                      int out_h, int out_w, uint8_t* workspace, size_t workspace_bytes, int cpu_threads,
                                      void* stream);

/* ---- backward of the backbone (second slice: VGG-16) ------------------------- */

/*
 * Weight (and bias) gradient of a stride-1 convolution (kw in {1, 3, 5, 7}, kh <= 7: every filter shape of the three
 * backbones -- 1x1, 3x3, 5x5, 1x7, 7x1; any padding) on the tcgen05 tensor cores:
 *     dw[co][ky][kx][ci] += inv_scale * sum_{img,y,x} dz[img,y,x,co] * x[img, y+ky-pad_h, x+kx-pad_w, ci]
 *     dbias[co]          += inv_scale * sum_{img,y,x} dz[img,y,x,co]                      (dbias may be NULL)
 * Replaces autograd's conv2d weight/bias backward for vgg16.features.* (backbone.py:88-99) when the backbone is
 * trained (scripts/train_volleyball_stage2_dynamic.py:12).  The K dimension of this GEMM is the pixel index, so
 * both NHWC fp16 operands are consumed as MN-major UMMA tiles straight from the forward kernel's TMA boxes.
 * x  : fp16 NHWC [n, h, w, x_c_stride], channels [0, c_in) used (the conv's saved input), c_in % 8 == 0 (a partial
 *      last 64-channel block is zero-filled by the TMA unit: Inception-v3's 32 / 48 / 80 / 96 / 160 / 288 channels)
 * dz : fp16 NHWC [n, oh, ow, dz_c_stride], oh = h + 2*pad_h - kh + 1 (gradient w.r.t. the conv output BEFORE ReLU,
 *      possibly multiplied by a loss scale S; then *inv_scale = 1/S, a device scalar; NULL = 1)
 * dw : fp32 [c_out][kh][kw][c_in] -- ACCUMULATED with atomics: zero-fill before the first call of a step (the
 *      frames of a step may arrive in several calls); dbias likewise.
 */
DIN_API int din_conv2d_wgrad_nhwc_f16(const void* x, const void* dz, float* dw, float* dbias, const float* inv_scale,
                                      int n, int h, int w, int c_in, int x_c_stride, int c_out, int dz_c_stride,
                                      int kh, int kw, int pad_h, int pad_w, void* stream);

/*
 * RoIAlign backward: dfm[frame, corner, :] += bilinear weight * dcrops[m, bin, :] for the four corners of every
 * valid sample point (extrapolated samples are constants: no gradient).  fp32 vector atomics.
 * Replaces the backward of the external crop_and_resize extension (infer_model.py:178-181).
 * dcrops: fp32 [m][crop_h*crop_w][d] (the layout of din_roi_align_nhwc_f16's output);  dfm: fp32 NHWC
 * [n_img, h, w, fm_c_stride], zero-filled by the caller (or holding a gradient to add to).
 */
DIN_API int din_roi_align_bwd_f32(const float* dcrops, const float* boxes, const int32_t* box_ind, float* dfm,
                                  int n_img, int h, int w, int d, int fm_c_stride, int m, int crop_h, int crop_w,
                                  void* stream);

/*
 * fp32 gradient -> fp16 times a dynamic power-of-two loss scale S = 2^floor(log2(target / max|x|)), so that the
 * fp16 tensor-core backward of the backbone stays inside fp16's normal range; S is undone exactly by the
 * weight-gradient kernels (inv_scale).  scale_ws: fp32 [4] device workspace: [0] max|x| bits, [1] S, [2] 1/S.
 */
DIN_API int din_grad_to_f16(const float* x, void* y, float* scale_ws, long long count, float target, void* stream);

/*
 * Backward of ReLU (+ MaxPool2d(2,2)) between two backbone convolutions, NHWC fp16:
 *   pool == 0: dz = dy * [y > 0];     pool != 0: dy is [n, h/2, w/2, c]; the gradient goes to the FIRST maximum of
 *   each 2x2 window (torch's max_pool2d index rule) if it is positive; uncovered rows / columns get 0.
 * y: the saved ReLU output [n, h, w, c].  Replaces autograd through nn.ReLU / nn.MaxPool2d of vgg16.features.
 */
DIN_API int din_relu_pool_bwd_nhwc_f16(const void* y, const void* dy, void* dz, int n, int h, int w, int c, int pool,
                                       void* stream);

/*
 * Weight / bias gradient of the stem convolution -- VGG-16 features.0 (64 x 3 x 3x3, stride 1, pad 1), ResNet-18
 * conv1 (64 x 3 x 7x7, stride 2, pad 3; BN folded: the caller un-folds) or Inception-v3 Conv2d_1a_3x3 (32 x 3 x 3x3,
 * stride 2, pad 0; dw [32][3][3][3], dz [n, oh, ow, 32]) -- with prep_images recomputed on the raw
 * frames (fp32 NCHW, or uint8 NHWC when x_is_u8), on the tensor cores:
 *   dw [64][3][kh][kw] (OIHW) += inv_scale * sum_pixels dz (x) prep(x)_shifted;   dbias [64] += inv_scale * sum dz.
 * dz: fp16 NHWC [n, oh, ow, 64].  ACCUMULATES (atomics): zero-fill before the first call of a step.
 */
DIN_API int din_stem_wgrad(const void* x, int x_is_u8, const void* dz, float* dw, float* dbias, const float* inv_scale,
                           int n, int h, int w, int c_out, int kh, int kw, int stride, int pad, int prep,
                           void* stream);

/* ---- ResNet-18 backward helpers (stride-2 convolutions, 3x3/2 max-pool, eval-mode BatchNorm folded into the convs) -- */

/*
 * dst[n, 2*oy, 2*ox, :] (+)= src[n, oy, ox, :]  (NHWC fp16; dst is [n,h,w,c], src [n,oh,ow,c]).
 * accumulate == 0: zero insertion -- the gradient of a stride-2 convolution's output laid out on its input grid, so
 *   that dX = din_conv2d_nhwc_f16(dZ_up, rot180(W)^T) and dW = din_conv2d_wgrad_nhwc_f16(X, dZ_up) are the stride-1
 *   kernels (resnet18 layer{2,3,4}.0.conv1, backbone.py:115-132);  accumulate != 0: adds src at the even positions
 *   (data gradient of the 1x1 stride-2 shortcut, layer{2,3,4}.0.downsample.0).
 */
DIN_API int din_scatter2_nhwc_f16(const void* src, void* dst, int n, int h, int w, int c, int oh, int ow, int accumulate,
                                  void* stream);

/* y = a + b, fp16, count elements (multiple of 8): merging the two gradient branches of a residual block. */
DIN_API int din_add_f16(const void* a, const void* b, void* y, long long count, void* stream);

/*
 * Backward of resnet18.relu + resnet18.maxpool (MaxPool2d(3, 2, 1)): x = the saved ReLU output [n,h,w,c], dy the
 * gradient of the pooled map [n,oh,ow,c] -> dz [n,h,w,c]: every window sends its gradient to its FIRST maximum (scan
 * order, padding never wins, as torch), masked by x > 0.
 */
DIN_API int din_maxpool3s2_relu_bwd_nhwc_f16(const void* x, const void* dy, void* dz, int n, int h, int w, int c,
                                             void* stream);

/* ---- Inception-v3 backward helpers (concat slices, 3x3/2 max-pools without padding, the multiscale resize) ---------- */

/* The same routing for any MaxPool2d(3, 2, pad) with pad in {0, 1} over channel slices (x / dy / dz live in buffers with
 * their own channel strides): F.max_pool2d(x, 3, 2) of the Inception-v3 trunk (backbone.py:50,56) and of Mixed_6a's pool
 * branch.  x is a ReLU output everywhere in these backbones, so the x > 0 mask is the ReLU backward of the layer below.
 * accumulate != 0: dz += (another branch already wrote its share of the same tensor's gradient).  c <= 288. */
DIN_API int din_maxpool3s2_bwd_nhwc_f16(const void* x, const void* dy, void* dz, int n, int h, int w, int c,
                                        int x_c_stride, int dy_c_stride, int dz_c_stride, int pad, int accumulate,
                                        void* stream);

/* dz[r, 0:c] = dy[r, 0:c] * [y[r, 0:c] > 0], rows r of three channel-strided fp16 buffers: the ReLU backward of a branch
 * whose output is a channel slice of an Inception block's concat buffer (torch.cat backward + threshold_backward). */
DIN_API int din_relu_bwd_slice_nhwc_f16(const void* y, const void* dy, void* dz, long long rows, int c, int y_c_stride,
                                        int dy_c_stride, int dz_c_stride, void* stream);

/* Backward of din_upsample_bilinear_nhwc_f16 (align_corners=True): dx [n,h,w,c] = sum over the outputs [n,oh,ow,c] that
 * read each source pixel of weight * dy, gathered per source pixel with the forward's fp32 weights (no atomics).
 * Replaces upsample_bilinear2d_backward under infer_model.py:169. */
DIN_API int din_upsample_bilinear_bwd_nhwc_f16(const void* dy, void* dx, int n, int h, int w, int c, int dy_c_stride,
                                               int dx_c_stride, int oh, int ow, void* stream);

/* db[c] += inv_scale * sum over rows of dz[r, c] (fp16 in, fp32 atomics out; dz may be a channel slice): the bias /
 * BatchNorm-shift gradient of a branch whose bias is applied outside its convolution (the Inception pool branches). */
DIN_API int din_colsum_nhwc_f16(const void* dz, float* db, long long rows, int c, int c_stride, const float* inv_scale,
                                void* stream);

/*
 * Gradient of an eval-mode BatchNorm's weight when the BN is folded into its convolution (z = gamma*xhat + beta):
 *   dgamma[c] += inv_scale * sum_p dz[p,c] * ((zsrc[p,c] - sub[p,c]) - beta[c]) / gamma[c]      (sub may be NULL)
 * zsrc: the saved activation (post-ReLU serves: dz is zero wherever the ReLU clipped); sub: the residual that was
 * added before the ReLU (BasicBlock.bn2).  d(beta) is the convolution's bias gradient.  Rows of c fp16 values.
 */
DIN_API int din_bn_gamma_grad_f16(const void* dz, const void* zsrc, const void* sub, const float* gamma,
                                  const float* beta, float* dgamma, const float* inv_scale, long long rows, int c,
                                  void* stream);

/* w[r][:] *= scale[r]: un-folds the BN scale from a folded convolution's weight gradient. */
DIN_API int din_scale_rows_f32(float* w, const float* scale, long long rows, long long cols, void* stream);

/* d(gamma) of an eval-mode BatchNorm folded into its convolution, from quantities the backward already has:
 *   dgamma[r] = rsqrt(running_var[r] + eps) * ( <w[r,:], dwf[r,:]> - running_mean[r] * dbeta[r] )
 * w: the UN-folded fp32 weight [rows][cols] (OIHW rows); dwf: gradient w.r.t. the folded weight in the same order
 * (before din_scale_rows_f32); dbeta: the folded convolution's bias gradient.  Overwrites dgamma [rows].
 * Replaces autograd of nn.BatchNorm2d.weight in eval mode (backbone.py:115-132 under cfg.set_bn_eval). */
DIN_API int din_bn_fold_grads_f32(const float* w, const float* dwf, const float* dbeta, const float* running_mean,
                          const float* running_var, float eps, float* dgamma, int rows, long long cols, void* stream);

/* ---- BatchNorm2d on batch statistics (module.train() without cfg.set_bn_eval: config.py:80, train_net_dynamic.py:161;
 *      nn.BatchNorm2d layers of backbone/backbone.py:115-132).  z: raw convolution output [rows][c] (NHWC), fp16 or
 *      (z_is_f32) fp32 -- fp32 keeps the pre-normalisation values exact, so that y is rounded once, like the folded
 *      eval-mode epilogue; c % 8 == 0.  csrc/bn_train.cu. */
/* sum[c], sumsq[c] (fp64, overwritten) = per-channel sum / sum of squares of z over all rows */
DIN_API int din_bn_stats(const void* z, int z_is_f32, long long rows, int c, double* sum, double* sumsq, void* stream);
/* mean, biased var -> invstd = 1/sqrt(var + eps); scale = gamma * invstd; shift = beta - mean * scale;
 * running_mean / running_var (or both NULL) updated in place: (1 - momentum) * old + momentum * new (unbiased var). */
DIN_API int din_bn_finalize_f32(const double* sum, const double* sumsq, long long count, const float* gamma,
                        const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                        float* scale, float* shift, float* mean, float* invstd, int c, void* stream);
/* y = [relu]( z * scale[c] + shift[c] [+ residual] ), fp16 */
DIN_API int din_bn_apply(const void* z, int z_is_f32, const float* scale, const float* shift, const void* residual,
                 void* y, long long rows, int c, int relu, void* stream);
/* g = dY (already masked by the ReLU, fp16, times the loss scale S); xhat = (z - mean) * invstd:
 *   dz = gamma * invstd * (g - sum(g)/rows - xhat * sum(g*xhat)/rows)            (fp16)
 *   dbeta += sum(g) * inv_scale,  dgamma += sum(g*xhat) * inv_scale               (fp32 [c], or both NULL)
 * sums: fp32 [2*c] scratch (overwritten).  inv_scale: device scalar 1/S or NULL. */
DIN_API int din_bn_bwd(const void* g, const void* z, int z_is_f32, const float* mean, const float* invstd,
               const float* gamma, float* sums, void* dz, float* dbeta, float* dgamma, const float* inv_scale,
               long long rows, int c, void* stream);

/* ---- Dynamic_TCE_volleyball's context encoding (infer_model.py:237-468, infer_module/TCE_STBiP_module.py:224-310) ----
 * Every actor attends over its frame's feature map, per head:
 *   a[n][px] = <q[n], x[px]>,  A = softmax over px,  ctx[n] = sum_px A[n][px] x[px],   x = img + posbias.
 * q, ctx  : fp32 [heads][frames*n][128]  (emb_roi output per head / the attended context)
 * img     : fp32 [frames][pixels][heads*128]  = the heads' downsample2 1x1 convolutions of the feature map, without bias
 *           (one din_conv2d_nhwc_f16 with out_f32)
 * posbias : fp32 [pixels][heads*128] = downsample2(Context_PositionEmbeddingSine table) + bias (a constant of the plan)
 * Replaces: torch.matmul / F.softmax / torch.matmul at TCE_STBiP_module.py:274-280.  n <= 16, n*pixels*4 B <= ~130 KB. */
DIN_API int din_context_attention_f32(const float* q, const float* img, const float* posbias, float* ctx, int frames,
                                      int n, int pixels, int heads, void* stream);

/* Backward of din_context_attention_f32 (autograd through torch.matmul / F.softmax / torch.matmul at
 * TCE_STBiP_module.py:274-280 when scripts/train_volleyball_stage2_dynamic_tce.py trains Dynamic_TCE_volleyball):
 *   dq [heads][frames*n][128], dimg [frames][pixels][heads*128] (the gradient of the heads' downsample2 output: the feature map
 *   is key AND value, both terms are summed) from dctx [heads][frames*n][128]; scores / softmax recomputed with the forward's
 *   arithmetic; no atomics.  dq_add (or NULL): added to dq -- the gradient q also receives through layernorm1's residual. */
DIN_API int din_context_attention_bwd_f32(const float* q, const float* img, const float* posbias, const float* dctx,
                                          const float* dq_add, float* dq, float* dimg, int frames, int n, int pixels,
                                          int heads, void* stream);

/* ---- data-parallel training: gradients -> one flat fp32 buffer (csrc/flat.cu) -----------------------------------
 * flat[dst_offset + i] = scale * src[i] for every job, ALL jobs in one launch (<= 96 per launch, more are split).
 * The flat buffer is what the step's all-reduce runs on (ncclAllReduce through torch.distributed); `scale` carries
 * the rank's weight local_clips / global_clips so that the SUM over ranks is the global-batch mean gradient.
 * Replaces: the reduce_add_coalesced on GPU 0 inside nn.DataParallel's backward (train_net_dynamic.py:96,220-224). */
typedef struct DinFlatJob {
  const void* src;          /* fp32 gradient tensor (device, contiguous) */
  int64_t dst_offset;       /* element offset into `flat` */
  int64_t numel;
} DinFlatJob;
DIN_API int din_pack_flat_f32(const DinFlatJob* jobs, int n_jobs, float* flat, float scale, void* stream);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* DIN_SM100_H_ */
