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
 * Implicit-GEMM convolution on the tcgen05 tensor cores (TMA-staged fp16 operands, fp32 TMEM
 * accumulators), fused bias (+ residual) (+ ReLU) epilogue.  Also serves as the dense GEMM of the
 * path (kh = kw = 1, n = h = 1, w = rows).
 * Replaces: every Conv2d(+BN)(+ReLU) of the truncated torchvision backbones (backbone.py:10-132) and
 *   nn.Linear fc_emb_1 (infer_model.py:50,184).
 */
typedef struct DinConvDesc {
  int32_t n, h, w;          /* input spatial extent (NHWC) */
  int32_t c_in;             /* input channels used by this conv; multiple of 64 */
  int32_t x_c_stride;       /* channel count of the buffer x lives in (>= c_in; x may be a channel slice) */
  int32_t c_out;            /* output channels; multiple of 8 */
  int32_t y_c_stride;       /* channel count of the buffer y lives in (>= c_out; concat-by-offset writes) */
  int32_t kh, kw;           /* filter taps */
  int32_t stride;           /* 1 or 2 */
  int32_t pad_h, pad_w;     /* zero padding */
  int32_t relu;             /* fuse ReLU */
  int32_t out_f32;          /* 0: y is fp16, 1: y is fp32 */
} DinConvDesc;

/*
 * x        : fp16, element (img, yy, xx, c) at x[((img*h + yy)*w + xx)*x_c_stride + c]; 16-byte aligned
 * w_packed : fp16 [c_out][kh][kw][c_in] (K-major rows; see din_pack_conv_weight_f16); 16-byte aligned
 * bias     : fp32 [c_out] or NULL
 * residual : fp16, same indexing as y with y_c_stride, added before ReLU; or NULL
 * y        : fp16 or fp32 (out_f32), element (img, oy, ox, co) at y[((img*oh + oy)*ow + ox)*y_c_stride + co]
 */
DIN_API int din_conv2d_nhwc_f16(const DinConvDesc* desc, const void* x, const void* w_packed, const float* bias,
                        const void* residual, void* y, void* stream);

/* OIHW fp32 [c_out, c_in, kh, kw] (optionally scaled per output channel by `scale`, for BN folding;
 * NULL = 1) -> fp16 [c_out][kh][kw][c_in_padded], zero-filled for c_in <= c < c_in_padded.
 * Device-to-device. */
DIN_API int din_pack_conv_weight_f16(const float* w_oihw, const float* scale, void* w_packed, int c_out, int c_in,
                             int c_in_padded, int kh, int kw, void* stream);

/*
 * Max pooling, NHWC fp16.  Replaces nn.MaxPool2d(2,2) in vgg16.features, resnet18.maxpool (3,2,1),
 * F.max_pool2d(3,2) at backbone.py:50,56.  Padding elements never win (-inf), as in torch.
 */
DIN_API int din_maxpool2d_nhwc_f16(const void* x, void* y, int n, int h, int w, int c, int k, int stride, int pad,
                           void* stream);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* DIN_SM100_H_ */
