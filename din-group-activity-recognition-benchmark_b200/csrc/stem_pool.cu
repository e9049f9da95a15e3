// stem_pool.cu — the two HBM-bound backbone helpers around the tcgen05 convolutions:
//   * stem convolution (3 input channels) fused with prep_images and the NCHW fp32 -> NHWC fp16
//     relayout  (reference: utils.py:8-19 + first conv of backbone.py:88-99 / 115-132 / 44);
//   * NHWC fp16 max pooling (reference: nn.MaxPool2d inside vgg16.features / resnet18.maxpool,
//     F.max_pool2d at backbone.py:50,56).
#include <cfloat>
#include <cstdlib>

#include "din_common.cuh"

namespace {

constexpr int kStemTileH = 4;    // output rows per CTA
constexpr int kStemTileW = 64;   // output cols per CTA
constexpr int kStemThreads = 256;

// One CTA = 4 x 64 output pixels x c_out channels.  Warp w: output row (w & 3), channel half (w >> 2);
// lane l: pixels l and l + 32 of that row.  The (prep'd, zero-padded) input patch and the whole filter
// bank live in shared memory; weight reads are warp-uniform broadcasts.
template <int CPT>  // output channels per thread (c_out = 2 * CPT)
__global__ void __launch_bounds__(kStemThreads)
stem_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 __half* __restrict__ y, int h, int wd, int oh, int ow, int kh, int kw, int stride, int pad,
                 int relu, int prep) {
  extern __shared__ float smem_f[];
  const int c_out = 2 * CPT;
  const int K = 3 * kh * kw;
  const int in_th = (kStemTileH - 1) * stride + kh;
  const int in_tw = (kStemTileW - 1) * stride + kw;
  const int in_tw_p = in_tw | 1;  // odd row pitch: fewer bank conflicts for strided reads
  float* ws = smem_f;                       // [K][c_out]
  float* xs = smem_f + K * c_out;           // [3][in_th][in_tw_p]

  const int img = blockIdx.z;
  const int oy0 = blockIdx.y * kStemTileH;
  const int ox0 = blockIdx.x * kStemTileW;

  // weights: OIHW [c_out][3][kh][kw] -> ws[(c*kh + ky)*kw + kx][o]
  for (int i = threadIdx.x; i < K * c_out; i += kStemThreads) {
    const int o = i / K;
    const int k = i - o * K;
    ws[k * c_out + o] = w[i];
  }
  const int iy0 = oy0 * stride - pad;
  const int ix0 = ox0 * stride - pad;
  const float* xi = x + static_cast<size_t>(img) * 3 * h * wd;
  for (int i = threadIdx.x; i < 3 * in_th * in_tw; i += kStemThreads) {
    const int c = i / (in_th * in_tw);
    const int r = i - c * (in_th * in_tw);
    const int yy = r / in_tw;
    const int xx = r - yy * in_tw;
    const int gy = iy0 + yy, gx = ix0 + xx;
    float v = 0.0f;
    if (gy >= 0 && gy < h && gx >= 0 && gx < wd) {
      v = __ldg(xi + (static_cast<size_t>(c) * h + gy) * wd + gx);
      // prep_images, same three roundings as the reference (utils.py:14-17): div, sub, mul
      if (prep) v = __fmul_rn(__fsub_rn(__fdiv_rn(v, 255.0f), 0.5f), 2.0f);
    }
    xs[(c * in_th + yy) * in_tw_p + xx] = v;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = warp & 3;
  const int half_id = warp >> 2;
  float acc0[CPT], acc1[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const float b = bias ? __ldg(bias + half_id * CPT + j) : 0.0f;
    acc0[j] = b;
    acc1[j] = b;
  }
  const float* wbase = ws + half_id * CPT;
  for (int c = 0; c < 3; ++c) {
    for (int ky = 0; ky < kh; ++ky) {
      const float* xrow = xs + (c * in_th + row * stride + ky) * in_tw_p;
      for (int kx = 0; kx < kw; ++kx) {
        const float a0 = xrow[lane * stride + kx];
        const float a1 = xrow[(lane + 32) * stride + kx];
        const float4* wv = reinterpret_cast<const float4*>(wbase + ((c * kh + ky) * kw + kx) * c_out);
#pragma unroll
        for (int j = 0; j < CPT / 4; ++j) {
          const float4 t = wv[j];
          acc0[4 * j + 0] = fmaf(a0, t.x, acc0[4 * j + 0]);
          acc0[4 * j + 1] = fmaf(a0, t.y, acc0[4 * j + 1]);
          acc0[4 * j + 2] = fmaf(a0, t.z, acc0[4 * j + 2]);
          acc0[4 * j + 3] = fmaf(a0, t.w, acc0[4 * j + 3]);
          acc1[4 * j + 0] = fmaf(a1, t.x, acc1[4 * j + 0]);
          acc1[4 * j + 1] = fmaf(a1, t.y, acc1[4 * j + 1]);
          acc1[4 * j + 2] = fmaf(a1, t.z, acc1[4 * j + 2]);
          acc1[4 * j + 3] = fmaf(a1, t.w, acc1[4 * j + 3]);
        }
      }
    }
  }
  const int oy = oy0 + row;
  if (oy >= oh) return;
#pragma unroll
  for (int pxi = 0; pxi < 2; ++pxi) {
    const int ox = ox0 + lane + 32 * pxi;
    if (ox >= ow) continue;
    const float* a = pxi ? acc1 : acc0;
    __half* yp = y + ((static_cast<size_t>(img) * oh + oy) * ow + ox) * c_out + half_id * CPT;
#pragma unroll
    for (int j = 0; j < CPT; j += 8) {
      uint4 o;
      __half2* h2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v0 = a[j + 2 * e], v1 = a[j + 2 * e + 1];
        if (relu) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
        h2[e] = __floats2half2_rn(v0, v1);
      }
      *reinterpret_cast<uint4*>(yp + j) = o;
    }
  }
}

// 8 channels (one 16-byte vector) of one output pixel per thread.  AVG = 0: max pooling (padding never
// wins); AVG = 1: average pooling with count_include_pad=True (torch's default, what torchvision's
// Inception blocks use): always divides by k*k.
template <int AVG>
__global__ void pool_nhwc_kernel(const __half* __restrict__ x, __half* __restrict__ y, int n, int h, int w, int c8, int xs8,
                 int ys8, int oh, int ow, int k, int stride, int pad) {
  const size_t total = static_cast<size_t>(n) * oh * ow * c8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % c8);
    size_t t = i / c8;
    const int ox = static_cast<int>(t % ow);
    t /= ow;
    const int oy = static_cast<int>(t % oh);
    const int img = static_cast<int>(t / oh);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = AVG ? 0.0f : -FLT_MAX;
    for (int ky = 0; ky < k; ++ky) {
      const int iy = oy * stride - pad + ky;
      if (iy < 0 || iy >= h) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ix = ox * stride - pad + kx;
        if (ix < 0 || ix >= w) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + ((static_cast<size_t>(img) * h + iy) * w + ix) * xs8 + cv);
        const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h2[e]);
          if (AVG) { acc[2 * e] += f.x; acc[2 * e + 1] += f.y; }
          else { acc[2 * e] = fmaxf(acc[2 * e], f.x); acc[2 * e + 1] = fmaxf(acc[2 * e + 1], f.y); }
        }
      }
    }
    uint4 o;
    __half2* oh2 = reinterpret_cast<__half2*>(&o);
    const float sc = AVG ? 1.0f / static_cast<float>(k * k) : 1.0f;
#pragma unroll
    for (int e = 0; e < 4; ++e) oh2[e] = __floats2half2_rn(acc[2 * e] * sc, acc[2 * e + 1] * sc);
    reinterpret_cast<uint4*>(y)[((static_cast<size_t>(img) * oh + oy) * ow + ox) * ys8 + cv] = o;
  }
}

// The same for the window shapes the backbones use (3x3 stride 1 / 2, 2x2 stride 2), one output per thread, huge grid:
// every tap is an UNCONDITIONAL load from a clamped coordinate (selected against the padding value afterwards), so all
// K*K loads are in flight at once; pool_nhwc_kernel's per-tap `continue` put each load in its own branch and its
// grid-stride loop kept one output per thread in flight: 4.5x off the HBM roofline on Inception's average pools.
// Max pooling compares packed halves (exact); the average accumulates in fp32 in the window's scan order.
template <int AVG, int K, int S>
__global__ void __launch_bounds__(256)
pool_fixed_kernel(const __half* __restrict__ x, __half* __restrict__ y, int n, int h, int w, int c8, int xs8, int ys8,
                  int oh, int ow, int pad) {
  const size_t total = static_cast<size_t>(n) * oh * ow * c8;
  const size_t i = blockIdx.x * static_cast<size_t>(256) + threadIdx.x;
  if (i >= total) return;
  const int cv = static_cast<int>(i % c8);
  size_t t = i / c8;
  const int ox = static_cast<int>(t % ow);
  t /= ow;
  const int oy = static_cast<int>(t % oh);
  const int img = static_cast<int>(t / oh);
  const uint4* base = reinterpret_cast<const uint4*>(x) + static_cast<size_t>(img) * h * w * xs8 + cv;
  uint4 v[K * K];
  bool ok[K * K];
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const int iy = oy * S - pad + ky, ix = ox * S - pad + kx;
      ok[ky * K + kx] = iy >= 0 && iy < h && ix >= 0 && ix < w;
      const int cy = min(max(iy, 0), h - 1), cx = min(max(ix, 0), w - 1);
      v[ky * K + kx] = __ldg(base + (static_cast<size_t>(cy) * w + cx) * xs8);
    }
  }
  uint4 o;
  __half2* oh2 = reinterpret_cast<__half2*>(&o);
  if (AVG) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < K * K; ++q) {
      const __half2* h2 = reinterpret_cast<const __half2*>(&v[q]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        acc[2 * e] += ok[q] ? f.x : 0.0f;
        acc[2 * e + 1] += ok[q] ? f.y : 0.0f;
      }
    }
    const float sc = 1.0f / static_cast<float>(K * K);       // count_include_pad = True
#pragma unroll
    for (int e = 0; e < 4; ++e) oh2[e] = __floats2half2_rn(acc[2 * e] * sc, acc[2 * e + 1] * sc);
  } else {
    const __half2 ninf = __half2half2(__ushort_as_half(static_cast<unsigned short>(0xFC00)));   // -inf: padding never wins
    __half2 m[4] = {ninf, ninf, ninf, ninf};
#pragma unroll
    for (int q = 0; q < K * K; ++q) {
      const __half2* h2 = reinterpret_cast<const __half2*>(&v[q]);
#pragma unroll
      for (int e = 0; e < 4; ++e) m[e] = __hmax2(m[e], ok[q] ? h2[e] : ninf);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) oh2[e] = m[e];
  }
  reinterpret_cast<uint4*>(y)[((static_cast<size_t>(img) * oh + oy) * ow + ox) * ys8 + cv] = o;
}

// 3x3 stride-1 average pooling (Inception's pool branches) with a vertical sliding window: one thread owns 8 channels of
// one output COLUMN segment of kAvgRows rows and keeps the 3x3 window's nine 16-byte vectors in registers, loading only
// the three new ones per output row.  With one output per thread every input row is fetched by three different rows of
// CTAs (3x L2 -> SM traffic, the measured limit: 3.6 ms for 6.5 GB of compulsory bytes per step); here 1.25x.  The nine
// values are still summed in scan order, so the result is bit-identical to pool_fixed_kernel<1, 3, 1>.
constexpr int kAvgRows = 8;

// BIAS: y = (relu)(avg + bias[c]) -- the tail of an Inception pool branch whose 1x1 convolution ran before the pool
// (din_conv2d_branches_nhwc_f16).
template <bool BIAS>
__global__ void __launch_bounds__(256)
avgpool3s1_strip_kernel(const __half* __restrict__ x, __half* __restrict__ y, int n, int h, int w, int c8, int xs8, int ys8,
                        const float* __restrict__ bias = nullptr, int relu = 0) {
  const int strips = (h + kAvgRows - 1) / kAvgRows;
  const size_t total = static_cast<size_t>(n) * strips * w * c8;
  const size_t i = blockIdx.x * static_cast<size_t>(256) + threadIdx.x;
  if (i >= total) return;
  const int cv = static_cast<int>(i % c8);
  size_t t = i / c8;
  const int ox = static_cast<int>(t % w);
  t /= w;
  const int strip = static_cast<int>(t % strips);
  const int img = static_cast<int>(t / strips);
  const uint4* base = reinterpret_cast<const uint4*>(x) + static_cast<size_t>(img) * h * w * xs8 + cv;
  const bool okx[3] = {ox - 1 >= 0, true, ox + 1 < w};
  const int cx[3] = {max(ox - 1, 0), ox, min(ox + 1, w - 1)};
  uint4 win[3][3];
  bool oky[3];
  auto load_row = [&](int iy, uint4 (&row)[3]) -> bool {
    const int cy = min(max(iy, 0), h - 1);
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) row[kx] = __ldg(base + (static_cast<size_t>(cy) * w + cx[kx]) * xs8);
    return iy >= 0 && iy < h;
  };
  float bv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (BIAS) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * cv), b1 = __ldg(reinterpret_cast<const float4*>(bias) + 2 * cv + 1);
    bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
  }
  const int oy0 = strip * kAvgRows;
  oky[0] = load_row(oy0 - 1, win[0]);
  oky[1] = load_row(oy0, win[1]);
#pragma unroll
  for (int r = 0; r < kAvgRows; ++r) {
    const int oy = oy0 + r;
    if (oy >= h) break;
    oky[2] = load_row(oy + 1, win[2]);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&win[ky][kx]);
        const bool ok = oky[ky] && okx[kx];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h2[e]);
          acc[2 * e] += ok ? f.x : 0.0f;
          acc[2 * e + 1] += ok ? f.y : 0.0f;
        }
      }
    }
    uint4 o;
    __half2* oh2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float a = acc[2 * e] * (1.0f / 9.0f), b = acc[2 * e + 1] * (1.0f / 9.0f);
      if (BIAS) {
        a += bv[2 * e]; b += bv[2 * e + 1];
        if (relu) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
      }
      oh2[e] = __floats2half2_rn(a, b);
    }
    reinterpret_cast<uint4*>(y)[((static_cast<size_t>(img) * h + oy) * w + ox) * ys8 + cv] = o;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) { win[0][kx] = win[1][kx]; win[1][kx] = win[2][kx]; }
    oky[0] = oky[1];
    oky[1] = oky[2];
  }
}

// Bilinear resize, align_corners=True (F.interpolate at infer_model.py:169), NHWC fp16; torch's upsample_bilinear2d
// arithmetic: scale = (in - 1) / (out - 1), (1-ly)*((1-lx)*a + lx*b) + ly*((1-lx)*c + lx*d).
// One thread per (8 channels, output column, strip of kUpRows output rows): the two source rows of an
// output row stay in registers and are re-used by the next output row (an upscale by ~2 advances the source row by 0 or
// 1), so a strip of 8 output rows loads ~10 vectors instead of 32.  One output per thread was bound by L2 -> SM traffic
// (4 x 16 bytes loaded per 16 bytes stored: 6.7 GB per 80 Inception-v3 frames, 1.2 ms for a 1.7 GB output).  Same
// arithmetic per output element as the one-output-per-thread kernel it replaced: bit-identical.
constexpr int kUpRows = 8;

__global__ void __launch_bounds__(256)
upsample_bilinear_strip_kernel(const __half* __restrict__ x, __half* __restrict__ y, int n, int h, int w, int c8,
                               int xs8, int ys8, int oh, int ow) {
  const int strips = (oh + kUpRows - 1) / kUpRows;
  const size_t total = static_cast<size_t>(n) * strips * ow * c8;
  const size_t i = blockIdx.x * static_cast<size_t>(256) + threadIdx.x;
  if (i >= total) return;
  const float sy = oh > 1 ? static_cast<float>(h - 1) / static_cast<float>(oh - 1) : 0.0f;
  const float sx = ow > 1 ? static_cast<float>(w - 1) / static_cast<float>(ow - 1) : 0.0f;
  const int cv = static_cast<int>(i % c8);
  size_t t = i / c8;
  const int ox = static_cast<int>(t % ow);
  t /= ow;
  const int strip = static_cast<int>(t % strips);
  const int img = static_cast<int>(t / strips);
  const float fx = sx * static_cast<float>(ox);
  const int x0 = min(static_cast<int>(fx), w - 1);
  const int x1 = min(x0 + 1, w - 1);
  const float lx = fx - static_cast<float>(x0);
  const uint4* base = reinterpret_cast<const uint4*>(x) + static_cast<size_t>(img) * h * w * xs8 + cv;
  uint4 v00 = make_uint4(0, 0, 0, 0), v01 = v00, v10 = v00, v11 = v00;
  int cy0 = -1, cy1 = -1;
#pragma unroll
  for (int r = 0; r < kUpRows; ++r) {
    const int oy = strip * kUpRows + r;
    if (oy >= oh) break;
    const float fy = sy * static_cast<float>(oy);
    const int y0 = min(static_cast<int>(fy), h - 1);
    const int y1 = min(y0 + 1, h - 1);
    const float ly = fy - static_cast<float>(y0);
    if (y0 != cy0) {
      if (y0 == cy1) { v00 = v10; v01 = v11; }
      else {
        v00 = __ldg(base + (static_cast<size_t>(y0) * w + x0) * xs8);
        v01 = __ldg(base + (static_cast<size_t>(y0) * w + x1) * xs8);
      }
      cy0 = y0;
    }
    if (y1 != cy1) {
      v10 = __ldg(base + (static_cast<size_t>(y1) * w + x0) * xs8);
      v11 = __ldg(base + (static_cast<size_t>(y1) * w + x1) * xs8);
      cy1 = y1;
    }
    const __half2* a = reinterpret_cast<const __half2*>(&v00);
    const __half2* b = reinterpret_cast<const __half2*>(&v01);
    const __half2* c = reinterpret_cast<const __half2*>(&v10);
    const __half2* d = reinterpret_cast<const __half2*>(&v11);
    uint4 o;
    __half2* oh2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = __half22float2(a[e]), fb = __half22float2(b[e]);
      const float2 fc = __half22float2(c[e]), fd = __half22float2(d[e]);
      const float r0 = (1.0f - ly) * ((1.0f - lx) * fa.x + lx * fb.x) + ly * ((1.0f - lx) * fc.x + lx * fd.x);
      const float r1 = (1.0f - ly) * ((1.0f - lx) * fa.y + lx * fb.y) + ly * ((1.0f - lx) * fc.y + lx * fd.y);
      oh2[e] = __floats2half2_rn(r0, r1);
    }
    reinterpret_cast<uint4*>(y)[((static_cast<size_t>(img) * oh + oy) * ow + ox) * ys8 + cv] = o;
  }
}

}  // namespace

extern "C" int din_stem_conv_nchw_f32(const float* x, const float* w, const float* bias, void* y, int n, int h,
                                      int w_in, int c_out, int kh, int kw, int stride, int pad, int relu, int prep,
                                      void* stream) {
  DIN_CHECK_ARG(x && w && y, "din_stem_conv_nchw_f32: null pointer");
  DIN_CHECK_ARG(n > 0 && h > 0 && w_in > 0, "din_stem_conv_nchw_f32: bad extent n=%d h=%d w=%d", n, h, w_in);
  DIN_CHECK_ARG(c_out == 64 || c_out == 32, "din_stem_conv_nchw_f32: c_out=%d unsupported (32 or 64)", c_out);
  DIN_CHECK_ARG(kh >= 1 && kw >= 1 && kh * kw <= 49, "din_stem_conv_nchw_f32: bad filter %dx%d", kh, kw);
  DIN_CHECK_ARG(stride >= 1 && stride <= 2 && pad >= 0, "din_stem_conv_nchw_f32: bad stride/pad");
  DIN_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 15) == 0, "din_stem_conv_nchw_f32: y must be 16-byte aligned");
  DIN_CHECK_ARG(n <= 65535, "din_stem_conv_nchw_f32: n=%d exceeds grid.z", n);
  const int oh = (h + 2 * pad - kh) / stride + 1;
  const int ow = (w_in + 2 * pad - kw) / stride + 1;
  DIN_CHECK_ARG(oh > 0 && ow > 0, "din_stem_conv_nchw_f32: empty output");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    // production path: im2col-in-smem + tcgen05.mma (stem_tc.cu) for the three backbone stems.  Other
    // filter geometries (and DIN_STEM_SIMT=1, for A/B measurements) use the generic CUDA-core direct
    // convolution below — still a CUDA kernel of this library, not a fallback to another backend.
    const char* e = std::getenv("DIN_STEM_SIMT");
    if (!(e && e[0] == '1')) {
      const int rc = din_stem_tc_launch(x, 0, w, bias, y, n, h, w_in, c_out, kh, kw, stride, pad, relu, prep, st);
      if (rc != DIN_ERR_UNSUPPORTED) return rc;
    }
  }
  const int K = 3 * kh * kw;
  const int in_th = (kStemTileH - 1) * stride + kh;
  const int in_tw_p = ((kStemTileW - 1) * stride + kw) | 1;
  const size_t smem = (static_cast<size_t>(K) * c_out + 3 * in_th * in_tw_p) * sizeof(float);
  dim3 grid((ow + kStemTileW - 1) / kStemTileW, (oh + kStemTileH - 1) / kStemTileH, n);
  if (c_out == 64) {
    DIN_OPT_IN_SMEM(stem_conv_kernel<32>, smem);
    stem_conv_kernel<32><<<grid, kStemThreads, smem, st>>>(x, w, bias, static_cast<__half*>(y), h, w_in, oh, ow,
                                                           kh, kw, stride, pad, relu, prep);
  } else {
    DIN_OPT_IN_SMEM(stem_conv_kernel<16>, smem);
    stem_conv_kernel<16><<<grid, kStemThreads, smem, st>>>(x, w, bias, static_cast<__half*>(y), h, w_in, oh, ow,
                                                           kh, kw, stride, pad, relu, prep);
  }
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

static int pool_common(const char* who, int avg, const void* x, void* y, int n, int h, int w, int c, int x_c_stride,
                       int y_c_stride, int k, int stride, int pad, void* stream) {
  DIN_CHECK_ARG(x && y, "%s: null pointer", who);
  DIN_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0,
                "%s: bad shape n=%d h=%d w=%d c=%d (c must be a multiple of 8)", who, n, h, w, c);
  DIN_CHECK_ARG(x_c_stride >= c && y_c_stride >= c && x_c_stride % 8 == 0 && y_c_stride % 8 == 0,
                "%s: channel strides %d / %d must be >= c and multiples of 8", who, x_c_stride, y_c_stride);
  DIN_CHECK_ARG(k >= 1 && stride >= 1 && pad >= 0 && pad < k, "%s: bad window", who);
  DIN_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                "%s: pointers must be 16-byte aligned", who);
  const int oh = (h + 2 * pad - k) / stride + 1;
  const int ow = (w + 2 * pad - k) / stride + 1;
  DIN_CHECK_ARG(oh > 0 && ow > 0, "%s: empty output", who);
  const size_t total = static_cast<size_t>(n) * oh * ow * (c / 8);
  const int sms = din_num_sms();
  const int grid = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(sms > 0 ? sms : 148) * 16));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    const size_t blocks = (total + 255) / 256;
    const __half* xp = static_cast<const __half*>(x);
    __half* yp = static_cast<__half*>(y);
    bool done = true;
    if (blocks > 0x7FFFFFFFull) done = false;
#define DIN_POOL_FIXED(A, KK, SS)                                                                                  \
    pool_fixed_kernel<A, KK, SS><<<static_cast<int>(blocks), 256, 0, st>>>(xp, yp, n, h, w, c / 8, x_c_stride / 8, \
                                                                             y_c_stride / 8, oh, ow, pad)
    else if (avg && k == 3 && stride == 1 && pad == 1) {
      const size_t strip_threads = static_cast<size_t>(n) * ((h + kAvgRows - 1) / kAvgRows) * w * (c / 8);
      avgpool3s1_strip_kernel<false><<<static_cast<int>((strip_threads + 255) / 256), 256, 0, st>>>(
          xp, yp, n, h, w, c / 8, x_c_stride / 8, y_c_stride / 8);
    }
    else if (avg && k == 3 && stride == 1) DIN_POOL_FIXED(1, 3, 1);
    else if (!avg && k == 3 && stride == 2) DIN_POOL_FIXED(0, 3, 2);
    else if (!avg && k == 2 && stride == 2) DIN_POOL_FIXED(0, 2, 2);
    else if (!avg && k == 3 && stride == 1) DIN_POOL_FIXED(0, 3, 1);
    else done = false;
#undef DIN_POOL_FIXED
    if (done) {
      DIN_CHECK_CUDA(cudaGetLastError());
      return DIN_OK;
    }
  }
  if (avg)
    pool_nhwc_kernel<1><<<grid, 256, 0, st>>>(static_cast<const __half*>(x), static_cast<__half*>(y), n, h, w, c / 8,
                                              x_c_stride / 8, y_c_stride / 8, oh, ow, k, stride, pad);
  else
    pool_nhwc_kernel<0><<<grid, 256, 0, st>>>(static_cast<const __half*>(x), static_cast<__half*>(y), n, h, w, c / 8,
                                              x_c_stride / 8, y_c_stride / 8, oh, ow, k, stride, pad);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_stem_conv_nhwc_u8(const uint8_t* x, const float* w, const float* bias, void* y, int n, int h,
                                     int w_in, int c_out, int kh, int kw, int stride, int pad, int relu, int prep,
                                     void* stream) {
  DIN_CHECK_ARG(x && w && y, "din_stem_conv_nhwc_u8: null pointer");
  DIN_CHECK_ARG(n > 0 && h > 0 && w_in > 0, "din_stem_conv_nhwc_u8: bad extent n=%d h=%d w=%d", n, h, w_in);
  DIN_CHECK_ARG(stride >= 1 && stride <= 2 && pad >= 0, "din_stem_conv_nhwc_u8: bad stride/pad");
  DIN_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 15) == 0, "din_stem_conv_nhwc_u8: y must be 16-byte aligned");
  DIN_CHECK_ARG((h + 2 * pad - kh) / stride + 1 > 0 && (w_in + 2 * pad - kw) / stride + 1 > 0,
                "din_stem_conv_nhwc_u8: empty output");
  const int rc = din_stem_tc_launch(x, 1, w, bias, y, n, h, w_in, c_out, kh, kw, stride, pad, relu, prep,
                                    static_cast<cudaStream_t>(stream));
  if (rc == DIN_ERR_UNSUPPORTED)
    return din_set_error(DIN_ERR_UNSUPPORTED,
                         "din_stem_conv_nhwc_u8: only the three backbone stems are instantiated (64x3x3 s1, 64x7x7 s2, "
                         "32x3x3 s2), got c_out=%d %dx%d s%d", c_out, kh, kw, stride);
  return rc;
}

extern "C" int din_stem7x7_pool_nhwc_f16(const void* x, int x_is_u8, const float* w, const float* bias, void* y, int n,
                                         int h, int w_in, int prep, void* stream) {
  const char* who = "din_stem7x7_pool_nhwc_f16";
  DIN_CHECK_ARG(x && w && y, "%s: null pointer", who);
  DIN_CHECK_ARG(n > 0 && h >= 7 && w_in >= 7, "%s: bad extent n=%d h=%d w=%d", who, n, h, w_in);
  DIN_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 15) == 0, "%s: y must be 16-byte aligned", who);
  const int rc = din_stem_pool_tc_launch(x, x_is_u8, w, bias, y, n, h, w_in, prep, static_cast<cudaStream_t>(stream));
  if (rc == DIN_ERR_UNSUPPORTED)
    return din_set_error(DIN_ERR_UNSUPPORTED, "%s: image rows must be 16-byte multiples and x 16-byte aligned (w=%d): run "
                         "din_stem_conv_* and din_maxpool2d_nhwc_f16 separately", who, w_in);
  return rc;
}

extern "C" int din_maxpool2d_nhwc_f16(const void* x, void* y, int n, int h, int w, int c, int x_c_stride,
                                      int y_c_stride, int k, int stride, int pad, void* stream) {
  return pool_common("din_maxpool2d_nhwc_f16", 0, x, y, n, h, w, c, x_c_stride, y_c_stride, k, stride, pad, stream);
}

extern "C" int din_avgpool2d_nhwc_f16(const void* x, void* y, int n, int h, int w, int c, int x_c_stride,
                                      int y_c_stride, int k, int stride, int pad, void* stream) {
  return pool_common("din_avgpool2d_nhwc_f16", 1, x, y, n, h, w, c, x_c_stride, y_c_stride, k, stride, pad, stream);
}

extern "C" int din_avgpool3_bias_relu_nhwc_f16(const void* x, void* y, const float* bias, int n, int h, int w, int c,
                                               int x_c_stride, int y_c_stride, int relu, void* stream) {
  const char* who = "din_avgpool3_bias_relu_nhwc_f16";
  DIN_CHECK_ARG(x && y && bias, "%s: null pointer", who);
  DIN_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0,
                "%s: bad shape n=%d h=%d w=%d c=%d (c must be a multiple of 8)", who, n, h, w, c);
  DIN_CHECK_ARG(x_c_stride >= c && y_c_stride >= c && x_c_stride % 8 == 0 && y_c_stride % 8 == 0,
                "%s: channel strides %d / %d must be >= c and multiples of 8", who, x_c_stride, y_c_stride);
  DIN_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "%s: pointers must be 16-byte aligned", who);
  const size_t strip_threads = static_cast<size_t>(n) * ((h + kAvgRows - 1) / kAvgRows) * w * (c / 8);
  DIN_CHECK_ARG((strip_threads + 255) / 256 <= 0x7FFFFFFFull, "%s: too many blocks", who);
  avgpool3s1_strip_kernel<true><<<static_cast<int>((strip_threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), static_cast<__half*>(y), n, h, w, c / 8, x_c_stride / 8, y_c_stride / 8, bias, relu);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_upsample_bilinear_nhwc_f16(const void* x, void* y, int n, int h, int w, int c, int x_c_stride,
                                              int y_c_stride, int oh, int ow, void* stream) {
  DIN_CHECK_ARG(x && y, "din_upsample_bilinear_nhwc_f16: null pointer");
  DIN_CHECK_ARG(n > 0 && h > 0 && w > 0 && oh > 0 && ow > 0 && c > 0 && c % 8 == 0,
                "din_upsample_bilinear_nhwc_f16: bad shape n=%d h=%d w=%d c=%d -> %dx%d", n, h, w, c, oh, ow);
  DIN_CHECK_ARG(x_c_stride >= c && y_c_stride >= c && x_c_stride % 8 == 0 && y_c_stride % 8 == 0,
                "din_upsample_bilinear_nhwc_f16: channel strides must be >= c and multiples of 8");
  DIN_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                "din_upsample_bilinear_nhwc_f16: pointers must be 16-byte aligned");
  const size_t total = static_cast<size_t>(n) * ((oh + kUpRows - 1) / kUpRows) * ow * (c / 8);
  DIN_CHECK_ARG((total + 255) / 256 <= 0x7FFFFFFFull, "din_upsample_bilinear_nhwc_f16: too many blocks");
  upsample_bilinear_strip_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), static_cast<__half*>(y), n, h, w, c / 8, x_c_stride / 8, y_c_stride / 8, oh, ow);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}
