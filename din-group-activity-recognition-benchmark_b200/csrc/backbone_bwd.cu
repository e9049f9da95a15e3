// backbone_bwd.cu — the HBM-bound pieces of the backbone's backward pass (SURVEY.md §8f rank 1, VGG-16 slice):
//   * RoIAlign backward (bilinear scatter of the crop gradients into the feature-map gradient);
//   * fp32 -> scaled fp16 conversion of the feature-map gradient (dynamic power-of-two loss scale);
//   * ReLU (+ MaxPool2d(2,2)) backward between two convolutions;
//   * weight / bias gradient of the 3-channel stem convolution (fused with prep_images on its input side).
// The tensor-core pieces are conv_wgrad_tcgen05.cu (dW) and conv_tcgen05.cu itself (dX = the forward kernel run
// on the 180-degree-rotated, transposed filter).
//
// Reference: autograd through roi_align (external crop_and_resize backward; infer_model.py:178-181),
// vgg16.features (nn.ReLU(inplace), nn.MaxPool2d(2,2), nn.Conv2d; backbone.py:88-99) and prep_images
// (utils.py:8-19) in `total_loss.backward()` (train_net_dynamic.py:220-224, cfg.train_backbone = True).
#include <cfloat>
#include <cstdlib>

#include <algorithm>

#include "din_common.cuh"
#include "din_head.cuh"

namespace {

using namespace din;

// ================================================================================================
// RoIAlign backward: dfm[img, corner, :] += w_corner * dcrop[m, bin, :]   (fp32 atomics; dfm zero-filled by caller)
// One warp per (box, bin), lanes sweep channels -- the forward's mapping.
// ================================================================================================
__global__ void __launch_bounds__(256)
roi_align_bwd_kernel(const float* __restrict__ dcrops, const float* __restrict__ boxes,
                     const int* __restrict__ box_ind, float* __restrict__ dfm, int n_img, int H, int W, int D,
                     int fm_c_stride, int M, int crop_h, int crop_w) {
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int bins = crop_h * crop_w;
  if (warp_global >= M * bins) return;
  const int m = warp_global / bins;
  const int bin = warp_global - m * bins;
  const int iy = bin / crop_w, ix = bin - iy * crop_w;
  const RoiSample sp = roi_sample_point(boxes, box_ind, m, iy, ix, n_img, H, W, crop_h, crop_w);
  if (!sp.ok) return;                                   // extrapolated sample: constant 0, no gradient
  // forward: top = TL + (TR - TL) xl; bot = BL + (BR - BL) xl; out = top + (bot - top) yl
  const float wtl = (1.0f - sp.xl) * (1.0f - sp.yl), wtr = sp.xl * (1.0f - sp.yl);
  const float wbl = (1.0f - sp.xl) * sp.yl, wbr = sp.xl * sp.yl;
  float* base = dfm + static_cast<size_t>(sp.img) * H * W * fm_c_stride;
  float* ptl = base + (static_cast<size_t>(sp.top) * W + sp.left) * fm_c_stride;
  float* ptr = base + (static_cast<size_t>(sp.top) * W + sp.right) * fm_c_stride;
  float* pbl = base + (static_cast<size_t>(sp.bot) * W + sp.left) * fm_c_stride;
  float* pbr = base + (static_cast<size_t>(sp.bot) * W + sp.right) * fm_c_stride;
  const float* g = dcrops + (static_cast<size_t>(m) * bins + bin) * D;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g + c));
    auto add = [&](float* p, float wgt) {
      if (wgt != 0.0f) atomicAdd(reinterpret_cast<float4*>(p + c), make_float4(v.x * wgt, v.y * wgt, v.z * wgt, v.w * wgt));
    };
    add(ptl, wtl); add(ptr, wtr); add(pbl, wbl); add(pbr, wbr);
  }
}

// ================================================================================================
// fp32 gradient -> fp16 with a dynamic power-of-two scale S = 2^floor(log2(target / max|x|)):
//   scale_ws[0] = max|x| (bit pattern, reduced with atomicMax), [1] = S, [2] = 1/S.
// The backbone's backward runs on fp16 tensor-core operands; S keeps the gradient inside fp16's normal range
// and is undone exactly (power of two) in the weight-gradient epilogues.
// ================================================================================================
__global__ void __launch_bounds__(256)
amax_kernel(const float* __restrict__ x, long long count, unsigned int* __restrict__ amax_bits) {
  float m = 0.0f;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < count;
       i += static_cast<long long>(gridDim.x) * 256)
    m = fmaxf(m, fabsf(__ldg(x + i)));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(amax_bits, __float_as_uint(m));   // non-negative floats order as uints
}

__global__ void __launch_bounds__(256)
scale_to_f16_kernel(const float* __restrict__ x, __half* __restrict__ y, long long count, float* __restrict__ scale_ws,
                    float target) {
  const float amax = __uint_as_float(*reinterpret_cast<const unsigned int*>(scale_ws));
  float s = 1.0f;
  if (amax > 0.0f && isfinite(amax)) {
    int e;
    frexpf(target / amax, &e);                            // target/amax = f * 2^e, f in [0.5, 1)
    e = max(-24, min(24, e - 1));
    s = ldexpf(1.0f, e);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { scale_ws[1] = s; scale_ws[2] = 1.0f / s; }
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < count;
       i += static_cast<long long>(gridDim.x) * 256)
    y[i] = __float2half_rn(__ldg(x + i) * s);
}

// ================================================================================================
// ReLU (+ MaxPool2d(2,2)) backward.  y = relu(conv) is the saved pre-pool activation.
//   no pool:  dz = dy * [y > 0]
//   pool   :  dz[yy,xx] = dp[yy/2, xx/2] if (yy,xx) is the FIRST maximum of its 2x2 window (scan order, as
//             torch's max_pool2d) and y > 0, else 0; rows / columns not covered by a window get 0.
// One thread per 8 channels of one full-resolution pixel.
// ================================================================================================
// All arithmetic on packed half2 (ncu on the first version: SM 86 % busy converting to fp32, DRAM 33 %).
// One thread per element and a huge grid: two persistent grid-stride variants that also accumulated the bias
// gradient on the way (1 and 4 elements per iteration) were measured 1.8x and 2x SLOWER than this kernel plus the
// separate column-sum pass (their loads end up in per-element branches and serialise); dropped.
__device__ __forceinline__ __half2 h2_gt(__half2 a, __half2 b) { return __hgt2(a, b); }   // 1.0 / 0.0 per lane

__global__ void __launch_bounds__(256)
relu_pool_bwd_kernel(const __half* __restrict__ y, const __half* __restrict__ dy, __half* __restrict__ dz, int n, int h,
                     int w, int c, int pool) {
  const int c8 = c >> 3;
  const long long total = static_cast<long long>(n) * h * w * c8;
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= total) return;
  const int oc = static_cast<int>(idx % c8);
  long long pix = idx / c8;
  const int xx = static_cast<int>(pix % w);
  pix /= w;
  const int yy = static_cast<int>(pix % h);
  const int img = static_cast<int>(pix / h);
  const uint4* y4 = reinterpret_cast<const uint4*>(y);
  const uint4 mine = __ldg(y4 + idx);
  const __half2* hm = reinterpret_cast<const __half2*>(&mine);
  const __half2 zero = __float2half2_rn(0.0f);
  uint4 out = make_uint4(0, 0, 0, 0);
  __half2* ho = reinterpret_cast<__half2*>(&out);
  if (!pool) {
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(dy) + idx);
    const __half2* hg = reinterpret_cast<const __half2*>(&g);
#pragma unroll
    for (int e = 0; e < 4; ++e) ho[e] = __hmul2(hg[e], h2_gt(hm[e], zero));
  } else {
    const int ph = h >> 1, pw = w >> 1;
    const int py = yy >> 1, px = xx >> 1;
    if (py < ph && px < pw) {
      const uint4 g = __ldg(reinterpret_cast<const uint4*>(dy) + ((static_cast<long long>(img) * ph + py) * pw + px) * c8 + oc);
      const __half2* hg = reinterpret_cast<const __half2*>(&g);
      const int my_pos = ((yy & 1) << 1) | (xx & 1);      // scan order inside the window
      uint4 win[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        win[q] = __ldg(y4 + ((static_cast<long long>(img) * h + (2 * py + (q >> 1))) * w + (2 * px + (q & 1))) * c8 + oc);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // first maximum in scan order: strictly greater than every earlier value, >= every later one, and > 0
        __half2 sel = h2_gt(hm[e], zero);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __half2 o = reinterpret_cast<const __half2*>(&win[q])[e];
          if (q < my_pos) sel = __hmul2(sel, __hgt2(hm[e], o));
          else if (q > my_pos) sel = __hmul2(sel, __hge2(hm[e], o));
        }
        ho[e] = __hmul2(hg[e], sel);
      }
    }
  }
  reinterpret_cast<uint4*>(dz)[idx] = out;
}

// The pooled case, one thread per (2x2 window, 8 channels): the four activations of a window are loaded ONCE (the
// pixel-per-thread form above loads them in each of the window's four threads: 6 loads per output vector, L1-bound at
// 3.2 TB/s of algorithmic traffic), the first maximum in scan order (torch's max_pool2d index rule) takes the pooled
// gradient if it is positive.  Windows run over ceil(h/2) x ceil(w/2): the odd last row / column lies outside every
// pooling window and receives zeros.  Same values as relu_pool_bwd_kernel, bit for bit.
__global__ void __launch_bounds__(256)
relu_pool2_bwd_window_kernel(const __half* __restrict__ y, const __half* __restrict__ dy, __half* __restrict__ dz, int n, int h,
                             int w, int c8) {
  const int wh = (h + 1) >> 1, ww = (w + 1) >> 1, ph = h >> 1, pw = w >> 1;
  const long long total = static_cast<long long>(n) * wh * ww * c8;
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= total) return;
  const int oc = static_cast<int>(idx % c8);
  long long t = idx / c8;
  const int px = static_cast<int>(t % ww);
  t /= ww;
  const int py = static_cast<int>(t % wh);
  const int img = static_cast<int>(t / wh);
  const uint4* y4 = reinterpret_cast<const uint4*>(y);
  uint4* dz4 = reinterpret_cast<uint4*>(dz);
  const long long base = ((static_cast<long long>(img) * h + 2 * py) * w + 2 * px) * c8 + oc;
  const bool col1 = 2 * px + 1 < w, row1 = 2 * py + 1 < h;
  if (py >= ph || px >= pw) {                                  // outside every pooling window
    const uint4 z = make_uint4(0, 0, 0, 0);
    dz4[base] = z;
    if (col1) dz4[base + c8] = z;
    if (row1) dz4[base + static_cast<long long>(w) * c8] = z;
    return;
  }
  const long long off[4] = {0, c8, static_cast<long long>(w) * c8, static_cast<long long>(w) * c8 + c8};
  uint4 win[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) win[q] = __ldg(y4 + base + off[q]);
  const uint4 g = __ldg(reinterpret_cast<const uint4*>(dy) + ((static_cast<long long>(img) * ph + py) * pw + px) * c8 + oc);
  const __half2* hg = reinterpret_cast<const __half2*>(&g);
  const __half2 zero = __float2half2_rn(0.0f);
  uint4 out[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    __half2* ho = reinterpret_cast<__half2*>(&out[q]);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __half2 me = reinterpret_cast<const __half2*>(&win[q])[e];
      __half2 sel = __hgt2(me, zero);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const __half2 o = reinterpret_cast<const __half2*>(&win[r])[e];
        if (r < q) sel = __hmul2(sel, __hgt2(me, o));        // strictly greater than every earlier value
        else if (r > q) sel = __hmul2(sel, __hge2(me, o));   // >= every later one: the first maximum wins
      }
      ho[e] = __hmul2(hg[e], sel);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) dz4[base + off[q]] = out[q];
}

// ================================================================================================
// Stem weight / bias gradient (3 input channels, 3x3 stride 1):
//   dW[co][c][ky][kx] += inv_scale * sum_pixels dz[p][co] * prep(x)[c][p + tap],   db[co] += inv_scale * sum dz
// Persistent CTAs walk strips of 128 output pixels; thread (co, kgroup) keeps 7 of the 28 (27 taps + bias)
// partial sums of its output channel in registers over ALL its strips and flushes them with one atomic each.
// ================================================================================================
constexpr int kSwThreads = 256;

template <bool U8>
__global__ void __launch_bounds__(kSwThreads)
stem_wgrad_kernel(const void* __restrict__ x, const __half* __restrict__ dz, float* __restrict__ dw,
                  float* __restrict__ db, const float* __restrict__ inv_scale, int n, int h, int w, int prep) {
  __shared__ float patch[9][132];                 // [c*3 + ky][px], prepped, zero padded
  __shared__ float dzs[128][64 + 1];
  const int strips_per_row = (w + 127) / 128;
  const long long num_strips = static_cast<long long>(n) * h * strips_per_row;
  const int co = threadIdx.x & 63, kg = threadIdx.x >> 6;       // kg: taps 7*kg .. 7*kg+6 (tap 27 = bias)
  float acc[7] = {0, 0, 0, 0, 0, 0, 0};
  const size_t plane = static_cast<size_t>(h) * w;
  for (long long s = blockIdx.x; s < num_strips; s += gridDim.x) {
    const int strip = static_cast<int>(s % strips_per_row);
    const long long row = s / strips_per_row;
    const int oy = static_cast<int>(row % h);
    const int img = static_cast<int>(row / h);
    const int x0 = strip * 128;
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * 130; i += kSwThreads) {
      const int r = i / 130, px = i - r * 130;
      const int ch = r / 3, ky = r - ch * 3;
      const int gy = oy + ky - 1, gx = x0 + px - 1;
      float v = 0.0f;
      if (gy >= 0 && gy < h && gx >= 0 && gx < w) {
        if (U8) v = static_cast<float>(__ldg(static_cast<const uint8_t*>(x) + (static_cast<size_t>(img) * plane + static_cast<size_t>(gy) * w + gx) * 3 + ch));
        else v = __ldg(static_cast<const float*>(x) + (static_cast<size_t>(img) * 3 + ch) * plane + static_cast<size_t>(gy) * w + gx);
        if (prep) v = __fmul_rn(__fmaf_rn(v, 1.0f / 255.0f, -0.5f), 2.0f);
      }
      patch[r][px] = v;
    }
    const __half* dzr = dz + (static_cast<size_t>(img) * plane + static_cast<size_t>(oy) * w + x0) * 64;
    for (int i = threadIdx.x; i < 128 * 8; i += kSwThreads) {
      const int px = i >> 3, o8 = i & 7;
      uint4 u = make_uint4(0, 0, 0, 0);
      if (x0 + px < w) u = __ldg(reinterpret_cast<const uint4*>(dzr + static_cast<size_t>(px) * 64) + o8);
      const __half* hh = reinterpret_cast<const __half*>(&u);
#pragma unroll
      for (int e = 0; e < 8; ++e) dzs[px][o8 * 8 + e] = __half2float(hh[e]);
    }
    __syncthreads();
#pragma unroll 4
    for (int px = 0; px < 128; ++px) {
      const float g = dzs[px][co];
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int k = kg * 7 + j;                  // k = (c*3 + ky)*3 + kx  (OIHW flatten), 27 = bias
        const float xv = (k < 27) ? patch[k / 3][px + (k % 3)] : 1.0f;
        acc[j] = fmaf(g, xv, acc[j]);
      }
    }
  }
  const float scl = inv_scale ? __ldg(inv_scale) : 1.0f;
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const int k = kg * 7 + j;
    if (k < 27) atomicAdd(dw + co * 27 + k, acc[j] * scl);
    else if (db != nullptr) atomicAdd(db + co, acc[j] * scl);
  }
}

// ================================================================================================
// ResNet-18 pieces (stride-2 convolutions, 3x3 stride-2 max-pool, eval-mode BatchNorm folded into the convolutions)
// ================================================================================================
// dst[n, 2*oy, 2*ox, :] (+)= src[n, oy, ox, :]; with accumulate == 0 every other dst pixel is zero-filled.
//   accumulate == 0: "zero insertion": dZ of a stride-2 convolution laid out on the input grid, so that BOTH its data
//                    gradient and its weight gradient are the stride-1 kernels' (dX = conv(dZ_up, rot W), dW = wgrad(X,
//                    dZ_up)): 4x the useful work on the three stride-2 3x3 layers, no new tensor-core code;
//   accumulate == 1: the 1x1 stride-2 shortcut's data gradient added at the even positions.
__global__ void __launch_bounds__(256)
scatter2_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int n, int h, int w, int c8, int oh, int ow,
                int accumulate) {
  const long long total = static_cast<long long>(n) * h * w * c8;
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= total) return;
  const int oc = static_cast<int>(idx % c8);
  long long pix = idx / c8;
  const int ix = static_cast<int>(pix % w);
  pix /= w;
  const int iy = static_cast<int>(pix % h);
  const int img = static_cast<int>(pix / h);
  const bool hit = !(iy & 1) && !(ix & 1) && (iy >> 1) < oh && (ix >> 1) < ow;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (hit) v = __ldg(reinterpret_cast<const uint4*>(src) + ((static_cast<long long>(img) * oh + (iy >> 1)) * ow + (ix >> 1)) * c8 + oc);
  uint4* d = reinterpret_cast<uint4*>(dst) + idx;
  if (accumulate) {
    if (!hit) return;
    uint4 cur = *d;
    __half2* a = reinterpret_cast<__half2*>(&cur);
    const __half2* b = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int e = 0; e < 4; ++e) a[e] = __hadd2(a[e], b[e]);
    *d = cur;
  } else {
    *d = v;
  }
}

// y = a + b (fp16, the residual branch merging two gradients)
__global__ void __launch_bounds__(256)
add_f16_kernel(const __half* __restrict__ a, const __half* __restrict__ b, __half* __restrict__ y, long long count8) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= count8) return;
  const uint4 va = __ldg(reinterpret_cast<const uint4*>(a) + i), vb = __ldg(reinterpret_cast<const uint4*>(b) + i);
  uint4 o;
  const __half2* ha = reinterpret_cast<const __half2*>(&va);
  const __half2* hb = reinterpret_cast<const __half2*>(&vb);
  __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) ho[e] = __hadd2(ha[e], hb[e]);
  reinterpret_cast<uint4*>(y)[i] = o;
}

// dz[r, 0:c] = dy[r, 0:c] * [y[r, 0:c] > 0] over channel slices of three buffers with their own channel strides (in
// 8-channel units): the ReLU backward of an Inception branch whose output is a slice of the block's concat buffer.
__global__ void __launch_bounds__(256)
relu_bwd_slice_kernel(const __half* __restrict__ y, const __half* __restrict__ dy, __half* __restrict__ dz, long long rows,
                      int c8, int ys8, int dys8, int dzs8) {
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= rows * c8) return;
  const int oc = static_cast<int>(idx % c8);
  const long long r = idx / c8;
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(y) + r * ys8 + oc);
  const uint4 g = __ldg(reinterpret_cast<const uint4*>(dy) + r * dys8 + oc);
  const __half2* ha = reinterpret_cast<const __half2*>(&a);
  const __half2* hg = reinterpret_cast<const __half2*>(&g);
  const __half2 zero = __float2half2_rn(0.0f);
  uint4 o;
  __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) ho[e] = __hmul2(hg[e], __hgt2(ha[e], zero));
  reinterpret_cast<uint4*>(dz)[r * dzs8 + oc] = o;
}

// Backward of the align_corners=True bilinear resize [n,h,w,c] -> [n,oh,ow,c] (F.interpolate at infer_model.py:169), gather
// form: source pixel (sy, sx) collects dy from every output pixel whose forward interpolation read it, with the forward's
// weights (same fp32 arithmetic: scale = (h-1)/(oh-1), y0 = int(scale * oy), ly = scale * oy - y0).  One thread per
// (8 channels, source pixel); candidates are the outputs oy with y0(oy) in {sy - 1, sy}: a window of 1/scale + 2 rows.
__global__ void __launch_bounds__(256)
upsample_bilinear_bwd_kernel(const __half* __restrict__ dy, __half* __restrict__ dx, int n, int h, int w, int c8, int dys8,
                             int dxs8, int oh, int ow) {
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= static_cast<long long>(n) * h * w * c8) return;
  const int oc = static_cast<int>(idx % c8);
  long long t = idx / c8;
  const int sx = static_cast<int>(t % w);
  t /= w;
  const int sy = static_cast<int>(t % h);
  const int img = static_cast<int>(t / h);
  const float scy = oh > 1 ? static_cast<float>(h - 1) / static_cast<float>(oh - 1) : 0.0f;
  const float scx = ow > 1 ? static_cast<float>(w - 1) / static_cast<float>(ow - 1) : 0.0f;
  // outputs that can touch source row sy: scy * oy in (sy - 1, sy + 1)
  int oy_lo = 0, oy_hi = oh - 1, ox_lo = 0, ox_hi = ow - 1;
  if (scy > 0.0f) {
    oy_lo = max(0, static_cast<int>(floorf(static_cast<float>(sy - 1) / scy)) - 1);
    oy_hi = min(oh - 1, static_cast<int>(ceilf(static_cast<float>(sy + 1) / scy)) + 1);
  }
  if (scx > 0.0f) {
    ox_lo = max(0, static_cast<int>(floorf(static_cast<float>(sx - 1) / scx)) - 1);
    ox_hi = min(ow - 1, static_cast<int>(ceilf(static_cast<float>(sx + 1) / scx)) + 1);
  }
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const uint4* g4 = reinterpret_cast<const uint4*>(dy) + static_cast<long long>(img) * oh * ow * dys8 + oc;
  for (int oy = oy_lo; oy <= oy_hi; ++oy) {
    const float fy = scy * static_cast<float>(oy);
    const int y0 = min(static_cast<int>(fy), h - 1), y1 = min(y0 + 1, h - 1);
    const float ly = fy - static_cast<float>(y0);
    const float wy = (y0 == sy ? 1.0f - ly : 0.0f) + (y1 == sy ? ly : 0.0f);
    if (wy == 0.0f) continue;
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      const float fx = scx * static_cast<float>(ox);
      const int x0 = min(static_cast<int>(fx), w - 1), x1 = min(x0 + 1, w - 1);
      const float lx = fx - static_cast<float>(x0);
      const float wx = (x0 == sx ? 1.0f - lx : 0.0f) + (x1 == sx ? lx : 0.0f);
      if (wx == 0.0f) continue;
      const uint4 g = __ldg(g4 + (static_cast<long long>(oy) * ow + ox) * dys8);
      const __half2* hg = reinterpret_cast<const __half2*>(&g);
      const float wgt = wy * wx;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(hg[e]);
        acc[2 * e] += wgt * f.x;
        acc[2 * e + 1] += wgt * f.y;
      }
    }
  }
  uint4 o;
  __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e) ho[e] = __floats2half2_rn(acc[2 * e], acc[2 * e + 1]);
  reinterpret_cast<uint4*>(dx)[((static_cast<long long>(img) * h + sy) * w + sx) * dxs8 + oc] = o;
}

// Backward of MaxPool2d(3, 2, 1) fused with the ReLU before it (resnet18.relu / .maxpool):
//   dz[iy,ix] = [x > 0] * sum over the (1, 2 or 4) windows containing (iy,ix) of  dy[window] * [(iy,ix) is the window's
//   FIRST maximum in scan order]   (padding is -inf and never wins, as torch).
// A CTA owns an 8 x 32 patch of input pixels.  Phase 1: the 5 x 17 windows touching the patch each find their first
// maximum once (9 loads per window = 3 per owned pixel; fp16x2 compare/select, the scan position kept as an fp16
// number) into shared memory.  Phase 2: every owned pixel compares its position with its windows' winners and gathers
// dy.  (The direct per-pixel gather re-scanned up to 4 windows = 36 loads per pixel: 2.2 ms per 20 720p frames; this
// form: see profiles/.)
constexpr int kMpTileH = 8, kMpTileW = 32;
constexpr int kMpWinH = kMpTileH / 2 + 1, kMpWinW = kMpTileW / 2 + 1;

// PAD = 1: resnet18.maxpool; PAD = 0: the F.max_pool2d(3, 2) of the Inception-v3 trunk (backbone.py:50,56) and of
// Mixed_6a's pool branch.  x / dy / dz may be channel slices of wider buffers (xs8 / dys8 / dzs8 = their channel strides
// in 8-channel units).
template <int PAD>
__global__ void __launch_bounds__(256)
maxpool3s2_relu_bwd_kernel(const __half* __restrict__ x, const __half* __restrict__ dy, __half* __restrict__ dz, int h,
                           int w, int c8, int oh, int ow, int tiles_x, int tiles_y, int xs8, int dys8, int dzs8,
                           int accumulate) {
  extern __shared__ uint4 win_idx[];                     // [kMpWinH * kMpWinW][c8] x 8 channels (fp16 scan position)
  const int tx = blockIdx.x % tiles_x, ty = (blockIdx.x / tiles_x) % tiles_y, img = blockIdx.x / (tiles_x * tiles_y);
  const int iy0 = ty * kMpTileH, ix0 = tx * kMpTileW;
  const int oy0 = (iy0 + PAD - 1) >> 1, ox0 = (ix0 + PAD - 1) >> 1;      // first window touching the patch (may be -1)
  const uint4* x4 = reinterpret_cast<const uint4*>(x) + static_cast<long long>(img) * h * w * xs8;
  const uint4* dy4 = reinterpret_cast<const uint4*>(dy) + static_cast<long long>(img) * oh * ow * dys8;
  uint4* dz4 = reinterpret_cast<uint4*>(dz) + static_cast<long long>(img) * h * w * dzs8;
  const __half2 zero = __float2half2_rn(0.0f);
  const __half2 ninf = __half2half2(__ushort_as_half(static_cast<unsigned short>(0xFC00)));

  for (int item = threadIdx.x; item < kMpWinH * kMpWinW * c8; item += 256) {
    const int oc = item % c8, wi = item / c8;
    const int oy = oy0 + wi / kMpWinW, ox = ox0 + wi % kMpWinW;
    uint4 res;
    __half2* idx = reinterpret_cast<__half2*>(&res);
    if (oy < 0 || ox < 0 || oy >= oh || ox >= ow) {
      idx[0] = idx[1] = idx[2] = idx[3] = __float2half2_rn(-1.0f);        // no such window: matches no position
    } else {
      uint4 v[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) {                                        // all nine loads in flight, addresses clamped
        const int yy = min(max(2 * oy - PAD + q / 3, 0), h - 1), xx = min(max(2 * ox - PAD + q % 3, 0), w - 1);
        v[q] = __ldg(x4 + (static_cast<long long>(yy) * w + xx) * xs8 + oc);
      }
      __half2 best[4] = {ninf, ninf, ninf, ninf};
      idx[0] = idx[1] = idx[2] = idx[3] = zero;
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        const int yy = 2 * oy - PAD + q / 3, xx = 2 * ox - PAD + q % 3;
        const bool ok = yy >= 0 && yy < h && xx >= 0 && xx < w;
        const __half2* hv = reinterpret_cast<const __half2*>(&v[q]);
        const __half2 q2 = __float2half2_rn(static_cast<float>(q));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __half2 val = ok ? hv[e] : ninf;
          const __half2 gt = __hgt2(val, best[e]);                         // strictly greater: the first maximum stays
          best[e] = __hmax2(best[e], val);
          idx[e] = __hfma2(gt, __hsub2(q2, idx[e]), idx[e]);
        }
      }
    }
    win_idx[item] = res;
  }
  __syncthreads();

  for (int item = threadIdx.x; item < kMpTileH * kMpTileW * c8; item += 256) {
    const int oc = item % c8, pi = item / c8;
    const int iy = iy0 + pi / kMpTileW, ix = ix0 + pi % kMpTileW;
    if (iy >= h || ix >= w) continue;
    const uint4 mine = __ldg(x4 + (static_cast<long long>(iy) * w + ix) * xs8 + oc);
    const __half2* hm = reinterpret_cast<const __half2*>(&mine);
    __half2 acc[4] = {zero, zero, zero, zero};
    for (int oy = (iy + PAD - 1) >> 1; oy <= ((iy + PAD) >> 1); ++oy) {    // windows with 2*oy - PAD <= iy <= 2*oy - PAD + 2
      if (oy < 0 || oy >= oh) continue;
      for (int ox = (ix + PAD - 1) >> 1; ox <= ((ix + PAD) >> 1); ++ox) {
        if (ox < 0 || ox >= ow) continue;
        const int my_pos = (iy - (2 * oy - PAD)) * 3 + (ix - (2 * ox - PAD));
        const __half2 p2 = __float2half2_rn(static_cast<float>(my_pos));
        const uint4 wv = win_idx[((oy - oy0) * kMpWinW + (ox - ox0)) * c8 + oc];
        const uint4 g = __ldg(dy4 + (static_cast<long long>(oy) * ow + ox) * dys8 + oc);
        const __half2* hw2 = reinterpret_cast<const __half2*>(&wv);
        const __half2* hg = reinterpret_cast<const __half2*>(&g);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = __hfma2(hg[e], __heq2(hw2[e], p2), acc[e]);
      }
    }
    uint4 out;
    __half2* ho = reinterpret_cast<__half2*>(&out);
#pragma unroll
    for (int e = 0; e < 4; ++e) ho[e] = __hmul2(acc[e], __hgt2(hm[e], zero));
    uint4* dst = dz4 + (static_cast<long long>(iy) * w + ix) * dzs8 + oc;
    if (accumulate) {                                   // dz already holds another branch's gradient of the same tensor
      const uint4 prev = *dst;
      const __half2* hp = reinterpret_cast<const __half2*>(&prev);
#pragma unroll
      for (int e = 0; e < 4; ++e) ho[e] = __hadd2(ho[e], hp[e]);
    }
    *dst = out;
  }
}

// Eval-mode BatchNorm folded into a convolution: z = gamma * xhat + beta.  d(gamma)[c] += inv_scale * sum_p dz * xhat with
// xhat = (z - beta) / gamma recovered from the saved activation (z = zsrc - sub; wherever the ReLU zeroed z, dz is zero
// too, so the post-ReLU tensor serves).  d(beta) is the convolution's bias gradient (conv_wgrad's dbias).
__global__ void __launch_bounds__(256)
bn_gamma_grad_kernel(const __half* __restrict__ dz, const __half* __restrict__ zsrc, const __half* __restrict__ sub,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ dgamma,
                     long long rows, int c, const float* __restrict__ inv_scale) {
  __shared__ float red[256][8 + 1];
  const int octs = c >> 3;
  const int lanes_r = 256 / octs;
  const int oct = threadIdx.x % octs, rl = threadIdx.x / octs;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (rl < lanes_r) {
    float be[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) be[e] = __ldg(beta + oct * 8 + e);
    // two rows in flight per thread (all loads issued before the arithmetic): the kernel is latency-bound otherwise
    const long long step = static_cast<long long>(gridDim.x) * lanes_r;
    for (long long r = static_cast<long long>(blockIdx.x) * lanes_r + rl; r < rows; r += 2 * step) {
      const long long r2 = r + step;
      const bool two = r2 < rows;
      const long long rb = two ? r2 : r;
      const uint4 g = __ldg(reinterpret_cast<const uint4*>(dz + r * c) + oct);
      const uint4 z = __ldg(reinterpret_cast<const uint4*>(zsrc + r * c) + oct);
      const uint4 g2 = __ldg(reinterpret_cast<const uint4*>(dz + rb * c) + oct);
      const uint4 z2 = __ldg(reinterpret_cast<const uint4*>(zsrc + rb * c) + oct);
      uint4 sb = make_uint4(0, 0, 0, 0), sb2 = make_uint4(0, 0, 0, 0);
      if (sub != nullptr) {
        sb = __ldg(reinterpret_cast<const uint4*>(sub + r * c) + oct);
        sb2 = __ldg(reinterpret_cast<const uint4*>(sub + rb * c) + oct);
      }
      const __half* hg = reinterpret_cast<const __half*>(&g);
      const __half* hz = reinterpret_cast<const __half*>(&z);
      const __half* hs = reinterpret_cast<const __half*>(&sb);
      const __half* hg2 = reinterpret_cast<const __half*>(&g2);
      const __half* hz2 = reinterpret_cast<const __half*>(&z2);
      const __half* hs2 = reinterpret_cast<const __half*>(&sb2);
      const float w2 = two ? 1.0f : 0.0f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        acc[e] = fmaf(__half2float(hg[e]), (__half2float(hz[e]) - __half2float(hs[e])) - be[e], acc[e]);
        acc[e] = fmaf(w2 * __half2float(hg2[e]), (__half2float(hz2[e]) - __half2float(hs2[e])) - be[e], acc[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[threadIdx.x][e] = acc[e];
  __syncthreads();
  if (threadIdx.x < octs) {
    const float scl = inv_scale ? __ldg(inv_scale) : 1.0f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float s = 0.0f;
      for (int l = 0; l < lanes_r; ++l) s += red[l * octs + threadIdx.x][e];
      const float g = __ldg(gamma + threadIdx.x * 8 + e);
      if (g != 0.0f) atomicAdd(dgamma + threadIdx.x * 8 + e, s * scl / g);   // gamma == 0: xhat is not recoverable
    }
  }
}

// d(gamma) of an eval-mode BatchNorm folded into its convolution, WITHOUT touching the activations: with c = conv(x, W) the
// un-normalised output, z = gamma * (c - mean) * invstd + beta, so
//   d(gamma) = invstd * (sum_p dz * c - mean * sum_p dz) = invstd * (<W, dWf> - mean * d(beta))
// because sum_p dz[p] * c[p] = sum_k W[k] * (sum_p dz[p] * x[p + k]) = <W, dWf>, dWf being the gradient w.r.t. the FOLDED
// weight that the wgrad kernel already produced.  Exact in fp32, no division by gamma (bn_gamma_grad_kernel recovers
// xhat as (z - beta) / gamma from fp16 activations: fine for random init, noise-amplifying for the near-zero gammas a
// trained checkpoint holds), and no pass over HBM-sized tensors.  One block per output channel.
__global__ void __launch_bounds__(128)
bn_fold_grads_kernel(const float* __restrict__ w, const float* __restrict__ dwf, const float* __restrict__ dbeta,
                     const float* __restrict__ running_mean, const float* __restrict__ running_var, float eps,
                     float* __restrict__ dgamma, long long cols) {
  __shared__ float red[33];
  const long long row = blockIdx.x;
  const float* wr = w + row * cols;
  const float* dr = dwf + row * cols;
  float acc = 0.0f;
  for (long long i = threadIdx.x; i < cols; i += 128) acc = fmaf(__ldg(wr + i), __ldg(dr + i), acc);
  const float dot = block_sum<128>(acc, red);
  if (threadIdx.x == 0)
    dgamma[row] = rsqrtf(__ldg(running_var + row) + eps) * (dot - __ldg(running_mean + row) * __ldg(dbeta + row));
}

// w[r][:] *= scale[r]   (gradient of the un-folded convolution weight: dW = dW_folded * gamma / sqrt(var + eps))
__global__ void __launch_bounds__(256)
scale_rows_kernel(float* __restrict__ w, const float* __restrict__ scale, long long rows, long long cols) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= rows * cols) return;
  w[i] *= __ldg(scale + i / cols);
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" int din_roi_align_bwd_f32(const float* dcrops, const float* boxes, const int32_t* box_ind, float* dfm,
                                     int n_img, int h, int w, int d, int fm_c_stride, int m, int crop_h, int crop_w,
                                     void* stream) {
  DIN_CHECK_ARG(dcrops && boxes && box_ind && dfm, "din_roi_align_bwd_f32: null pointer");
  DIN_CHECK_ARG(n_img > 0 && h > 1 && w > 1 && m > 0, "din_roi_align_bwd_f32: bad extent n=%d h=%d w=%d m=%d", n_img, h,
                w, m);
  DIN_CHECK_ARG(d > 0 && d % 4 == 0 && fm_c_stride >= d && fm_c_stride % 4 == 0,
                "din_roi_align_bwd_f32: d=%d / fm_c_stride=%d must be multiples of 4", d, fm_c_stride);
  DIN_CHECK_ARG(crop_h > 1 && crop_w > 1, "din_roi_align_bwd_f32: crop must be > 1");
  DIN_CHECK_ARG(((reinterpret_cast<uintptr_t>(dcrops) | reinterpret_cast<uintptr_t>(dfm)) & 15) == 0,
                "din_roi_align_bwd_f32: pointers must be 16-byte aligned");
  const long long warps = static_cast<long long>(m) * crop_h * crop_w;
  roi_align_bwd_kernel<<<static_cast<int>((warps + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dcrops, boxes, box_ind, dfm, n_img, h, w, d, fm_c_stride, m, crop_h, crop_w);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_grad_to_f16(const float* x, void* y, float* scale_ws, long long count, float target, void* stream) {
  DIN_CHECK_ARG(x && y && scale_ws, "din_grad_to_f16: null pointer");
  DIN_CHECK_ARG(count > 0 && target > 0.0f, "din_grad_to_f16: bad count %lld / target %g", count, target);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DIN_CHECK_CUDA(cudaMemsetAsync(scale_ws, 0, 4 * sizeof(float), st));
  const int sms = din_num_sms();
  long long blocks = (count + 255) / 256;
  const long long cap = static_cast<long long>(sms > 0 ? sms : 148) * 8;
  if (blocks > cap) blocks = cap;
  amax_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(x, count, reinterpret_cast<unsigned int*>(scale_ws));
  DIN_CHECK_CUDA(cudaGetLastError());
  scale_to_f16_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(x, static_cast<__half*>(y), count, scale_ws, target);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_relu_pool_bwd_nhwc_f16(const void* y, const void* dy, void* dz, int n, int h, int w, int c, int pool,
                                          void* stream) {
  DIN_CHECK_ARG(y && dy && dz, "din_relu_pool_bwd_nhwc_f16: null pointer");
  DIN_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "din_relu_pool_bwd_nhwc_f16: bad shape n=%d h=%d w=%d c=%d",
                n, h, w, c);
  DIN_CHECK_ARG(!pool || (h >= 2 && w >= 2), "din_relu_pool_bwd_nhwc_f16: pooling needs at least 2x2");
  DIN_CHECK_ARG(((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dz)) & 15) == 0,
                "din_relu_pool_bwd_nhwc_f16: pointers must be 16-byte aligned");
  const long long total = static_cast<long long>(n) * h * w * (c / 8);
  DIN_CHECK_ARG((total + 255) / 256 <= INT32_MAX, "din_relu_pool_bwd_nhwc_f16: too large");
  {
    const char* e = std::getenv("DIN_RELU_POOL_PIXEL");       // =1: the pixel-per-thread kernel for pooled layers too (A/B)
    if (pool && !(e && e[0] == '1')) {
      const long long wins = static_cast<long long>(n) * ((h + 1) / 2) * ((w + 1) / 2) * (c / 8);
      relu_pool2_bwd_window_kernel<<<static_cast<int>((wins + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
          static_cast<const __half*>(y), static_cast<const __half*>(dy), static_cast<__half*>(dz), n, h, w, c / 8);
      DIN_CHECK_CUDA(cudaGetLastError());
      return DIN_OK;
    }
  }
  relu_pool_bwd_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(y), static_cast<const __half*>(dy), static_cast<__half*>(dz), n, h, w, c, pool);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_stem_wgrad(const void* x, int x_is_u8, const void* dz, float* dw, float* dbias,
                              const float* inv_scale, int n, int h, int w, int c_out, int kh, int kw, int stride, int pad,
                              int prep, void* stream) {
  DIN_CHECK_ARG(x && dz && dw, "din_stem_wgrad: null pointer");
  DIN_CHECK_ARG(n > 0 && h > 0 && w > 0, "din_stem_wgrad: bad extent n=%d h=%d w=%d", n, h, w);
  DIN_CHECK_ARG(kh == kw && ((c_out == 64 && kh == 3 && stride == 1 && pad == 1) || (c_out == 64 && kh == 7 && stride == 2 && pad == 3) ||
                             (c_out == 32 && kh == 3 && stride == 2 && pad == 0)),
                "din_stem_wgrad: only the VGG-16 (64 x 3x3 s1 p1), ResNet-18 (64 x 7x7 s2 p3) and Inception-v3 (32 x 3x3 s2 p0) "
                "stems are implemented");
  DIN_CHECK_ARG((reinterpret_cast<uintptr_t>(dz) & 15) == 0, "din_stem_wgrad: dz must be 16-byte aligned");
  {
    // production path: tensor cores (stem_tc.cu).  DIN_STEM_WGRAD_SIMT=1 keeps the CUDA-core kernel below for A/B
    // measurements -- still a kernel of this library.
    const char* e = std::getenv("DIN_STEM_WGRAD_SIMT");
    if (!(e && e[0] == '1') || kh != 3 || stride != 1)
      return din_stem_wgrad_tc_launch(x, x_is_u8, dz, dw, dbias, inv_scale, n, h, w, c_out, kh, stride, pad, prep,
                                      static_cast<cudaStream_t>(stream));
  }
  const int sms = din_num_sms();
  const long long strips = static_cast<long long>(n) * h * ((w + 127) / 128);
  long long grid = static_cast<long long>(sms > 0 ? sms : 148) * 3;
  if (grid > strips) grid = strips;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (x_is_u8)
    stem_wgrad_kernel<true><<<static_cast<int>(grid), kSwThreads, 0, st>>>(x, static_cast<const __half*>(dz), dw, dbias,
                                                                          inv_scale, n, h, w, prep);
  else
    stem_wgrad_kernel<false><<<static_cast<int>(grid), kSwThreads, 0, st>>>(x, static_cast<const __half*>(dz), dw, dbias,
                                                                           inv_scale, n, h, w, prep);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_scatter2_nhwc_f16(const void* src, void* dst, int n, int h, int w, int c, int oh, int ow, int accumulate,
                                     void* stream) {
  DIN_CHECK_ARG(src && dst, "din_scatter2_nhwc_f16: null pointer");
  DIN_CHECK_ARG(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && oh > 0 && ow > 0 && 2 * (oh - 1) < h && 2 * (ow - 1) < w,
                "din_scatter2_nhwc_f16: bad shape n=%d h=%d w=%d c=%d oh=%d ow=%d", n, h, w, c, oh, ow);
  DIN_CHECK_ARG(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0,
                "din_scatter2_nhwc_f16: pointers must be 16-byte aligned");
  const long long total = static_cast<long long>(n) * h * w * (c / 8);
  DIN_CHECK_ARG((total + 255) / 256 <= INT32_MAX, "din_scatter2_nhwc_f16: too large");
  scatter2_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(src), static_cast<__half*>(dst), n, h, w, c / 8, oh, ow, accumulate);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_add_f16(const void* a, const void* b, void* y, long long count, void* stream) {
  DIN_CHECK_ARG(a && b && y, "din_add_f16: null pointer");
  DIN_CHECK_ARG(count > 0 && count % 8 == 0, "din_add_f16: count=%lld must be a positive multiple of 8", count);
  DIN_CHECK_ARG(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15) == 0,
                "din_add_f16: pointers must be 16-byte aligned");
  const long long n8 = count / 8;
  DIN_CHECK_ARG((n8 + 255) / 256 <= INT32_MAX, "din_add_f16: too large");
  add_f16_kernel<<<static_cast<int>((n8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(a), static_cast<const __half*>(b), static_cast<__half*>(y), n8);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_relu_bwd_slice_nhwc_f16(const void* y, const void* dy, void* dz, long long rows, int c, int y_c_stride,
                                          int dy_c_stride, int dz_c_stride, void* stream) {
  const char* who = "din_relu_bwd_slice_nhwc_f16";
  DIN_CHECK_ARG(y && dy && dz, "%s: null pointer", who);
  DIN_CHECK_ARG(rows > 0 && c > 0 && c % 8 == 0 && y_c_stride >= c && dy_c_stride >= c && dz_c_stride >= c &&
                    (y_c_stride | dy_c_stride | dz_c_stride) % 8 == 0, "%s: bad shape rows=%lld c=%d", who, rows, c);
  DIN_CHECK_ARG(((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dz)) & 15) == 0,
                "%s: pointers must be 16-byte aligned", who);
  const long long total = rows * (c / 8);
  DIN_CHECK_ARG((total + 255) / 256 <= INT32_MAX, "%s: too large", who);
  relu_bwd_slice_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(y), static_cast<const __half*>(dy), static_cast<__half*>(dz), rows, c / 8, y_c_stride / 8,
      dy_c_stride / 8, dz_c_stride / 8);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_upsample_bilinear_bwd_nhwc_f16(const void* dy, void* dx, int n, int h, int w, int c, int dy_c_stride,
                                                  int dx_c_stride, int oh, int ow, void* stream) {
  const char* who = "din_upsample_bilinear_bwd_nhwc_f16";
  DIN_CHECK_ARG(dy && dx, "%s: null pointer", who);
  DIN_CHECK_ARG(n > 0 && h > 0 && w > 0 && oh > 0 && ow > 0 && c > 0 && c % 8 == 0 && dy_c_stride >= c && dx_c_stride >= c &&
                    (dy_c_stride | dx_c_stride) % 8 == 0, "%s: bad shape", who);
  DIN_CHECK_ARG(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0,
                "%s: pointers must be 16-byte aligned", who);
  const long long total = static_cast<long long>(n) * h * w * (c / 8);
  DIN_CHECK_ARG((total + 255) / 256 <= INT32_MAX, "%s: too large", who);
  upsample_bilinear_bwd_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(dy), static_cast<__half*>(dx), n, h, w, c / 8, dy_c_stride / 8, dx_c_stride / 8, oh, ow);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_maxpool3s2_bwd_nhwc_f16(const void* x, const void* dy, void* dz, int n, int h, int w, int c,
                                          int x_c_stride, int dy_c_stride, int dz_c_stride, int pad, int accumulate,
                                          void* stream) {
  const char* who = "din_maxpool3s2_bwd_nhwc_f16";
  DIN_CHECK_ARG(x && dy && dz, "%s: null pointer", who);
  DIN_CHECK_ARG(n > 0 && h >= 3 - 2 * pad && w >= 3 - 2 * pad && c > 0 && c % 8 == 0 && (pad == 0 || pad == 1), "%s: bad shape", who);
  DIN_CHECK_ARG(x_c_stride >= c && dy_c_stride >= c && dz_c_stride >= c && (x_c_stride | dy_c_stride | dz_c_stride) % 8 == 0,
                "%s: channel strides must be >= c and multiples of 8", who);
  DIN_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dz)) & 15) == 0,
                "%s: pointers must be 16-byte aligned", who);
  const int oh = (h + 2 * pad - 3) / 2 + 1, ow = (w + 2 * pad - 3) / 2 + 1;
  const int tiles_x = (w + kMpTileW - 1) / kMpTileW, tiles_y = (h + kMpTileH - 1) / kMpTileH;
  const long long ctas = static_cast<long long>(n) * tiles_x * tiles_y;
  const size_t smem = static_cast<size_t>(kMpWinH) * kMpWinW * (c / 8) * sizeof(uint4);
  DIN_CHECK_ARG(ctas <= INT32_MAX && smem <= 48 * 1024, "%s: too large (c <= 288)", who);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (pad == 1)
    maxpool3s2_relu_bwd_kernel<1><<<static_cast<int>(ctas), 256, smem, st>>>(
        static_cast<const __half*>(x), static_cast<const __half*>(dy), static_cast<__half*>(dz), h, w, c / 8, oh, ow, tiles_x,
        tiles_y, x_c_stride / 8, dy_c_stride / 8, dz_c_stride / 8, accumulate);
  else
    maxpool3s2_relu_bwd_kernel<0><<<static_cast<int>(ctas), 256, smem, st>>>(
        static_cast<const __half*>(x), static_cast<const __half*>(dy), static_cast<__half*>(dz), h, w, c / 8, oh, ow, tiles_x,
        tiles_y, x_c_stride / 8, dy_c_stride / 8, dz_c_stride / 8, accumulate);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_maxpool3s2_relu_bwd_nhwc_f16(const void* x, const void* dy, void* dz, int n, int h, int w, int c,
                                                void* stream) {
  return din_maxpool3s2_bwd_nhwc_f16(x, dy, dz, n, h, w, c, c, c, c, 1, 0, stream);
}

extern "C" int din_bn_gamma_grad_f16(const void* dz, const void* zsrc, const void* sub, const float* gamma,
                                     const float* beta, float* dgamma, const float* inv_scale, long long rows, int c,
                                     void* stream) {
  DIN_CHECK_ARG(dz && zsrc && gamma && beta && dgamma, "din_bn_gamma_grad_f16: null pointer");
  DIN_CHECK_ARG(rows > 0 && c > 0 && c % 8 == 0 && c <= 2048, "din_bn_gamma_grad_f16: bad shape rows=%lld c=%d", rows, c);
  const int sms = din_num_sms();
  int g = 8 * (sms > 0 ? sms : 148);
  const long long per = 256 / (c / 8);
  // >= 16 rows per thread: every CTA ends in c atomics on the same c addresses
  if (static_cast<long long>(g) * per * 16 > rows) g = static_cast<int>(std::max<long long>(1, rows / (per * 16)));
  bn_gamma_grad_kernel<<<g, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(dz), static_cast<const __half*>(zsrc), static_cast<const __half*>(sub), gamma, beta,
      dgamma, rows, c, inv_scale);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_bn_fold_grads_f32(const float* w, const float* dwf, const float* dbeta, const float* running_mean,
                                     const float* running_var, float eps, float* dgamma, int rows, long long cols,
                                     void* stream) {
  DIN_CHECK_ARG(w && dwf && dbeta && running_mean && running_var && dgamma, "din_bn_fold_grads_f32: null pointer");
  DIN_CHECK_ARG(rows > 0 && cols > 0, "din_bn_fold_grads_f32: bad shape rows=%d cols=%lld", rows, cols);
  bn_fold_grads_kernel<<<rows, 128, 0, static_cast<cudaStream_t>(stream)>>>(w, dwf, dbeta, running_mean, running_var, eps,
                                                                            dgamma, cols);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_scale_rows_f32(float* w, const float* scale, long long rows, long long cols, void* stream) {
  DIN_CHECK_ARG(w && scale && rows > 0 && cols > 0, "din_scale_rows_f32: bad arguments");
  const long long total = rows * cols;
  DIN_CHECK_ARG((total + 255) / 256 <= INT32_MAX, "din_scale_rows_f32: too large");
  scale_rows_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(w, scale, rows,
                                                                                                         cols);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}
