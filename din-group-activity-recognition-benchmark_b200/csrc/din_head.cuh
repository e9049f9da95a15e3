// din_head.cuh — device helpers shared by the forward (head.cu) and backward (head_bwd.cu) kernels of the
// person-level head.
#pragma once

#include "din_common.cuh"

namespace din {

// ================================================================================================
// block-wide sum helper
// ================================================================================================
template <int kThreads>
__device__ __forceinline__ float block_sum(float v, float* red /* >= 33 floats */) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // protect `red` from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < (kThreads / 32) ? red[lane] : 0.0f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}


// ================================================================================================
// RoIAlign sample point (crop_and_resize with longcw/RoIAlign.pytorch's box transform), shared by the forward
// gather and the backward scatter
// ================================================================================================
struct RoiSample {
  bool ok;                 // sample point inside [0, H-1] x [0, W-1] and a valid frame index
  int img, top, bot, left, right;
  float yl, xl;            // bilinear fractions
};

__device__ __forceinline__ RoiSample roi_sample_point(const float* __restrict__ boxes, const int* __restrict__ box_ind,
                                                      int m, int iy, int ix, int n_img, int H, int W, int crop_h,
                                                      int crop_w) {
  const float x1 = __ldg(boxes + 4 * m + 0), y1 = __ldg(boxes + 4 * m + 1);
  const float x2 = __ldg(boxes + 4 * m + 2), y2 = __ldg(boxes + 4 * m + 3);
  const int b = __ldg(box_ind + m);
  // RoIAlign.forward (transform_fpcoor=True): same fp32 op order as the published implementation, no
  // FMA contraction, so floor()/validity decisions agree with the CPU restatement.
  const float Wm1 = static_cast<float>(W - 1), Hm1 = static_cast<float>(H - 1);
  const float spacing_w = __fdiv_rn(__fsub_rn(x2, x1), static_cast<float>(crop_w));
  const float spacing_h = __fdiv_rn(__fsub_rn(y2, y1), static_cast<float>(crop_h));
  const float nx0 = __fdiv_rn(__fsub_rn(__fadd_rn(x1, __fdiv_rn(spacing_w, 2.0f)), 0.5f), Wm1);
  const float ny0 = __fdiv_rn(__fsub_rn(__fadd_rn(y1, __fdiv_rn(spacing_h, 2.0f)), 0.5f), Hm1);
  const float nw = __fdiv_rn(__fmul_rn(spacing_w, static_cast<float>(crop_w - 1)), Wm1);
  const float nh = __fdiv_rn(__fmul_rn(spacing_h, static_cast<float>(crop_h - 1)), Hm1);
  const float by1 = ny0, bx1 = nx0, by2 = __fadd_rn(ny0, nh), bx2 = __fadd_rn(nx0, nw);
  // crop_and_resize
  const float height_scale = __fdiv_rn(__fmul_rn(__fsub_rn(by2, by1), Hm1), static_cast<float>(crop_h - 1));
  const float width_scale = __fdiv_rn(__fmul_rn(__fsub_rn(bx2, bx1), Wm1), static_cast<float>(crop_w - 1));
  const float in_y = __fadd_rn(__fmul_rn(by1, Hm1), __fmul_rn(static_cast<float>(iy), height_scale));
  const float in_x = __fadd_rn(__fmul_rn(bx1, Wm1), __fmul_rn(static_cast<float>(ix), width_scale));

  RoiSample sp;
  sp.img = b;
  sp.ok = (b >= 0) && (b < n_img) && (in_y >= 0.0f) && (in_y <= Hm1) && (in_x >= 0.0f) && (in_x <= Wm1);
  sp.top = static_cast<int>(floorf(in_y)); sp.bot = static_cast<int>(ceilf(in_y));
  sp.left = static_cast<int>(floorf(in_x)); sp.right = static_cast<int>(ceilf(in_x));
  sp.yl = in_y - static_cast<float>(sp.top);
  sp.xl = in_x - static_cast<float>(sp.left);
  return sp;
}

// ================================================================================================
// Dynamic Relation / Dynamic Walk: launch geometry and the affinity-conv phase shared by forward and backward
// ================================================================================================
constexpr int kDinThreads = 256;
constexpr int kDinWarps = kDinThreads / 32;
constexpr int kDinMaxN = 16;     // actors per frame (12 Volleyball, 13 Collective)
constexpr int kDinMaxK2 = 9;     // taps
constexpr int kDinMaxOutPerWarp = 4;  // ceil(27 / 8)


// One CTA = one (clip, frame).  Stages the kt input rows the frame's actors need (zero outside the clip / beyond
// the clip's real actors) in `xs` [kt][N][C] and computes the 3k² (or 2k²) conv outputs of every actor into
// `conv_s` [N][n_out]: p_conv rows (T-axis offsets, then N-axis offsets) followed by scale_conv rows
// (dynamic_infer_module.py:191-196).  warp w owns outputs o = w, w+8, ...; lanes sweep channels.
// Ends with __syncthreads(): xs and conv_s are ready for every thread.
__device__ __forceinline__ void din_stage_rows_and_conv(const float* __restrict__ xb, const float* __restrict__ w_tap,
                                                        const float* __restrict__ b_cat, float* xs, float* conv_s,
                                                        int t, int T, int N, int Nb, int C, int kt, int kn,
                                                        int ratio, int n_out, int dy0, int dx0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C4 = C >> 2;
  // ---- stage the kt rows
  for (int i = threadIdx.x; i < kt * N * C4; i += kDinThreads) {
    const int c4 = i % C4;
    const int rn = i / C4;
    const int n = rn % N;
    const int ky = rn / N;
    const int tt = t + dy0 + ky * ratio;
    float4 v = make_float4(0, 0, 0, 0);
    if (tt >= 0 && tt < T && n < Nb) v = __ldg(reinterpret_cast<const float4*>(xb + (static_cast<size_t>(tt) * N + n) * C) + c4);
    reinterpret_cast<float4*>(xs)[i] = v;
  }
  __syncthreads();

  // ---- phase 1: conv outputs.  warp w owns outputs o = w, w+8, ...; lanes sweep channels.
  {
    float acc[kDinMaxOutPerWarp][kDinMaxN];
#pragma unroll
    for (int a = 0; a < kDinMaxOutPerWarp; ++a)
#pragma unroll
      for (int n = 0; n < kDinMaxN; ++n) acc[a][n] = 0.0f;
    for (int ky = 0; ky < kt; ++ky) {
      for (int kx = 0; kx < kn; ++kx) {
        const int tap = ky * kn + kx;
        const int dx = dx0 + kx * ratio;
        const float* wt = w_tap + static_cast<size_t>(tap) * n_out * C;
        const float* xrow = xs + static_cast<size_t>(ky) * N * C;
        for (int c4 = lane; c4 < C4; c4 += 32) {
          float4 wv[kDinMaxOutPerWarp];
#pragma unroll
          for (int a = 0; a < kDinMaxOutPerWarp; ++a) {
            const int o = warp + a * kDinWarps;
            wv[a] = (o < n_out) ? __ldg(reinterpret_cast<const float4*>(wt + static_cast<size_t>(o) * C) + c4)
                                : make_float4(0, 0, 0, 0);
          }
#pragma unroll
          for (int n = 0; n < kDinMaxN; ++n) {
            const int nn = n + dx;
            if (n < Nb && nn >= 0 && nn < Nb) {
              const float4 xv = reinterpret_cast<const float4*>(xrow + static_cast<size_t>(nn) * C)[c4];
#pragma unroll
              for (int a = 0; a < kDinMaxOutPerWarp; ++a) {
                acc[a][n] = fmaf(wv[a].x, xv.x, acc[a][n]);
                acc[a][n] = fmaf(wv[a].y, xv.y, acc[a][n]);
                acc[a][n] = fmaf(wv[a].z, xv.z, acc[a][n]);
                acc[a][n] = fmaf(wv[a].w, xv.w, acc[a][n]);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int a = 0; a < kDinMaxOutPerWarp; ++a) {
      const int o = warp + a * kDinWarps;
#pragma unroll
      for (int n = 0; n < kDinMaxN; ++n) {
        const float v = warp_sum(acc[a][n]);
        if (lane == 0 && o < n_out && n < Nb) conv_s[n * n_out + o] = v + __ldg(b_cat + o);
      }
    }
  }
  __syncthreads();

}

}  // namespace din
