// head.cu — everything after the backbone: RoIAlign, embedding LayerNorm, the lite branch, the fused
// Dynamic Relation / Dynamic Walk step, fusion LayerNorm and the group read-out.  All fp32 except the
// RoIAlign input/output (fp16 feature map in, fp16 fc_emb operand out).
//
// Reference call sites:
//   RoIAlign                  infer_model.py:178-181 (external longcw/RoIAlign.pytorch crop_and_resize)
//   nl_emb_1 + ReLU           infer_model.py:185-186
//   point_conv/point_ln       infer_model.py:188-193
//   Dynamic_Person_Inference  infer_module/dynamic_infer_module.py:121-151, 184-282, 344-404
//   fusion + dpi_nl           infer_model.py:203-216 (volleyball), :1298-1301 (collective)
//   read-out                  infer_model.py:224-232 (volleyball), :1311-1313 (collective)
#include <cfloat>
#include <cstdio>
#include <cstdlib>

#include <cooperative_groups.h>

#include "din_common.cuh"
#include "din_head.cuh"

namespace {

using namespace din;

// ================================================================================================
// RoIAlign (crop_and_resize, one bilinear sample per bin, zero extrapolation)
// ================================================================================================
// One warp per (box, bin); lanes sweep the channel dimension 8 fp16 at a time (16-byte vectors), so
// every global access is a fully coalesced 512-byte warp transaction on the NHWC map.  The output row
// of box m is [bin][channel] — the K-order the packed fc_emb weight uses — so RoIAlign's result is
// directly the A operand of the embedding GEMM.
// OutT = __half: the operand of the tensor-core embedding GEMM.  OutT = float: the un-rounded crops for the fp32
// embedding path taken when there are fewer actor rows than half an MMA tile (engine.py: embed).
template <typename OutT>
__global__ void __launch_bounds__(256)
roi_align_kernel(const __half* __restrict__ fm, const float* __restrict__ boxes, const int* __restrict__ box_ind,
                 OutT* __restrict__ out, int n_img, int H, int W, int D, int fm_c_stride, int M, int crop_h,
                 int crop_w) {
  constexpr bool kF32 = sizeof(OutT) == 4;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int bins = crop_h * crop_w;
  if (warp_global >= M * bins) return;
  const int m = warp_global / bins;
  const int bin = warp_global - m * bins;
  const int iy = bin / crop_w;
  const int ix = bin - iy * crop_w;

  const RoiSample sp = roi_sample_point(boxes, box_ind, m, iy, ix, n_img, H, W, crop_h, crop_w);
  const int b = sp.img;

  OutT* op = out + (static_cast<size_t>(m) * bins + bin) * D;
  if (!sp.ok) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int c = lane * 8; c < D; c += 256) {
      *reinterpret_cast<uint4*>(op + c) = z;
      if (kF32) *reinterpret_cast<uint4*>(op + c + 4) = z;
    }
    return;
  }
  const int top = sp.top, bot = sp.bot, left = sp.left, right = sp.right;
  const float yl = sp.yl, xl = sp.xl;
  const __half* base = fm + static_cast<size_t>(b) * H * W * fm_c_stride;
  const __half* ptl = base + (static_cast<size_t>(top) * W + left) * fm_c_stride;
  const __half* ptr = base + (static_cast<size_t>(top) * W + right) * fm_c_stride;
  const __half* pbl = base + (static_cast<size_t>(bot) * W + left) * fm_c_stride;
  const __half* pbr = base + (static_cast<size_t>(bot) * W + right) * fm_c_stride;
  for (int c = lane * 8; c < D; c += 256) {
    const uint4 vtl = __ldg(reinterpret_cast<const uint4*>(ptl + c));
    const uint4 vtr = __ldg(reinterpret_cast<const uint4*>(ptr + c));
    const uint4 vbl = __ldg(reinterpret_cast<const uint4*>(pbl + c));
    const uint4 vbr = __ldg(reinterpret_cast<const uint4*>(pbr + c));
    const __half2* htl = reinterpret_cast<const __half2*>(&vtl);
    const __half2* htr = reinterpret_cast<const __half2*>(&vtr);
    const __half2* hbl = reinterpret_cast<const __half2*>(&vbl);
    const __half2* hbr = reinterpret_cast<const __half2*>(&vbr);
    float r[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 tl = __half22float2(htl[e]), tr = __half22float2(htr[e]);
      const float2 bl = __half22float2(hbl[e]), br = __half22float2(hbr[e]);
      const float t0 = tl.x + (tr.x - tl.x) * xl, t1 = tl.y + (tr.y - tl.y) * xl;
      const float b0 = bl.x + (br.x - bl.x) * xl, b1 = bl.y + (br.y - bl.y) * xl;
      r[2 * e] = t0 + (b0 - t0) * yl;
      r[2 * e + 1] = t1 + (b1 - t1) * yl;
    }
    if constexpr (kF32) {
      *reinterpret_cast<float4*>(op + c) = make_float4(r[0], r[1], r[2], r[3]);
      *reinterpret_cast<float4*>(op + c + 4) = make_float4(r[4], r[5], r[6], r[7]);
    } else {
      uint4 o;
      __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int e = 0; e < 4; ++e) ho[e] = __floats2half2_rn(r[2 * e], r[2 * e + 1]);
      *reinterpret_cast<uint4*>(op + c) = o;
    }
  }
}

// ================================================================================================
// LayerNorm over strided groups, with optional pre-add, ReLU and post-add
//   y = [relu]( LN(x (+ pre)) * gamma + beta ) (+ post)
// One CTA per group.  A group is `rows` x `cols` elements, element (r, c) at
//   base(g) + r*row_stride + c,   base(g) = (g / n_inner)*outer_stride + (g % n_inner)*inner_stride,
// gamma/beta indexed r*cols + c.  Covers nl_emb_1 (rows=1, one group per person), point_ln / dpi_nl
// / hier_LN (one group per clip, rows=1, cols=T*N*C) and Collective's LayerNorm([T, C]) over the
// [N, T, C] permutation of a [T, N, C] tensor (rows=T, row_stride=N*C, inner_stride=C).
// Two passes over L2-resident data (mean, then centred variance), biased variance as torch.
// ================================================================================================
constexpr int kLnThreads = 512;

__global__ void __launch_bounds__(kLnThreads)
group_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ pre, const float* __restrict__ post,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y,
                       int n_inner, long long outer_stride, long long inner_stride, int rows,
                       long long row_stride, int cols, float eps, int relu, const int* __restrict__ n_valid) {
  __shared__ float red[33];
  const int g = blockIdx.x;
  const int go = g / n_inner, gi = g - go * n_inner;
  if (n_valid != nullptr && gi >= __ldg(n_valid + go)) return;  // padded actor (Collective)
  const size_t base = static_cast<size_t>(go) * outer_stride + static_cast<size_t>(gi) * inner_stride;
  const int cols4 = cols >> 2;
  const int total4 = rows * cols4;
  float s = 0.0f;
  for (int i = threadIdx.x; i < total4; i += kLnThreads) {
    const int r = i / cols4, c4 = i - r * cols4;
    const size_t off = base + static_cast<size_t>(r) * row_stride + 4 * c4;
    float4 v = __ldg(reinterpret_cast<const float4*>(x + off));
    if (pre != nullptr) {
      const float4 p = __ldg(reinterpret_cast<const float4*>(pre + off));
      v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float n_el = static_cast<float>(rows) * static_cast<float>(cols);
  const float mean = block_sum<kLnThreads>(s, red) / n_el;
  float q = 0.0f;
  for (int i = threadIdx.x; i < total4; i += kLnThreads) {
    const int r = i / cols4, c4 = i - r * cols4;
    const size_t off = base + static_cast<size_t>(r) * row_stride + 4 * c4;
    float4 v = __ldg(reinterpret_cast<const float4*>(x + off));
    if (pre != nullptr) {
      const float4 p = __ldg(reinterpret_cast<const float4*>(pre + off));
      v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  const float var = block_sum<kLnThreads>(q, red) / n_el;
  const float rstd = rsqrtf(var + eps);
  for (int i = threadIdx.x; i < total4; i += kLnThreads) {
    const int r = i / cols4, c4 = i - r * cols4;
    const size_t off = base + static_cast<size_t>(r) * row_stride + 4 * c4;
    float4 v = __ldg(reinterpret_cast<const float4*>(x + off));
    if (pre != nullptr) {
      const float4 p = __ldg(reinterpret_cast<const float4*>(pre + off));
      v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    const size_t goff = static_cast<size_t>(r) * cols + 4 * c4;
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + goff));
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta + goff));
    float4 o;
    o.x = (v.x - mean) * rstd * ga.x + be.x;
    o.y = (v.y - mean) * rstd * ga.y + be.y;
    o.z = (v.z - mean) * rstd * ga.z + be.z;
    o.w = (v.w - mean) * rstd * ga.w + be.w;
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    if (post != nullptr) {
      const float4 p = __ldg(reinterpret_cast<const float4*>(post + off));
      o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
    }
    *reinterpret_cast<float4*>(y + off) = o;
  }
}

// ================================================================================================
// fp32 linear:  y[M,N] = x[M,K] · w[N,K]^T (+ bias) (ReLU)      (point_conv, hidden_weight)
// 64x64 tile, K step 16, 256 threads x (4x4) outputs, operands staged transposed in shared memory.
// ================================================================================================
constexpr int kLinBM = 64, kLinBN = 64, kLinBK = 16;

__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ y, int M, int N, int K, int relu, int accumulate) {
  __shared__ float xs[kLinBK][kLinBM + 4];
  __shared__ float ws[kLinBK][kLinBN + 4];
  const int m0 = blockIdx.y * kLinBM, n0 = blockIdx.x * kLinBN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  const int lr = threadIdx.x >> 2;        // 0..63: tile row loaded by this thread
  const int lk = (threadIdx.x & 3) * 4;   // 0,4,8,12
  for (int k0 = 0; k0 < K; k0 += kLinBK) {
    float4 xv = make_float4(0, 0, 0, 0), wv = make_float4(0, 0, 0, 0);
    if (m0 + lr < M) xv = __ldg(reinterpret_cast<const float4*>(x + static_cast<size_t>(m0 + lr) * K + k0 + lk));
    if (n0 + lr < N) wv = __ldg(reinterpret_cast<const float4*>(w + static_cast<size_t>(n0 + lr) * K + k0 + lk));
    __syncthreads();
    xs[lk + 0][lr] = xv.x; xs[lk + 1][lr] = xv.y; xs[lk + 2][lr] = xv.z; xs[lk + 3][lr] = xv.w;
    ws[lk + 0][lr] = wv.x; ws[lk + 1][lr] = wv.y; ws[lk + 2][lr] = wv.z; ws[lk + 3][lr] = wv.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kLinBK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&xs[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&ws[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.0f);
      if (relu) v = fmaxf(v, 0.0f);
      if (accumulate) v += y[static_cast<size_t>(m) * N + n];
      y[static_cast<size_t>(m) * N + n] = v;
    }
  }
}

// ================================================================================================
// Dynamic Relation + Dynamic Walk for one sampling ratio (dynamic_infer_ratio, :184-282), fused:
//   phase 1  offsets/relation logits = the two small convs over the T x N grid (p_conv, scale_conv):
//            one CTA per (clip, frame) keeps the kt input rows it needs in shared memory and streams
//            the tap-major weight [tap][3k²][C] once, reused across all N actors of the frame;
//   phase 2  per actor (one warp each): softmax over the k² taps with __shfl-free uniform math,
//            sample positions p = grid + tap + offset, detached floor, clamp-then-weight bilinear
//            blend of 4 corners gathered from the zero-padded clip, relation-weighted sum over taps;
//   y[b,t,n,:] (+)= coef * result      (coef = 1/len(ratios) or beta[r]; accumulate for ratios > first)
// Nothing but x is read from HBM and nothing but y is written: the 4 x [B,T,N,k²,C] gathers and the
// ~40 elementwise temporaries of the reference never exist.
// ================================================================================================
__global__ void __launch_bounds__(kDinThreads)
dynamic_infer_kernel(const float* __restrict__ x, const float* __restrict__ w_tap, const float* __restrict__ b_cat,
                     float* __restrict__ y, int T, int N, int C, int kt, int kn, int ratio, int scale_factor,
                     const float* __restrict__ coef_ptr, float coef_scalar, int accumulate,
                     const int* __restrict__ n_valid) {
  extern __shared__ float smem_f[];
  const int b = blockIdx.x / T;
  const int t = blockIdx.x - b * T;
  const int Nb = n_valid ? min(__ldg(n_valid + b), N) : N;  // real actors of this clip
  const int k2 = kt * kn;
  const int n_out = scale_factor ? 3 * k2 : 2 * k2;
  const int pt = (kt - 1) / 2 * ratio, pl = (kn - 1) / 2 * ratio;
  const int dy0 = -(((kt - 1) * ratio + 1) / 2);  // == python -(field-1)//2 for field = (k-1)*ratio+1
  const int dx0 = -(((kn - 1) * ratio + 1) / 2);
  float* xs = smem_f;                               // [kt][N][C] input rows t + dy (zero outside the clip)
  float* conv_s = smem_f + static_cast<size_t>(kt) * N * C;   // [N][n_out]
  const float* xb = x + static_cast<size_t>(b) * T * N * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C4 = C >> 2;

  din_stage_rows_and_conv(xb, w_tap, b_cat, xs, conv_s, t, T, N, Nb, C, kt, kn, ratio, n_out, dy0, dx0);

  // ---- phase 2: one warp per actor
  const float coef = coef_ptr ? __ldg(coef_ptr) : coef_scalar;
  const int Hp = T + 2 * pt, Wp = N + 2 * pl;   // padded extents; with n_valid the clip is [T, Nb]
  const int Wp_b = Nb + 2 * pl;
  for (int n = warp; n < Nb; n += kDinWarps) {
    const float* cs = conv_s + n * n_out;
    // relation weights: softmax over taps (uniform across the warp)
    float rel[kDinMaxK2];
    if (scale_factor) {
      float mx = -FLT_MAX;
      for (int k = 0; k < k2; ++k) mx = fmaxf(mx, cs[2 * k2 + k]);
      float den = 0.0f;
      for (int k = 0; k < k2; ++k) { rel[k] = expf(cs[2 * k2 + k] - mx); den += rel[k]; }
      for (int k = 0; k < k2; ++k) rel[k] = rel[k] / den;
    } else {
      for (int k = 0; k < k2; ++k) rel[k] = 1.0f / static_cast<float>(k2);
    }
    (void)Wp;
    float* yp = y + ((static_cast<size_t>(b) * T + t) * N + n) * C;
    for (int c4 = lane; c4 < C4; c4 += 32) {
      float4 o = make_float4(0, 0, 0, 0);
      for (int ky = 0; ky < kt; ++ky) {
        for (int kx = 0; kx < kn; ++kx) {
          const int k = ky * kn + kx;
          float py = static_cast<float>(pt + t + dy0 + ky * ratio) + cs[k];
          float px = static_cast<float>(pl + n + dx0 + kx * ratio) + cs[k2 + k];
          float ly = floorf(py), lx = floorf(px);
          float ry = ly + 1.0f, rx = lx + 1.0f;
          const float my = static_cast<float>(Hp - 1), mxx = static_cast<float>(Wp_b - 1);
          ly = fminf(fmaxf(ly, 0.0f), my); ry = fminf(fmaxf(ry, 0.0f), my); py = fminf(fmaxf(py, 0.0f), my);
          lx = fminf(fmaxf(lx, 0.0f), mxx); rx = fminf(fmaxf(rx, 0.0f), mxx); px = fminf(fmaxf(px, 0.0f), mxx);
          const float wly = 1.0f - fabsf(py - ly), wry = 1.0f - fabsf(py - ry);
          const float wlx = 1.0f - fabsf(px - lx), wrx = 1.0f - fabsf(px - rx);
          const int ily = static_cast<int>(ly) - pt, iry = static_cast<int>(ry) - pt;
          const int ilx = static_cast<int>(lx) - pl, irx = static_cast<int>(rx) - pl;
          auto fetch = [&](int ty, int tx) -> float4 {
            if (ty < 0 || ty >= T || tx < 0 || tx >= Nb) return make_float4(0, 0, 0, 0);
            return __ldg(reinterpret_cast<const float4*>(xb + (static_cast<size_t>(ty) * N + tx) * C) + c4);
          };
          const float4 lt = fetch(ily, ilx), rb = fetch(iry, irx), lb = fetch(iry, ilx), rt = fetch(ily, irx);
          const float wlt = wly * wlx, wrb = wry * wrx, wlb = wry * wlx, wrt = wly * wrx;
          // reference order: lt + rb + lb + rt, then * relation, summed over taps (:255-258, :278)
          const float fx = ((lt.x * wlt + rb.x * wrb) + lb.x * wlb) + rt.x * wrt;
          const float fy = ((lt.y * wlt + rb.y * wrb) + lb.y * wlb) + rt.y * wrt;
          const float fz = ((lt.z * wlt + rb.z * wrb) + lb.z * wlb) + rt.z * wrt;
          const float fw = ((lt.w * wlt + rb.w * wrb) + lb.w * wlb) + rt.w * wrt;
          o.x += fx * rel[k]; o.y += fy * rel[k]; o.z += fz * rel[k]; o.w += fw * rel[k];
        }
      }
      float4* dst = reinterpret_cast<float4*>(yp) + c4;
      if (accumulate) {
        const float4 prev = *dst;
        o.x = prev.x + coef * o.x; o.y = prev.y + coef * o.y; o.z = prev.z + coef * o.z; o.w = prev.w + coef * o.w;
      } else {
        o.x *= coef; o.y *= coef; o.z *= coef; o.w *= coef;
      }
      *dst = o;
    }
  }
}

// ================================================================================================
// The same step, parallel over channels: a thread-block CLUSTER per (clip, frame group).
//
// dynamic_infer_kernel above runs one CTA per (clip, frame): 80 CTAs for the bench's 8 clips, each streaming the whole
// tap-major weight (1 MB at C = 1024) from L2 through dependent loads -- ncu: 161 us at C = 1024 and 28 us at C = 128
// with 12 % of the warp slots active, for 0.5 GFLOP and 9 MB of compulsory traffic.  Here the channel dimension is
// split over the CTAs of a cluster (C/8 channels each at C = 1024, C/2 at C = 128) and the frames of a clip over
// `n_tg` clusters, so that the launch fills the chip:
//   load     each CTA stages the clip's whole [T][N][chunk] channel slice (zero beyond the clip's real actors) and its
//            [k2][n_out][chunk] slice of the weight in shared memory ONCE (coalesced 16-byte loads, all in flight);
//   phase 1  partial affinity-conv sums over the CTA's channels for its frames' actors: a warp owns 4 of the 3k2
//            outputs, lanes sweep channels, __shfl_xor reductions (as before);
//   exchange the partial sums are combined across the cluster through DISTRIBUTED SHARED MEMORY in rank order
//            (every CTA reads all ranks' partials: identical, deterministic totals everywhere), + bias;
//   phase 2  every CTA applies the now complete offsets / relation weights to ITS channels: softmax over the taps,
//            detached floor, clamp-then-weight corners, all gathers served from the staged slab in shared memory.
// Identical arithmetic per element except the association of the conv sums over channel chunks.
// ================================================================================================
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(kDinThreads)
dynamic_infer_cluster_kernel(const float* __restrict__ x, const float* __restrict__ w_tap,
                             const float* __restrict__ b_cat, float* __restrict__ y, int T, int N, int C, int kt, int kn,
                             int ratio, int scale_factor, const float* __restrict__ coef_ptr, float coef_scalar,
                             int accumulate, const int* __restrict__ n_valid, int n_tg, int chunk, int f_max) {
  extern __shared__ float smem_f[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());        // channel chunk
  const int n_chunks = static_cast<int>(cluster.num_blocks());
  const int b = blockIdx.y / n_tg, tg = blockIdx.y - b * n_tg;
  const int t0 = (tg * T) / n_tg, t1 = ((tg + 1) * T) / n_tg;     // frames of this cluster
  const int c0 = rank * chunk;
  const int Nb = n_valid ? min(__ldg(n_valid + b), N) : N;
  const int k2 = kt * kn;
  const int n_out = scale_factor ? 3 * k2 : 2 * k2;
  const int pt = (kt - 1) / 2 * ratio, pl = (kn - 1) / 2 * ratio;
  const int dy0 = -(((kt - 1) * ratio + 1) / 2);
  const int dx0 = -(((kn - 1) * ratio + 1) / 2);
  const int ch4 = chunk >> 2;
  float* slab = smem_f;                                            // [T][N][chunk]
  float* wts = slab + static_cast<size_t>(T) * N * chunk;          // [k2][n_out][chunk]
  float* part = wts + static_cast<size_t>(k2) * n_out * chunk;     // [f_max * N][n_out] this CTA's partial sums
  float* conv_s = part + static_cast<size_t>(f_max) * N * n_out;   // [f_max * N][n_out] totals (+ bias)
  const float* xb = x + static_cast<size_t>(b) * T * N * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- load: the clip's channel slice and this CTA's weight slice, as asynchronous 16-byte copies (cp.async): every
  //      copy of the CTA is in flight at once, no registers in between.  (The first version looped over __ldg + st.shared:
  //      15 + 30 dependent L2 round trips per thread, ~35 us of the kernel's 60 us per CTA.)
  for (int i = threadIdx.x; i < T * N * ch4; i += kDinThreads) {
    const int c4 = i % ch4, tn = i / ch4;
    const int n = tn % N;
    float4* dst = reinterpret_cast<float4*>(slab) + i;
    if (n < Nb) {
      const float4* src = reinterpret_cast<const float4*>(xb + static_cast<size_t>(tn) * C + c0) + c4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    } else {
      *dst = make_float4(0, 0, 0, 0);
    }
  }
  for (int i = threadIdx.x; i < k2 * n_out * ch4; i += kDinThreads) {
    const int c4 = i % ch4, row = i / ch4;
    const float4* src = reinterpret_cast<const float4*>(w_tap + static_cast<size_t>(row) * C + c0) + c4;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(reinterpret_cast<float4*>(wts) + i)), "l"(src)
                 : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- phase 1: partial conv sums over this CTA's channels, frame by frame
  for (int t = t0; t < t1; ++t) {
    float acc[kDinMaxOutPerWarp][kDinMaxN];
#pragma unroll
    for (int a = 0; a < kDinMaxOutPerWarp; ++a)
#pragma unroll
      for (int n = 0; n < kDinMaxN; ++n) acc[a][n] = 0.0f;
    for (int ky = 0; ky < kt; ++ky) {
      const int tt = t + dy0 + ky * ratio;
      if (tt < 0 || tt >= T) continue;                             // zero rows outside the clip
      const float* xrow = slab + static_cast<size_t>(tt) * N * chunk;
      for (int kx = 0; kx < kn; ++kx) {
        const int dx = dx0 + kx * ratio;
        const float* wt = wts + static_cast<size_t>(ky * kn + kx) * n_out * chunk;
        for (int c4 = lane; c4 < ch4; c4 += 32) {
          float4 wv[kDinMaxOutPerWarp];
#pragma unroll
          for (int a = 0; a < kDinMaxOutPerWarp; ++a) {
            const int o = warp + a * kDinWarps;
            wv[a] = (o < n_out) ? reinterpret_cast<const float4*>(wt + static_cast<size_t>(o) * chunk)[c4]
                                : make_float4(0, 0, 0, 0);
          }
#pragma unroll
          for (int n = 0; n < kDinMaxN; ++n) {
            const int nn = n + dx;
            if (n < Nb && nn >= 0 && nn < Nb) {
              const float4 xv = reinterpret_cast<const float4*>(xrow + static_cast<size_t>(nn) * chunk)[c4];
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
    float* pf = part + static_cast<size_t>(t - t0) * N * n_out;
#pragma unroll
    for (int a = 0; a < kDinMaxOutPerWarp; ++a) {
      const int o = warp + a * kDinWarps;
#pragma unroll
      for (int n = 0; n < kDinMaxN; ++n) {
        const float v = warp_sum(acc[a][n]);
        if (lane == 0 && o < n_out && n < Nb) pf[n * n_out + o] = v;
      }
    }
  }

  // ---- exchange: totals = sum over the cluster's ranks (in rank order) + bias
  cluster.sync();
  {
    const int total = (t1 - t0) * N * n_out;
    for (int i = threadIdx.x; i < total; i += kDinThreads) {
      const int n = (i / n_out) % N;
      if (n >= Nb) continue;
      float s = 0.0f;
      for (int r = 0; r < n_chunks; ++r) s += cluster.map_shared_rank(part, r)[i];
      conv_s[i] = s + __ldg(b_cat + (i % n_out));
    }
  }
  cluster.sync();                                                  // nobody overwrites / leaves while peers still read

  // ---- phase 2: one warp per (frame, actor), this CTA's channels
  const float coef = coef_ptr ? __ldg(coef_ptr) : coef_scalar;
  const int Hp = T + 2 * pt;
  const int Wp_b = Nb + 2 * pl;
  const int nodes = (t1 - t0) * Nb;
  for (int node = warp; node < nodes; node += kDinWarps) {
    const int tf = node / Nb, n = node - tf * Nb;
    const int t = t0 + tf;
    const float* cs = conv_s + (static_cast<size_t>(tf) * N + n) * n_out;
    float rel[kDinMaxK2];
    if (scale_factor) {
      float mx = -FLT_MAX;
      for (int k = 0; k < k2; ++k) mx = fmaxf(mx, cs[2 * k2 + k]);
      float den = 0.0f;
      for (int k = 0; k < k2; ++k) { rel[k] = expf(cs[2 * k2 + k] - mx); den += rel[k]; }
      for (int k = 0; k < k2; ++k) rel[k] = rel[k] / den;
    } else {
      for (int k = 0; k < k2; ++k) rel[k] = 1.0f / static_cast<float>(k2);
    }
    float* yp = y + ((static_cast<size_t>(b) * T + t) * N + n) * C + c0;
    for (int c4 = lane; c4 < ch4; c4 += 32) {
      float4 o = make_float4(0, 0, 0, 0);
      for (int ky = 0; ky < kt; ++ky) {
        for (int kx = 0; kx < kn; ++kx) {
          const int k = ky * kn + kx;
          float py = static_cast<float>(pt + t + dy0 + ky * ratio) + cs[k];
          float px = static_cast<float>(pl + n + dx0 + kx * ratio) + cs[k2 + k];
          float ly = floorf(py), lx = floorf(px);
          float ry = ly + 1.0f, rx = lx + 1.0f;
          const float my = static_cast<float>(Hp - 1), mxx = static_cast<float>(Wp_b - 1);
          ly = fminf(fmaxf(ly, 0.0f), my); ry = fminf(fmaxf(ry, 0.0f), my); py = fminf(fmaxf(py, 0.0f), my);
          lx = fminf(fmaxf(lx, 0.0f), mxx); rx = fminf(fmaxf(rx, 0.0f), mxx); px = fminf(fmaxf(px, 0.0f), mxx);
          const float wly = 1.0f - fabsf(py - ly), wry = 1.0f - fabsf(py - ry);
          const float wlx = 1.0f - fabsf(px - lx), wrx = 1.0f - fabsf(px - rx);
          const int ily = static_cast<int>(ly) - pt, iry = static_cast<int>(ry) - pt;
          const int ilx = static_cast<int>(lx) - pl, irx = static_cast<int>(rx) - pl;
          auto fetch = [&](int ty, int tx) -> float4 {
            if (ty < 0 || ty >= T || tx < 0 || tx >= Nb) return make_float4(0, 0, 0, 0);
            return reinterpret_cast<const float4*>(slab + (static_cast<size_t>(ty) * N + tx) * chunk)[c4];
          };
          const float4 lt = fetch(ily, ilx), rb = fetch(iry, irx), lb = fetch(iry, ilx), rt = fetch(ily, irx);
          const float wlt = wly * wlx, wrb = wry * wrx, wlb = wry * wlx, wrt = wly * wrx;
          // reference order: lt + rb + lb + rt, then * relation, summed over taps (:255-258, :278)
          const float fx = ((lt.x * wlt + rb.x * wrb) + lb.x * wlb) + rt.x * wrt;
          const float fy = ((lt.y * wlt + rb.y * wrb) + lb.y * wlb) + rt.y * wrt;
          const float fz = ((lt.z * wlt + rb.z * wrb) + lb.z * wlb) + rt.z * wrt;
          const float fw = ((lt.w * wlt + rb.w * wrb) + lb.w * wlb) + rt.w * wrt;
          o.x += fx * rel[k]; o.y += fy * rel[k]; o.z += fz * rel[k]; o.w += fw * rel[k];
        }
      }
      float4* dst = reinterpret_cast<float4*>(yp) + c4;
      if (accumulate) {
        const float4 prev = *dst;
        o.x = prev.x + coef * o.x; o.y = prev.y + coef * o.y; o.z = prev.z + coef * o.z; o.w = prev.w + coef * o.w;
      } else {
        o.x *= coef; o.y *= coef; o.z *= coef; o.w *= coef;
      }
      *dst = o;
    }
  }
}

// ================================================================================================
// Context attention of Dynamic_TCE_volleyball (TCE_STBiP_module.py:252-287, one layer of 4 heads): every actor attends
// over the frame's feature map,
//     a[n][px] = <q[n], x[px]>,  A = softmax_px(a),  ctx[n] = sum_px A[n][px] * x[px],   x[px] = img[px] + posbias[px]
// with img = downsample2(feature map) (the four heads' 1x1 convolutions run as ONE tcgen05 GEMM, fp32 out) and posbias =
// downsample2(position embedding) + bias precomputed per plan (the embedding is a constant of the map size).
// One CTA per (frame, head): q and the attention row of every actor live in shared memory, the map streams through a
// 32-pixel tile twice (scores, then the weighted sum); nothing but ctx is written.
//   q, ctx: [H][M][128] (head-major, M = frames * actors);  img: [F][P][H*128] fp32;  posbias: [P][H*128] fp32.
// ================================================================================================
constexpr int kCtxDim = 128;
constexpr int kCtxTile = 32;
constexpr int kCtxThreads = 256;

__global__ void __launch_bounds__(kCtxThreads)
context_attention_kernel(const float* __restrict__ q, const float* __restrict__ img, const float* __restrict__ posbias,
                         float* __restrict__ ctx, int F, int N, int P, int H) {
  extern __shared__ float smem_f[];
  const int f = blockIdx.x / H, h = blockIdx.x - f * H;
  const int M = F * N;
  float* qs = smem_f;                                   // [N][128]
  float* tile = qs + N * kCtxDim;                       // [32][129]
  float* att = tile + kCtxTile * (kCtxDim + 1);         // [N][P]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HC = H * kCtxDim;
  for (int i = tid; i < N * kCtxDim; i += kCtxThreads)
    qs[i] = __ldg(q + (static_cast<size_t>(h) * M + static_cast<size_t>(f) * N) * kCtxDim + i);
  const float* ib = img + static_cast<size_t>(f) * P * HC + h * kCtxDim;
  const float* pb = posbias + h * kCtxDim;

  auto load_tile = [&](int p0) {
    for (int i = tid; i < kCtxTile * (kCtxDim / 4); i += kCtxThreads) {
      const int px = i / (kCtxDim / 4), c4 = i - px * (kCtxDim / 4);
      float4 v = make_float4(0, 0, 0, 0);
      if (p0 + px < P) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(ib + static_cast<size_t>(p0 + px) * HC) + c4);
        const float4 b = __ldg(reinterpret_cast<const float4*>(pb + static_cast<size_t>(p0 + px) * HC) + c4);
        v = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
      }
      float* t = tile + px * (kCtxDim + 1) + 4 * c4;
      t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
    }
  };

  // ---- pass 1: scores
  for (int p0 = 0; p0 < P; p0 += kCtxTile) {
    __syncthreads();
    load_tile(p0);
    __syncthreads();
    for (int i = tid; i < N * kCtxTile; i += kCtxThreads) {
      const int n = i / kCtxTile, px = i - n * kCtxTile;
      if (p0 + px < P) {
        const float* qq = qs + n * kCtxDim;
        const float* xx = tile + px * (kCtxDim + 1);
        float acc = 0.0f;
#pragma unroll 8
        for (int c = 0; c < kCtxDim; ++c) acc = fmaf(qq[c], xx[c], acc);
        att[n * P + p0 + px] = acc;
      }
    }
  }
  __syncthreads();
  // ---- softmax over the map, one warp per actor
  for (int n = warp; n < N; n += kCtxThreads / 32) {
    float* a = att + n * P;
    float mx = -FLT_MAX;
    for (int i = lane; i < P; i += 32) mx = fmaxf(mx, a[i]);
    mx = warp_max(mx);
    float den = 0.0f;
    for (int i = lane; i < P; i += 32) { const float e = expf(a[i] - mx); a[i] = e; den += e; }
    den = warp_sum(den);
    const float inv = 1.0f / den;
    for (int i = lane; i < P; i += 32) a[i] *= inv;
  }
  // ---- pass 2: weighted sum; thread = (channel, actor parity)
  const int c = tid & (kCtxDim - 1), n0 = tid >> 7;          // 256 threads: n0 in {0, 1}
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
  for (int p0 = 0; p0 < P; p0 += kCtxTile) {
    __syncthreads();
    load_tile(p0);
    __syncthreads();
    const int lim = min(kCtxTile, P - p0);
    for (int px = 0; px < lim; ++px) {
      const float xv = tile[px * (kCtxDim + 1) + c];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int n = n0 + 2 * k;
        if (n < N) acc[k] = fmaf(att[n * P + p0 + px], xv, acc[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int n = n0 + 2 * k;
    if (n < N) ctx[(static_cast<size_t>(h) * M + static_cast<size_t>(f) * N + n) * kCtxDim + c] = acc[k];
  }
}

// Backward of context_attention_kernel (autograd through torch.matmul / F.softmax / torch.matmul of
// TCE_STBiP_module.py:274-280 when Dynamic_TCE_volleyball trains).  With x[px] = img[px] + posbias[px]:
//     dA[n][px] = <dctx[n], x[px]>,   ds = A * (dA - sum_px A * dA),
//     dq[n] = sum_px ds[n][px] * x[px],   dimg[px] = sum_n (ds[n][px] * q[n] + A[n][px] * dctx[n])
// (x is both key and value, so its gradient collects both terms; posbias is a constant of the plan.)
// One CTA per (frame, head), as the forward: scores and softmax recomputed with the forward's arithmetic, A and dA / ds for
// all actors in shared memory, the map streamed three times through the 32-pixel tile; every output element is written by
// exactly one thread (no atomics: deterministic).
__global__ void __launch_bounds__(kCtxThreads)
context_attention_bwd_kernel(const float* __restrict__ q, const float* __restrict__ img, const float* __restrict__ posbias,
                             const float* __restrict__ dctx, const float* __restrict__ dq_add, float* __restrict__ dq,
                             float* __restrict__ dimg, int F, int N, int P, int H) {
  extern __shared__ float smem_f[];
  const int f = blockIdx.x / H, h = blockIdx.x - f * H;
  const int M = F * N;
  float* qs = smem_f;                                   // [N][128]
  float* gs = qs + N * kCtxDim;                         // [N][128]  dctx
  float* tile = gs + N * kCtxDim;                       // [32][129]
  float* att = tile + kCtxTile * (kCtxDim + 1);         // [N][P]  A
  float* dsm = att + N * P;                             // [N][P]  dA, then ds
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HC = H * kCtxDim;
  const size_t row0 = (static_cast<size_t>(h) * M + static_cast<size_t>(f) * N) * kCtxDim;
  for (int i = tid; i < N * kCtxDim; i += kCtxThreads) {
    qs[i] = __ldg(q + row0 + i);
    gs[i] = __ldg(dctx + row0 + i);
  }
  const float* ib = img + static_cast<size_t>(f) * P * HC + h * kCtxDim;
  const float* pb = posbias + h * kCtxDim;
  float* ob = dimg + static_cast<size_t>(f) * P * HC + h * kCtxDim;

  auto load_tile = [&](int p0) {
    for (int i = tid; i < kCtxTile * (kCtxDim / 4); i += kCtxThreads) {
      const int px = i / (kCtxDim / 4), c4 = i - px * (kCtxDim / 4);
      float4 v = make_float4(0, 0, 0, 0);
      if (p0 + px < P) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(ib + static_cast<size_t>(p0 + px) * HC) + c4);
        const float4 b = __ldg(reinterpret_cast<const float4*>(pb + static_cast<size_t>(p0 + px) * HC) + c4);
        v = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
      }
      float* t = tile + px * (kCtxDim + 1) + 4 * c4;
      t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
    }
  };

  // ---- pass 1: scores and dA (same dot products, against q and dctx)
  for (int p0 = 0; p0 < P; p0 += kCtxTile) {
    __syncthreads();
    load_tile(p0);
    __syncthreads();
    for (int i = tid; i < N * kCtxTile; i += kCtxThreads) {
      const int n = i / kCtxTile, px = i - n * kCtxTile;
      if (p0 + px < P) {
        const float* qq = qs + n * kCtxDim;
        const float* gg = gs + n * kCtxDim;
        const float* xx = tile + px * (kCtxDim + 1);
        float acc = 0.0f, accg = 0.0f;
#pragma unroll 8
        for (int c = 0; c < kCtxDim; ++c) {
          acc = fmaf(qq[c], xx[c], acc);
          accg = fmaf(gg[c], xx[c], accg);
        }
        att[n * P + p0 + px] = acc;
        dsm[n * P + p0 + px] = accg;
      }
    }
  }
  __syncthreads();
  // ---- softmax (the forward's arithmetic), then ds = A * (dA - <A, dA>), one warp per actor
  for (int n = warp; n < N; n += kCtxThreads / 32) {
    float* a = att + n * P;
    float* d = dsm + n * P;
    float mx = -FLT_MAX;
    for (int i = lane; i < P; i += 32) mx = fmaxf(mx, a[i]);
    mx = warp_max(mx);
    float den = 0.0f;
    for (int i = lane; i < P; i += 32) { const float e = expf(a[i] - mx); a[i] = e; den += e; }
    den = warp_sum(den);
    const float inv = 1.0f / den;
    float dot = 0.0f;
    for (int i = lane; i < P; i += 32) { a[i] *= inv; dot = fmaf(a[i], d[i], dot); }
    dot = warp_sum(dot);
    for (int i = lane; i < P; i += 32) d[i] = a[i] * (d[i] - dot);
  }
  // ---- pass 2: dq (thread = (channel, actor parity), as the forward's weighted sum) and dimg (thread = (pixel, channel))
  const int c = tid & (kCtxDim - 1), n0 = tid >> 7;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
  for (int p0 = 0; p0 < P; p0 += kCtxTile) {
    __syncthreads();
    load_tile(p0);
    __syncthreads();
    const int lim = min(kCtxTile, P - p0);
    for (int px = 0; px < lim; ++px) {
      const float xv = tile[px * (kCtxDim + 1) + c];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int n = n0 + 2 * k;
        if (n < N) acc[k] = fmaf(dsm[n * P + p0 + px], xv, acc[k]);
      }
    }
    for (int i = tid; i < lim * kCtxDim; i += kCtxThreads) {
      const int px = i >> 7, cc = i & (kCtxDim - 1);
      float o = 0.0f;
      for (int n = 0; n < N; ++n)
        o = fmaf(dsm[n * P + p0 + px], qs[n * kCtxDim + cc], fmaf(att[n * P + p0 + px], gs[n * kCtxDim + cc], o));
      ob[static_cast<size_t>(p0 + px) * HC + cc] = o;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int n = n0 + 2 * k;
    if (n < N) {                                          // dq_add: the gradient q receives through the residual of layernorm1
      const size_t o = row0 + static_cast<size_t>(n) * kCtxDim + c;
      dq[o] = acc[k] + (dq_add != nullptr ? __ldg(dq_add + o) : 0.0f);
    }
  }
}

// ================================================================================================
// read-out: max over actors -> fc_activities -> mean over frames     (one CTA per clip)
// ================================================================================================
constexpr int kRoThreads = 256;

__global__ void __launch_bounds__(kRoThreads)
readout_kernel(const float* __restrict__ s, const float* __restrict__ w, const float* __restrict__ bias,
               float* __restrict__ logits, int T, int N, int C, int A, const int* __restrict__ n_valid) {
  extern __shared__ float pooled[];  // [C]
  __shared__ float score[64];        // A <= 64 running sum over frames
  const int b = blockIdx.x;
  const int Nb = n_valid ? min(__ldg(n_valid + b), N) : N;
  if (Nb < 1) {   // the reference's torch.max over an empty actor axis raises (infer_model.py:1311); -FLT_MAX logits would not
    if (threadIdx.x == 0) printf("din_readout_f32: clip %d has no valid actor\n", b);
    __trap();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 64) score[threadIdx.x] = 0.0f;
  for (int t = 0; t < T; ++t) {
    __syncthreads();
    const float* st = s + (static_cast<size_t>(b) * T + t) * N * C;
    for (int c = threadIdx.x; c < C; c += kRoThreads) {
      float m = -FLT_MAX;
      for (int n = 0; n < Nb; ++n) m = fmaxf(m, __ldg(st + static_cast<size_t>(n) * C + c));
      pooled[c] = m;
    }
    __syncthreads();
    for (int a = warp; a < A; a += kRoThreads / 32) {
      float d = 0.0f;
      for (int c = lane; c < C; c += 32) d = fmaf(pooled[c], __ldg(w + static_cast<size_t>(a) * C + c), d);
      d = warp_sum(d);
      if (lane == 0) score[a] += d + __ldg(bias + a);
    }
  }
  __syncthreads();
  if (threadIdx.x < A) logits[static_cast<size_t>(b) * A + threadIdx.x] = score[threadIdx.x] / static_cast<float>(T);
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
static int roi_align_launch(const char* who, const void* fm, const float* boxes, const int32_t* box_ind, void* out,
                            bool out_f32, int n_img, int h, int w, int d, int fm_c_stride, int m, int crop_h,
                            int crop_w, void* stream) {
  DIN_CHECK_ARG(fm && boxes && box_ind && out, "%s: null pointer", who);
  DIN_CHECK_ARG(n_img > 0 && h > 1 && w > 1 && m > 0, "%s: bad extent n=%d h=%d w=%d m=%d", who, n_img, h, w, m);
  DIN_CHECK_ARG(d > 0 && d % 8 == 0 && fm_c_stride >= d && fm_c_stride % 8 == 0,
                "%s: d=%d / fm_c_stride=%d must be multiples of 8", who, d, fm_c_stride);
  DIN_CHECK_ARG(crop_h > 1 && crop_w > 1, "%s: crop must be > 1", who);
  DIN_CHECK_ARG((reinterpret_cast<uintptr_t>(fm) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                "%s: pointers must be 16-byte aligned", who);
  const long long warps = static_cast<long long>(m) * crop_h * crop_w;
  const int grid = static_cast<int>((warps + 7) / 8);
  if (out_f32)
    roi_align_kernel<float><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __half*>(fm), boxes, box_ind, static_cast<float*>(out), n_img, h, w, d, fm_c_stride, m,
        crop_h, crop_w);
  else
    roi_align_kernel<__half><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __half*>(fm), boxes, box_ind, static_cast<__half*>(out), n_img, h, w, d, fm_c_stride, m,
        crop_h, crop_w);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_roi_align_nhwc_f16(const void* fm, const float* boxes, const int32_t* box_ind, void* out,
                                      int n_img, int h, int w, int d, int fm_c_stride, int m, int crop_h,
                                      int crop_w, void* stream) {
  return roi_align_launch("din_roi_align_nhwc_f16", fm, boxes, box_ind, out, false, n_img, h, w, d, fm_c_stride, m,
                          crop_h, crop_w, stream);
}

extern "C" int din_roi_align_nhwc_f16_f32out(const void* fm, const float* boxes, const int32_t* box_ind, float* out,
                                             int n_img, int h, int w, int d, int fm_c_stride, int m, int crop_h,
                                             int crop_w, void* stream) {
  return roi_align_launch("din_roi_align_nhwc_f16_f32out", fm, boxes, box_ind, out, true, n_img, h, w, d, fm_c_stride,
                          m, crop_h, crop_w, stream);
}

extern "C" int din_group_layernorm_f32(const float* x, const float* pre, const float* post, const float* gamma,
                                       const float* beta, float* y, int n_outer, int n_inner,
                                       long long outer_stride, long long inner_stride, int rows,
                                       long long row_stride, int cols, float eps, int relu,
                                       const int32_t* n_valid, void* stream) {
  DIN_CHECK_ARG(x && gamma && beta && y, "din_group_layernorm_f32: null pointer");
  DIN_CHECK_ARG(n_outer > 0 && n_inner > 0 && rows > 0 && cols > 0 && cols % 4 == 0,
                "din_group_layernorm_f32: bad shape outer=%d inner=%d rows=%d cols=%d (cols %% 4 == 0)", n_outer,
                n_inner, rows, cols);
  DIN_CHECK_ARG(outer_stride % 4 == 0 && inner_stride % 4 == 0 && row_stride % 4 == 0,
                "din_group_layernorm_f32: strides must be multiples of 4 elements");
  DIN_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                  reinterpret_cast<uintptr_t>(pre) | reinterpret_cast<uintptr_t>(post) |
                  reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
                "din_group_layernorm_f32: pointers must be 16-byte aligned");
  group_layernorm_kernel<<<n_outer * n_inner, kLnThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x, pre, post, gamma, beta, y, n_inner, outer_stride, inner_stride, rows, row_stride, cols, eps, relu, n_valid);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_linear_f32(const float* x, const float* w, const float* bias, float* y, int m, int n, int k,
                              int relu, int accumulate, void* stream) {
  DIN_CHECK_ARG(x && w && y, "din_linear_f32: null pointer");
  DIN_CHECK_ARG(m > 0 && n > 0 && k > 0 && k % 16 == 0, "din_linear_f32: bad shape m=%d n=%d k=%d (k %% 16 == 0)", m,
                n, k);
  DIN_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w)) & 15) == 0,
                "din_linear_f32: x and w must be 16-byte aligned");
  dim3 grid((n + kLinBN - 1) / kLinBN, (m + kLinBM - 1) / kLinBM);
  linear_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, w, bias, y, m, n, k, relu, accumulate);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_dynamic_infer_f32(const float* x, const float* w_tap, const float* b_cat, float* y, int b, int t,
                                     int n, int c, int kt, int kn, int ratio, int scale_factor,
                                     const float* coef_ptr, float coef_scalar, int accumulate,
                                     const int32_t* n_valid, void* stream) {
  DIN_CHECK_ARG(x && w_tap && b_cat && y, "din_dynamic_infer_f32: null pointer");
  DIN_CHECK_ARG(b > 0 && t > 0 && n > 0 && n <= kDinMaxN, "din_dynamic_infer_f32: bad extent b=%d t=%d n=%d (n <= %d)",
                b, t, n, kDinMaxN);
  DIN_CHECK_ARG(c > 0 && c % 4 == 0, "din_dynamic_infer_f32: c=%d must be a multiple of 4", c);
  DIN_CHECK_ARG(kt >= 1 && kn >= 1 && kt * kn <= kDinMaxK2 && (kt & 1) && (kn & 1),
                "din_dynamic_infer_f32: kernel %dx%d unsupported (odd, <= %d taps)", kt, kn, kDinMaxK2);
  DIN_CHECK_ARG(ratio >= 1, "din_dynamic_infer_f32: ratio=%d", ratio);
  const int n_out = (scale_factor ? 3 : 2) * kt * kn;
  DIN_CHECK_ARG(n_out <= kDinMaxOutPerWarp * kDinWarps, "din_dynamic_infer_f32: too many conv outputs");
  DIN_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                  reinterpret_cast<uintptr_t>(w_tap)) & 15) == 0,
                "din_dynamic_infer_f32: pointers must be 16-byte aligned");
  // ---- cluster formulation (channels over the CTAs of a cluster, frames over clusters): whenever the channel count
  //      splits into 2 / 4 / 8 chunks of a multiple of 32 channels and the staged slices fit shared memory
  {
    static const bool use_cluster = [] { const char* e = std::getenv("DIN_DI_CLUSTER"); return !(e && e[0] == '0'); }();
    // 128-channel chunks: one float4 per lane per row in both phases (a 64-channel chunk leaves half of every warp idle:
    // C = 128 split in two measured slower than the one-CTA-per-frame kernel); C = 128 runs as clusters of one CTA
    int n_chunks = 0;
    if (c % 128 == 0 && (c / 128 == 1 || c / 128 == 2 || c / 128 == 4 || c / 128 == 8)) n_chunks = c / 128;
    if (use_cluster && n_chunks > 0) {
      const int chunk = c / n_chunks;
      const int sms = din_num_sms();
      int n_tg = (2 * (sms > 0 ? sms : 148) + b * n_chunks - 1) / (b * n_chunks);     // ~2 CTAs per SM over the launch
      if (n_tg > t) n_tg = t;
      if (n_tg < 1) n_tg = 1;
      auto smem_for = [&](int tg) {
        const int f_max = (t + tg - 1) / tg;
        return (static_cast<size_t>(t) * n * chunk + static_cast<size_t>(kt) * kn * n_out * chunk +
                2 * static_cast<size_t>(f_max) * n * n_out) * sizeof(float);
      };
      while (n_tg < t && smem_for(n_tg) > 200 * 1024) ++n_tg;
      const size_t smem_c = smem_for(n_tg);
      if (smem_c <= 200 * 1024) {
        const int f_max = (t + n_tg - 1) / n_tg;
        DIN_OPT_IN_SMEM(dynamic_infer_cluster_kernel, smem_c);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(n_chunks, b * n_tg, 1);
        cfg.blockDim = dim3(kDinThreads, 1, 1);
        cfg.dynamicSmemBytes = smem_c;
        cfg.stream = static_cast<cudaStream_t>(stream);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = n_chunks; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        DIN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, dynamic_infer_cluster_kernel, x, w_tap, b_cat, y, t, n, c, kt, kn, ratio,
                                          scale_factor, coef_ptr, coef_scalar, accumulate,
                                          static_cast<const int*>(n_valid), n_tg, chunk, f_max));
        return DIN_OK;
      }
    }
  }
  const size_t smem = (static_cast<size_t>(kt) * n * c + static_cast<size_t>(n) * n_out) * sizeof(float);
  DIN_CHECK_ARG(smem <= 226 * 1024, "din_dynamic_infer_f32: kt*n*c too large for shared memory (%zu bytes)", smem);
  DIN_OPT_IN_SMEM(dynamic_infer_kernel, smem);
  dynamic_infer_kernel<<<b * t, kDinThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      x, w_tap, b_cat, y, t, n, c, kt, kn, ratio, scale_factor, coef_ptr, coef_scalar, accumulate, n_valid);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_readout_f32(const float* s, const float* w, const float* bias, float* logits, int b, int t,
                               int n, int c, int a, const int32_t* n_valid, void* stream) {
  DIN_CHECK_ARG(s && w && bias && logits, "din_readout_f32: null pointer");
  DIN_CHECK_ARG(b > 0 && t > 0 && n > 0 && c > 0 && a > 0 && a <= 64,
                "din_readout_f32: bad shape b=%d t=%d n=%d c=%d a=%d (a <= 64)", b, t, n, c, a);
  const size_t smem = static_cast<size_t>(c) * sizeof(float);
  DIN_CHECK_ARG(smem <= 48 * 1024, "din_readout_f32: c=%d too large", c);
  readout_kernel<<<b, kRoThreads, smem, static_cast<cudaStream_t>(stream)>>>(s, w, bias, logits, t, n, c, a, n_valid);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_context_attention_bwd_f32(const float* q, const float* img, const float* posbias, const float* dctx,
                                             const float* dq_add, float* dq, float* dimg, int frames, int n, int pixels,
                                             int heads, void* stream) {
  const char* who = "din_context_attention_bwd_f32";
  DIN_CHECK_ARG(q && img && posbias && dctx && dq && dimg, "%s: null pointer", who);
  DIN_CHECK_ARG(frames > 0 && n > 0 && n <= 16 && pixels > 0 && heads > 0,
                "%s: bad shape frames=%d n=%d (<= 16) pixels=%d heads=%d", who, frames, n, pixels, heads);
  DIN_CHECK_ARG(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(posbias) |
                  reinterpret_cast<uintptr_t>(dctx) | reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dimg)) & 15) == 0,
                "%s: pointers must be 16-byte aligned", who);
  const size_t smem = (2 * static_cast<size_t>(n) * kCtxDim + kCtxTile * (kCtxDim + 1) + 2 * static_cast<size_t>(n) * pixels) *
                      sizeof(float);
  DIN_CHECK_ARG(smem <= 220 * 1024, "%s: n * pixels = %d x %d does not fit shared memory", who, n, pixels);
  DIN_OPT_IN_SMEM(context_attention_bwd_kernel, smem);
  context_attention_bwd_kernel<<<frames * heads, kCtxThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      q, img, posbias, dctx, dq_add, dq, dimg, frames, n, pixels, heads);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_context_attention_f32(const float* q, const float* img, const float* posbias, float* ctx, int frames,
                                         int n, int pixels, int heads, void* stream) {
  DIN_CHECK_ARG(q && img && posbias && ctx, "din_context_attention_f32: null pointer");
  DIN_CHECK_ARG(frames > 0 && n > 0 && n <= 16 && pixels > 0 && heads > 0,
                "din_context_attention_f32: bad shape frames=%d n=%d (<= 16) pixels=%d heads=%d", frames, n, pixels, heads);
  DIN_CHECK_ARG(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(posbias) |
                  reinterpret_cast<uintptr_t>(ctx)) & 15) == 0,
                "din_context_attention_f32: pointers must be 16-byte aligned");
  const size_t smem = (static_cast<size_t>(n) * kCtxDim + kCtxTile * (kCtxDim + 1) + static_cast<size_t>(n) * pixels) *
                      sizeof(float);
  DIN_CHECK_ARG(smem <= 200 * 1024, "din_context_attention_f32: n * pixels = %d x %d does not fit shared memory", n, pixels);
  DIN_OPT_IN_SMEM(context_attention_kernel, smem);
  context_attention_kernel<<<frames * heads, kCtxThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      q, img, posbias, ctx, frames, n, pixels, heads);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}
