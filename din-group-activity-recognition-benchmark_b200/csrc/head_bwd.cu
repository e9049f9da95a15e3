// head_bwd.cu — backward of the person-level head (first slice: everything after the
// feature map, i.e. the stage-2 training step with a frozen backbone, config.py:39 `train_backbone = False`).
//
// What autograd does in the reference for these ops (train_net_dynamic.py:220-224 `total_loss.backward()`):
//   read-out                  infer_model.py:224-232     max over actors -> Linear -> mean over frames
//   dropout_global            infer_model.py:209,216     elementwise mask * 1/(1-p)
//   dpi_nl / point_ln / nl_emb_1 / hier_LN (+ReLU, + residual)   infer_model.py:185-193, 203-216
//   hidden_weight, point_conv, fc_emb_1                  nn.Linear / 1x1 conv
//   Dynamic_Person_Inference.dynamic_infer_ratio         infer_module/dynamic_infer_module.py:184-282
//       gradient reaches x through the four corner gathers and through p_conv / scale_conv; reaches the offsets
//       only through the bilinear weights (floor is taken on detached data, :208); torch.clamp passes gradient
//       where min <= p <= max; |.| has derivative sign(.) with sign(0) = 0.
// All fp32.  Parameter gradients are reduced in a fixed order (deterministic); only the scatter of the dynamic
// walk's corner gradients into dx uses fp32 atomics.
#include <cfloat>

#include "din_common.cuh"
#include "din_head.cuh"

namespace {

using namespace din;

// ================================================================================================
// General fp32 GEMM with arbitrary operand strides:  C[m,n] (+)= alpha * sum_k A(m,k) * B(k,n)
//   A(m,k) = a[m*sam + k*sak]  (fp32),  B(k,n) = b[k*sbk + n*sbn]  (fp32 or fp16)
// 64x64 tile, K step 16, 256 threads x (4x4) outputs.  Serves dX = dY.W, dW = dY^T.X of the head's linears.
// ================================================================================================
constexpr int kGBM = 64, kGBN = 64, kGBK = 16;

__device__ __forceinline__ float ld_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_as_float(const __half* p) { return __half2float(__ldg(p)); }

template <typename TB>
__global__ void __launch_bounds__(256)
gemm_strided_kernel(const float* __restrict__ a, long long sam, long long sak, const TB* __restrict__ b,
                    long long sbk, long long sbn, float* __restrict__ c, long long ldc, int M, int N, int K,
                    float alpha, int accumulate) {
  __shared__ float As[kGBK][kGBM + 4];
  __shared__ float Bs[kGBK][kGBN + 4];
  const int m0 = blockIdx.y * kGBM, n0 = blockIdx.x * kGBN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const bool a_k_fast = (sak == 1), b_k_fast = (sbk == 1);
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += kGBK) {
    float av[4], bv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + 256 * i;
      const int ka = a_k_fast ? (idx & 15) : (idx >> 6);
      const int ma = a_k_fast ? (idx >> 4) : (idx & 63);
      av[i] = (m0 + ma < M && k0 + ka < K) ? __ldg(a + static_cast<long long>(m0 + ma) * sam + static_cast<long long>(k0 + ka) * sak) : 0.0f;
      const int kb = b_k_fast ? (idx & 15) : (idx >> 6);
      const int nb = b_k_fast ? (idx >> 4) : (idx & 63);
      bv[i] = (n0 + nb < N && k0 + kb < K) ? ld_as_float(b + static_cast<long long>(k0 + kb) * sbk + static_cast<long long>(n0 + nb) * sbn) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + 256 * i;
      As[a_k_fast ? (idx & 15) : (idx >> 6)][a_k_fast ? (idx >> 4) : (idx & 63)] = av[i];
      Bs[b_k_fast ? (idx & 15) : (idx >> 6)][b_k_fast ? (idx >> 4) : (idx & 63)] = bv[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kGBK; ++k) {
      const float4 x4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 y4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float xa[4] = {x4.x, x4.y, x4.z, x4.w};
      const float yb[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], yb[j], acc[i][j]);
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
      float* dst = c + static_cast<long long>(m) * ldc + n;
      float v = alpha * acc[i][j];
      if (accumulate) v += *dst;
      *dst = v;
    }
  }
}

// y[n] = sum_m x[m*ld + n]   (bias gradients), one thread per column, fixed summation order
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, float* __restrict__ y, int M, int N, long long ld) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  float s = 0.0f;
  for (int m = 0; m < M; ++m) s += __ldg(x + static_cast<long long>(m) * ld + n);
  y[n] = s;
}

// y = x * mask * scale   (dropout forward and backward; mask is 0/1 bytes, NULL = all ones)
__global__ void __launch_bounds__(256)
scale_mask_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, float scale, float* __restrict__ y,
                  long long count) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= count) return;
  const float m = (mask == nullptr || mask[i]) ? scale : 0.0f;
  y[i] = x[i] * m;
}

// dz = dy * [y > 0]  (fp32 ReLU backward from the saved OUTPUT, as F.relu's autograd; base_model.py:120)
__global__ void __launch_bounds__(256)
relu_bwd_f32_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dz, long long count) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= count) return;
  dz[i] = y[i] > 0.0f ? dy[i] : 0.0f;
}

// ================================================================================================
// read-out backward.  Forward: pooled = max_n s; scores = pooled.W^T + bias; logits = mean_t scores.
//   kernel 1 (CTA per (clip, frame)): pooled + first arg-max (torch.max's index), dpooled = (dlogits/T).W,
//            ds = dpooled at the arg-max actor, 0 elsewhere;
//   kernel 2 (thread per (a, c)): dW[a,c] = sum_b dlogits[b,a]/T * sum_t pooled[b,t,c];  dbias[a] = sum_b dlogits[b,a].
// ================================================================================================
__global__ void __launch_bounds__(256)
readout_bwd_kernel(const float* __restrict__ s, const float* __restrict__ w, const float* __restrict__ dlogits,
                   float* __restrict__ ds, float* __restrict__ pooled_ws, int T, int N, int C, int A,
                   const int* __restrict__ n_valid) {
  __shared__ float dsc[64];
  const int bt = blockIdx.x;
  const int b = bt / T;
  const int Nb = n_valid ? min(__ldg(n_valid + b), N) : N;
  if (threadIdx.x < A) dsc[threadIdx.x] = __ldg(dlogits + static_cast<size_t>(b) * A + threadIdx.x) / static_cast<float>(T);
  __syncthreads();
  const float* st = s + static_cast<size_t>(bt) * N * C;
  float* dst = ds + static_cast<size_t>(bt) * N * C;
  for (int c = threadIdx.x; c < C; c += 256) {
    float m = -FLT_MAX;
    int arg = 0;
    for (int n = 0; n < Nb; ++n) {
      const float v = __ldg(st + static_cast<size_t>(n) * C + c);
      if (v > m) { m = v; arg = n; }
    }
    float dp = 0.0f;
    for (int a = 0; a < A; ++a) dp = fmaf(dsc[a], __ldg(w + static_cast<size_t>(a) * C + c), dp);
    for (int n = 0; n < N; ++n) dst[static_cast<size_t>(n) * C + c] = (n == arg && n < Nb) ? dp : 0.0f;
    pooled_ws[static_cast<size_t>(bt) * C + c] = m;
  }
}

__global__ void __launch_bounds__(256)
readout_wgrad_kernel(const float* __restrict__ pooled_ws, const float* __restrict__ dlogits, float* __restrict__ dw,
                     float* __restrict__ dbias, int B, int T, int C, int A) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= A * C) return;
  const int a = idx / C, c = idx - a * C;
  float acc = 0.0f, bacc = 0.0f;
  for (int b = 0; b < B; ++b) {
    const float g = __ldg(dlogits + static_cast<size_t>(b) * A + a);
    float ps = 0.0f;
    for (int t = 0; t < T; ++t) ps += __ldg(pooled_ws + (static_cast<size_t>(b) * T + t) * C + c);
    acc = fmaf(g / static_cast<float>(T), ps, acc);
    bacc += g;
  }
  dw[idx] = acc;
  if (c == 0) dbias[a] = bacc;
}

// ================================================================================================
// LayerNorm-over-strided-groups backward (same group geometry as group_layernorm_kernel in head.cu).
// Forward:  u = x (+ pre);  n = (u - mean) * rstd;  v = n*gamma + beta;  y = [relu](v) (+ post).
// Given dy:  dv = dy * [v > 0];  dn = dv*gamma;  du = rstd * (dn - mean(dn) - n * mean(dn*n));
//            dx (+)= du (the same du is the gradient of `pre`; dy itself is the gradient of `post`).
// One CTA per group; mean / rstd are recomputed with the forward's arithmetic (so the ReLU mask is the
// forward's) and written to stats_ws[g] = (mean, rstd) for the parameter-gradient kernel.
// ================================================================================================
constexpr int kLnbThreads = 512;

__global__ void __launch_bounds__(kLnbThreads)
group_layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ pre,
                           const float* __restrict__ gamma, const float* __restrict__ beta,
                           const float* __restrict__ dy, float* __restrict__ dx, float* __restrict__ stats_ws,
                           int n_inner, long long outer_stride, long long inner_stride, int rows,
                           long long row_stride, int cols, float eps, int relu, int accumulate,
                           const int* __restrict__ n_valid) {
  __shared__ float red[33];
  const int g = blockIdx.x;
  const int go = g / n_inner, gi = g - go * n_inner;
  if (n_valid != nullptr && gi >= __ldg(n_valid + go)) return;
  const size_t base = static_cast<size_t>(go) * outer_stride + static_cast<size_t>(gi) * inner_stride;
  const int total = rows * cols;
  const float n_el = static_cast<float>(rows) * static_cast<float>(cols);
  auto off_of = [&](int i) -> size_t {
    const int r = i / cols, c = i - r * cols;
    return base + static_cast<size_t>(r) * row_stride + c;
  };
  auto u_at = [&](size_t off) -> float {
    float v = __ldg(x + off);
    if (pre != nullptr) v += __ldg(pre + off);
    return v;
  };
  // the forward sums float4 lanes as (x+y)+(z+w) per thread; the statistics only need to agree to rounding,
  // the ReLU mask is evaluated on v recomputed below with the same mean / rstd this kernel uses for du
  float s = 0.0f;
  for (int i = threadIdx.x; i < total; i += kLnbThreads) s += u_at(off_of(i));
  const float mean = block_sum<kLnbThreads>(s, red) / n_el;
  float q = 0.0f;
  for (int i = threadIdx.x; i < total; i += kLnbThreads) {
    const float d = u_at(off_of(i)) - mean;
    q += d * d;
  }
  const float var = block_sum<kLnbThreads>(q, red) / n_el;
  const float rstd = rsqrtf(var + eps);
  if (threadIdx.x == 0) { stats_ws[2 * g] = mean; stats_ws[2 * g + 1] = rstd; }
  float s1 = 0.0f, s2 = 0.0f;
  for (int i = threadIdx.x; i < total; i += kLnbThreads) {
    const size_t off = off_of(i);
    const float nrm = (u_at(off) - mean) * rstd;
    const float ga = __ldg(gamma + i);
    float dv = __ldg(dy + off);
    if (relu && !(nrm * ga + __ldg(beta + i) > 0.0f)) dv = 0.0f;
    const float dn = dv * ga;
    s1 += dn;
    s2 += dn * nrm;
  }
  const float m1 = block_sum<kLnbThreads>(s1, red) / n_el;
  const float m2 = block_sum<kLnbThreads>(s2, red) / n_el;
  for (int i = threadIdx.x; i < total; i += kLnbThreads) {
    const size_t off = off_of(i);
    const float nrm = (u_at(off) - mean) * rstd;
    const float ga = __ldg(gamma + i);
    float dv = __ldg(dy + off);
    if (relu && !(nrm * ga + __ldg(beta + i) > 0.0f)) dv = 0.0f;
    float du = rstd * (dv * ga - m1 - nrm * m2);
    if (accumulate) du += dx[off];
    dx[off] = du;
  }
}

// dgamma[p] = sum_g dv*n,  dbeta[p] = sum_g dv   (p = r*cols + c; one thread per parameter element, groups in order)
__global__ void __launch_bounds__(256)
layernorm_param_grad_kernel(const float* __restrict__ x, const float* __restrict__ pre,
                            const float* __restrict__ gamma, const float* __restrict__ beta,
                            const float* __restrict__ dy, const float* __restrict__ stats_ws,
                            float* __restrict__ dgamma, float* __restrict__ dbeta, int n_outer, int n_inner,
                            long long outer_stride, long long inner_stride, int rows, long long row_stride, int cols,
                            int relu, const int* __restrict__ n_valid) {
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= rows * cols) return;
  const int r = p / cols, c = p - r * cols;
  const float ga = __ldg(gamma + p), be = __ldg(beta + p);
  float dg = 0.0f, db = 0.0f;
  for (int go = 0; go < n_outer; ++go) {
    const int ni = n_valid ? min(__ldg(n_valid + go), n_inner) : n_inner;
    for (int gi = 0; gi < ni; ++gi) {
      const int g = go * n_inner + gi;
      const size_t off = static_cast<size_t>(go) * outer_stride + static_cast<size_t>(gi) * inner_stride +
                         static_cast<size_t>(r) * row_stride + c;
      float u = __ldg(x + off);
      if (pre != nullptr) u += __ldg(pre + off);
      const float nrm = (u - __ldg(stats_ws + 2 * g)) * __ldg(stats_ws + 2 * g + 1);
      float dv = __ldg(dy + off);
      if (relu && !(nrm * ga + be > 0.0f)) dv = 0.0f;
      dg = fmaf(dv, nrm, dg);
      db += dv;
    }
  }
  dgamma[p] = dg;
  dbeta[p] = db;
}

// ================================================================================================
// Dynamic Relation + Dynamic Walk backward for one sampling ratio.
//   kernel A (CTA per (clip, frame), as the forward): recomputes the affinity convs, relation softmax and the
//     walk positions; per actor (one warp) and tap: the four <dy, corner> dot products give d(relation) and,
//     through the bilinear weights, d(offsets); coef*rel*w_corner*dy is scattered into dx at the corners.
//     Writes dconv[b,t,n,0..n_out) (gradient of the conv outputs) and the node's <dy, out> for d(coef).
//   kernel B (CTA per (clip, frame)): dx += conv^T(dconv) (exclusive ownership, no atomics).
//   kernel C (CTA per (tap, 128 channels)): dW_tap = sum_nodes dconv (x) x~, plus db and d(coef) in block (0,0).
// ================================================================================================
__device__ __forceinline__ float sgn(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }

__global__ void __launch_bounds__(kDinThreads)
din_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w_tap, const float* __restrict__ b_cat,
               const float* __restrict__ dy, float* __restrict__ dx, float* __restrict__ dconv,
               float* __restrict__ dcoef_part, int T, int N, int C, int kt, int kn, int ratio, int scale_factor,
               const float* __restrict__ coef_ptr, float coef_scalar, const int* __restrict__ n_valid) {
  extern __shared__ float smem_f[];
  const int b = blockIdx.x / T;
  const int t = blockIdx.x - b * T;
  const int Nb = n_valid ? min(__ldg(n_valid + b), N) : N;
  const int k2 = kt * kn;
  const int n_out = scale_factor ? 3 * k2 : 2 * k2;
  const int pt = (kt - 1) / 2 * ratio, pl = (kn - 1) / 2 * ratio;
  const int dy0 = -(((kt - 1) * ratio + 1) / 2);
  const int dx0 = -(((kn - 1) * ratio + 1) / 2);
  float* xs = smem_f;
  float* conv_s = smem_f + static_cast<size_t>(kt) * N * C;
  const float* xb = x + static_cast<size_t>(b) * T * N * C;
  float* dxb = dx + static_cast<size_t>(b) * T * N * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C4 = C >> 2;

  din_stage_rows_and_conv(xb, w_tap, b_cat, xs, conv_s, t, T, N, Nb, C, kt, kn, ratio, n_out, dy0, dx0);

  const float coef = coef_ptr ? __ldg(coef_ptr) : coef_scalar;
  const int Hp = T + 2 * pt;
  const int Wp_b = Nb + 2 * pl;
  const float my = static_cast<float>(Hp - 1), mxx = static_cast<float>(Wp_b - 1);
  for (int n = warp; n < N; n += kDinWarps) {
    const size_t node = (static_cast<size_t>(b) * T + t) * N + n;
    if (n >= Nb) {                                   // padded actor: no gradient
      if (lane < n_out) dconv[node * n_out + lane] = 0.0f;
      if (lane == 0 && dcoef_part != nullptr) dcoef_part[node] = 0.0f;
      continue;
    }
    const float* cs = conv_s + n * n_out;
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
    const float4* dy4 = reinterpret_cast<const float4*>(dy + node * C);
    float fdot[kDinMaxK2], dpy[kDinMaxK2], dpx[kDinMaxK2];
    for (int ky = 0; ky < kt; ++ky) {
      for (int kx = 0; kx < kn; ++kx) {
        const int k = ky * kn + kx;
        const float py_raw = static_cast<float>(pt + t + dy0 + ky * ratio) + cs[k];
        const float px_raw = static_cast<float>(pl + n + dx0 + kx * ratio) + cs[k2 + k];
        float ly = floorf(py_raw), lx = floorf(px_raw);
        float ry = ly + 1.0f, rx = lx + 1.0f;
        ly = fminf(fmaxf(ly, 0.0f), my); ry = fminf(fmaxf(ry, 0.0f), my);
        lx = fminf(fmaxf(lx, 0.0f), mxx); rx = fminf(fmaxf(rx, 0.0f), mxx);
        const float py = fminf(fmaxf(py_raw, 0.0f), my), px = fminf(fmaxf(px_raw, 0.0f), mxx);
        const float pass_y = (py_raw >= 0.0f && py_raw <= my) ? 1.0f : 0.0f;     // torch.clamp backward mask
        const float pass_x = (px_raw >= 0.0f && px_raw <= mxx) ? 1.0f : 0.0f;
        const float wly = 1.0f - fabsf(py - ly), wry = 1.0f - fabsf(py - ry);
        const float wlx = 1.0f - fabsf(px - lx), wrx = 1.0f - fabsf(px - rx);
        const int ily = static_cast<int>(ly) - pt, iry = static_cast<int>(ry) - pt;
        const int ilx = static_cast<int>(lx) - pl, irx = static_cast<int>(rx) - pl;
        const bool ok_lt = ily >= 0 && ily < T && ilx >= 0 && ilx < Nb;
        const bool ok_rb = iry >= 0 && iry < T && irx >= 0 && irx < Nb;
        const bool ok_lb = iry >= 0 && iry < T && ilx >= 0 && ilx < Nb;
        const bool ok_rt = ily >= 0 && ily < T && irx >= 0 && irx < Nb;
        const size_t o_lt = (static_cast<size_t>(ily) * N + ilx) * C, o_rb = (static_cast<size_t>(iry) * N + irx) * C;
        const size_t o_lb = (static_cast<size_t>(iry) * N + ilx) * C, o_rt = (static_cast<size_t>(ily) * N + irx) * C;
        const float wlt = wly * wlx, wrb = wry * wrx, wlb = wry * wlx, wrt = wly * wrx;
        const float a = coef * rel[k];
        float d_lt = 0.0f, d_rb = 0.0f, d_lb = 0.0f, d_rt = 0.0f;
        for (int c4 = lane; c4 < C4; c4 += 32) {
          const float4 g = __ldg(dy4 + c4);
          auto corner = [&](bool ok, size_t off, float wgt, float& dot) {
            if (!ok) return;
            const float4 v = __ldg(reinterpret_cast<const float4*>(xb + off) + c4);
            dot += (g.x * v.x + g.y * v.y) + (g.z * v.z + g.w * v.w);
            const float s = a * wgt;
            if (s != 0.0f) {
              float* d = dxb + off + 4 * c4;
              atomicAdd(d + 0, s * g.x); atomicAdd(d + 1, s * g.y); atomicAdd(d + 2, s * g.z); atomicAdd(d + 3, s * g.w);
            }
          };
          corner(ok_lt, o_lt, wlt, d_lt);
          corner(ok_rb, o_rb, wrb, d_rb);
          corner(ok_lb, o_lb, wlb, d_lb);
          corner(ok_rt, o_rt, wrt, d_rt);
        }
        d_lt = warp_sum(d_lt); d_rb = warp_sum(d_rb); d_lb = warp_sum(d_lb); d_rt = warp_sum(d_rt);
        fdot[k] = ((d_lt * wlt + d_rb * wrb) + d_lb * wlb) + d_rt * wrt;         // <dy, f_k>
        const float c_lt = a * d_lt, c_rb = a * d_rb, c_lb = a * d_lb, c_rt = a * d_rt;   // d(loss)/d(coe_corner)
        const float sy_l = sgn(py - ly), sy_r = sgn(py - ry), sx_l = sgn(px - lx), sx_r = sgn(px - rx);
        dpy[k] = -pass_y * (((c_lt * sy_l * wlx + c_rb * sy_r * wrx) + c_lb * sy_r * wlx) + c_rt * sy_l * wrx);
        dpx[k] = -pass_x * (((c_lt * sx_l * wly + c_rb * sx_r * wry) + c_lb * sx_l * wry) + c_rt * sx_r * wly);
      }
    }
    float out_dot = 0.0f;                                                         // <dy, sum_k rel_k f_k>
    for (int k = 0; k < k2; ++k) out_dot = fmaf(rel[k], fdot[k], out_dot);
    if (lane == 0) {
      float* dc = dconv + node * n_out;
      for (int k = 0; k < k2; ++k) {
        dc[k] = dpy[k];
        dc[k2 + k] = dpx[k];
        if (scale_factor) dc[2 * k2 + k] = coef * rel[k] * (fdot[k] - out_dot);   // softmax backward
      }
      if (dcoef_part != nullptr) dcoef_part[node] = out_dot;
    }
  }
}

// dx[b,t',n',:] += sum_tap sum_o dconv[b, t'-dy_tap, n'-dx_tap, o] * W[tap][o][:]
__global__ void __launch_bounds__(256)
din_conv_bwd_dx_kernel(const float* __restrict__ dconv, const float* __restrict__ w_tap, float* __restrict__ dx,
                       int T, int N, int C, int kt, int kn, int ratio, int n_out, const int* __restrict__ n_valid) {
  extern __shared__ float dcs[];                    // [kt][N][n_out] source rows of dconv (zero where invalid)
  const int b = blockIdx.x / T;
  const int t = blockIdx.x - b * T;
  const int Nb = n_valid ? min(__ldg(n_valid + b), N) : N;
  const int dy0 = -(((kt - 1) * ratio + 1) / 2);
  const int dx0 = -(((kn - 1) * ratio + 1) / 2);
  for (int i = threadIdx.x; i < kt * N * n_out; i += 256) {
    const int o = i % n_out;
    const int rn = i / n_out;
    const int n = rn % N, ky = rn / N;
    const int ts = t - (dy0 + ky * ratio);          // the node whose tap ky reads row t
    float v = 0.0f;
    if (ts >= 0 && ts < T && n < Nb) v = __ldg(dconv + ((static_cast<size_t>(b) * T + ts) * N + n) * n_out + o);
    dcs[i] = v;
  }
  __syncthreads();
  // thread = (actor slice, channel quad): with C = 128 only 32 quads exist, so the 256 threads also split the
  // actors (slice s owns actors s, s + n_slices, ...); with C = 1024 there is one slice and 16 accumulators
  const int C4 = C >> 2;
  const int n_slices = (C4 >= 256) ? 1 : (256 / C4 < kDinMaxN ? 256 / C4 : kDinMaxN);
  for (int item = threadIdx.x; item < C4 * n_slices; item += 256) {
    const int c4 = item % C4, slice = item / C4;
    float4 acc[kDinMaxN];
#pragma unroll
    for (int j = 0; j < kDinMaxN; ++j) acc[j] = make_float4(0, 0, 0, 0);
    for (int ky = 0; ky < kt; ++ky) {
      for (int kx = 0; kx < kn; ++kx) {
        const int tap = ky * kn + kx;
        const int dxk = dx0 + kx * ratio;
        for (int o = 0; o < n_out; ++o) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(w_tap + (static_cast<size_t>(tap) * n_out + o) * C) + c4);
#pragma unroll
          for (int j = 0; j < kDinMaxN; ++j) {
            const int n = slice + j * n_slices;
            const int ns = n - dxk;                  // source actor
            if (n < Nb && ns >= 0 && ns < Nb) {
              const float d = dcs[(ky * N + ns) * n_out + o];
              acc[j].x = fmaf(d, wv.x, acc[j].x); acc[j].y = fmaf(d, wv.y, acc[j].y);
              acc[j].z = fmaf(d, wv.z, acc[j].z); acc[j].w = fmaf(d, wv.w, acc[j].w);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kDinMaxN; ++j) {
      const int n = slice + j * n_slices;
      if (n < Nb) {
        float4* dst = reinterpret_cast<float4*>(dx + ((static_cast<size_t>(b) * T + t) * N + n) * C) + c4;
        float4 v = *dst;
        v.x += acc[j].x; v.y += acc[j].y; v.z += acc[j].z; v.w += acc[j].w;
        *dst = v;
      }
    }
  }
}

// dW[tap][o][c] = sum_{b,t,n} dconv[b,t,n,o] * x~[b, t+dy_tap, n+dx_tap, c], in two deterministic stages:
//   stage 1  grid (tap, 128-channel chunk, node split): partial sums over a contiguous range of nodes -> part[s]
//   stage 2  dW = sum_s part[s] (fixed order); one extra block reduces db and d(coef)
// (a single pass over all 960 nodes per block was a 0.6 ms latency chain at grid 9-72, ncu r1_v9)
constexpr int kDwThreads = 128;
constexpr int kDwNodes = 32;
constexpr int kDwMaxOut = 27;

__global__ void __launch_bounds__(kDwThreads)
din_conv_bwd_dw_partial_kernel(const float* __restrict__ x, const float* __restrict__ dconv, float* __restrict__ part,
                               int B, int T, int N, int C, int kt, int kn, int ratio, int n_out, int nodes_per_split,
                               const int* __restrict__ n_valid) {
  __shared__ float dcs[kDwNodes][kDwMaxOut + 1];
  const int tap = blockIdx.x;
  const int ky = tap / kn, kx = tap - ky * kn;
  const int dyk = -(((kt - 1) * ratio + 1) / 2) + ky * ratio;
  const int dxk = -(((kn - 1) * ratio + 1) / 2) + kx * ratio;
  const int c = blockIdx.y * kDwThreads + threadIdx.x;
  const int nodes = B * T * N;
  const int node_begin = blockIdx.z * nodes_per_split;
  const int node_end = min(nodes, node_begin + nodes_per_split);
  float acc[kDwMaxOut];
#pragma unroll
  for (int o = 0; o < kDwMaxOut; ++o) acc[o] = 0.0f;
  for (int n0 = node_begin; n0 < node_end; n0 += kDwNodes) {
    __syncthreads();
    for (int i = threadIdx.x; i < kDwNodes * n_out; i += kDwThreads) {
      const int j = i / n_out, o = i - j * n_out;
      dcs[j][o] = (n0 + j < node_end) ? __ldg(dconv + static_cast<size_t>(n0 + j) * n_out + o) : 0.0f;
    }
    __syncthreads();
    if (c < C) {
      for (int j = 0; j < kDwNodes && n0 + j < node_end; ++j) {
        const int node = n0 + j;
        const int n = node % N;
        const int bt = node / N;
        const int t = bt % T, b = bt / T;
        const int Nb = n_valid ? min(__ldg(n_valid + b), N) : N;
        const int ts = t + dyk, ns = n + dxk;
        if (n >= Nb || ts < 0 || ts >= T || ns < 0 || ns >= Nb) continue;
        const float xv = __ldg(x + ((static_cast<size_t>(b) * T + ts) * N + ns) * C + c);
#pragma unroll
        for (int o = 0; o < kDwMaxOut; ++o)
          if (o < n_out) acc[o] = fmaf(dcs[j][o], xv, acc[o]);
      }
    }
  }
  if (c < C) {
    float* dst = part + (static_cast<size_t>(blockIdx.z) * gridDim.x + tap) * n_out * C;
#pragma unroll
    for (int o = 0; o < kDwMaxOut; ++o)
      if (o < n_out) dst[static_cast<size_t>(o) * C + c] = acc[o];
  }
}

__global__ void __launch_bounds__(256)
din_conv_bwd_dw_reduce_kernel(const float* __restrict__ part, const float* __restrict__ dconv,
                              const float* __restrict__ dcoef_part, float* __restrict__ dw_tap,
                              float* __restrict__ db_cat, float* __restrict__ dcoef, int total /* k2*n_out*C */,
                              int splits, int nodes, int n_out) {
  __shared__ float red[33];
  if (blockIdx.x < gridDim.x - 1) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < total) {
      float s = 0.0f;
      for (int sp = 0; sp < splits; ++sp) s += __ldg(part + static_cast<size_t>(sp) * total + i);
      dw_tap[i] = s;
    }
    return;
  }
  // last block: bias gradient (padded actors hold zeros in dconv) and d(coef)
  if (threadIdx.x < n_out) {
    float s = 0.0f;
    for (int node = 0; node < nodes; ++node) s += __ldg(dconv + static_cast<size_t>(node) * n_out + threadIdx.x);
    db_cat[threadIdx.x] = s;
  }
  if (dcoef != nullptr) {
    float s = 0.0f;
    for (int node = threadIdx.x; node < nodes; node += 256) s += __ldg(dcoef_part + node);
    s = block_sum<256>(s, red);
    if (threadIdx.x == 0) dcoef[0] = s;
  }
}

// node splits of stage 1 and the workspace that follows from them
__host__ inline int din_bwd_splits(int nodes, int k2, int c) {
  const int blocks = k2 * ((c + kDwThreads - 1) / kDwThreads);
  int splits = (2 * 148 + blocks - 1) / blocks;
  const int max_splits = (nodes + kDwNodes - 1) / kDwNodes;
  if (splits > max_splits) splits = max_splits;
  return splits < 1 ? 1 : splits;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" int din_gemm_f32(const float* a, long long a_stride_m, long long a_stride_k, const void* b, int b_is_f16,
                            long long b_stride_k, long long b_stride_n, float* c, long long ldc, int m, int n, int k,
                            float alpha, int accumulate, void* stream) {
  DIN_CHECK_ARG(a && b && c, "din_gemm_f32: null pointer");
  DIN_CHECK_ARG(m > 0 && n > 0 && k > 0 && ldc >= n, "din_gemm_f32: bad shape m=%d n=%d k=%d ldc=%lld", m, n, k, ldc);
  dim3 grid((n + kGBN - 1) / kGBN, (m + kGBM - 1) / kGBM);
  DIN_CHECK_ARG(grid.y <= 65535, "din_gemm_f32: m=%d too large", m);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (b_is_f16)
    gemm_strided_kernel<__half><<<grid, 256, 0, st>>>(a, a_stride_m, a_stride_k, static_cast<const __half*>(b),
                                                      b_stride_k, b_stride_n, c, ldc, m, n, k, alpha, accumulate);
  else
    gemm_strided_kernel<float><<<grid, 256, 0, st>>>(a, a_stride_m, a_stride_k, static_cast<const float*>(b),
                                                     b_stride_k, b_stride_n, c, ldc, m, n, k, alpha, accumulate);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_colsum_f32(const float* x, float* y, int m, int n, long long ld, void* stream) {
  DIN_CHECK_ARG(x && y, "din_colsum_f32: null pointer");
  DIN_CHECK_ARG(m > 0 && n > 0 && ld >= n, "din_colsum_f32: bad shape m=%d n=%d ld=%lld", m, n, ld);
  colsum_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, m, n, ld);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_scale_mask_f32(const float* x, const uint8_t* mask, float scale, float* y, long long count,
                                  void* stream) {
  DIN_CHECK_ARG(x && y, "din_scale_mask_f32: null pointer");
  DIN_CHECK_ARG(count > 0 && (count + 255) / 256 <= INT32_MAX, "din_scale_mask_f32: bad count %lld", count);
  scale_mask_kernel<<<static_cast<int>((count + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, mask, scale, y, count);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_relu_bwd_f32(const float* y, const float* dy, float* dz, long long count, void* stream) {
  DIN_CHECK_ARG(y && dy && dz, "din_relu_bwd_f32: null pointer");
  DIN_CHECK_ARG(count > 0 && (count + 255) / 256 <= INT32_MAX, "din_relu_bwd_f32: bad count %lld", count);
  relu_bwd_f32_kernel<<<static_cast<int>((count + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, dy, dz,
                                                                                                            count);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_readout_bwd_f32(const float* s, const float* w, const float* dlogits, float* ds, float* pooled_ws,
                                   float* dw, float* dbias, int b, int t, int n, int c, int a,
                                   const int32_t* n_valid, void* stream) {
  DIN_CHECK_ARG(s && w && dlogits && ds && pooled_ws && dw && dbias, "din_readout_bwd_f32: null pointer");
  DIN_CHECK_ARG(b > 0 && t > 0 && n > 0 && c > 0 && a > 0 && a <= 64,
                "din_readout_bwd_f32: bad shape b=%d t=%d n=%d c=%d a=%d (a <= 64)", b, t, n, c, a);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  readout_bwd_kernel<<<b * t, 256, 0, st>>>(s, w, dlogits, ds, pooled_ws, t, n, c, a, n_valid);
  DIN_CHECK_CUDA(cudaGetLastError());
  readout_wgrad_kernel<<<(a * c + 255) / 256, 256, 0, st>>>(pooled_ws, dlogits, dw, dbias, b, t, c, a);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_group_layernorm_bwd_f32(const float* x, const float* pre, const float* gamma, const float* beta,
                                           const float* dy, float* dx, float* dgamma, float* dbeta, float* stats_ws,
                                           int n_outer, int n_inner, long long outer_stride, long long inner_stride,
                                           int rows, long long row_stride, int cols, float eps, int relu,
                                           int accumulate_dx, const int32_t* n_valid, void* stream) {
  DIN_CHECK_ARG(x && gamma && beta && dy && dx && stats_ws, "din_group_layernorm_bwd_f32: null pointer");
  DIN_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr), "din_group_layernorm_bwd_f32: dgamma and dbeta go together");
  DIN_CHECK_ARG(n_outer > 0 && n_inner > 0 && rows > 0 && cols > 0,
                "din_group_layernorm_bwd_f32: bad shape outer=%d inner=%d rows=%d cols=%d", n_outer, n_inner, rows, cols);
  DIN_CHECK_ARG(static_cast<long long>(rows) * cols <= INT32_MAX, "din_group_layernorm_bwd_f32: group too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  group_layernorm_bwd_kernel<<<n_outer * n_inner, kLnbThreads, 0, st>>>(
      x, pre, gamma, beta, dy, dx, stats_ws, n_inner, outer_stride, inner_stride, rows, row_stride, cols, eps, relu,
      accumulate_dx, n_valid);
  DIN_CHECK_CUDA(cudaGetLastError());
  if (dgamma != nullptr) {
    layernorm_param_grad_kernel<<<(rows * cols + 255) / 256, 256, 0, st>>>(
        x, pre, gamma, beta, dy, stats_ws, dgamma, dbeta, n_outer, n_inner, outer_stride, inner_stride, rows,
        row_stride, cols, relu, n_valid);
    DIN_CHECK_CUDA(cudaGetLastError());
  }
  return DIN_OK;
}

extern "C" int din_dynamic_infer_bwd_f32(const float* x, const float* w_tap, const float* b_cat, const float* dy,
                                         float* dx, float* dw_tap, float* db_cat, float* dcoef, float* ws, int b,
                                         int t, int n, int c, int kt, int kn, int ratio, int scale_factor,
                                         const float* coef_ptr, float coef_scalar, const int32_t* n_valid,
                                         void* stream) {
  DIN_CHECK_ARG(x && w_tap && b_cat && dy && dx && dw_tap && db_cat && ws, "din_dynamic_infer_bwd_f32: null pointer");
  DIN_CHECK_ARG(b > 0 && t > 0 && n > 0 && n <= kDinMaxN, "din_dynamic_infer_bwd_f32: bad extent b=%d t=%d n=%d (n <= %d)",
                b, t, n, kDinMaxN);
  DIN_CHECK_ARG(c > 0 && c % 4 == 0, "din_dynamic_infer_bwd_f32: c=%d must be a multiple of 4", c);
  DIN_CHECK_ARG(kt >= 1 && kn >= 1 && kt * kn <= kDinMaxK2 && (kt & 1) && (kn & 1),
                "din_dynamic_infer_bwd_f32: kernel %dx%d unsupported (odd, <= %d taps)", kt, kn, kDinMaxK2);
  DIN_CHECK_ARG(ratio >= 1, "din_dynamic_infer_bwd_f32: ratio=%d", ratio);
  const int n_out = (scale_factor ? 3 : 2) * kt * kn;
  DIN_CHECK_ARG(n_out <= kDwMaxOut, "din_dynamic_infer_bwd_f32: too many conv outputs");
  DIN_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) |
                  reinterpret_cast<uintptr_t>(w_tap)) & 15) == 0,
                "din_dynamic_infer_bwd_f32: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t nodes = static_cast<size_t>(b) * t * n;
  float* dconv = ws;                          // [nodes][n_out]
  float* dcoef_part = ws + nodes * n_out;     // [nodes]
  const size_t smem = (static_cast<size_t>(kt) * n * c + static_cast<size_t>(n) * n_out) * sizeof(float);
  DIN_CHECK_ARG(smem <= 220 * 1024, "din_dynamic_infer_bwd_f32: kt*n*c too large for shared memory (%zu bytes)", smem);
  DIN_OPT_IN_SMEM(din_bwd_kernel, smem);
  din_bwd_kernel<<<b * t, kDinThreads, smem, st>>>(x, w_tap, b_cat, dy, dx, dconv, dcoef_part, t, n, c, kt, kn, ratio,
                                                   scale_factor, coef_ptr, coef_scalar, n_valid);
  DIN_CHECK_CUDA(cudaGetLastError());
  const size_t smem2 = static_cast<size_t>(kt) * n * n_out * sizeof(float);
  din_conv_bwd_dx_kernel<<<b * t, 256, smem2, st>>>(dconv, w_tap, dx, t, n, c, kt, kn, ratio, n_out, n_valid);
  DIN_CHECK_CUDA(cudaGetLastError());
  const int k2 = kt * kn;
  const int splits = din_bwd_splits(static_cast<int>(nodes), k2, c);
  const int nodes_per_split = (static_cast<int>(nodes) + splits - 1) / splits;
  float* part = dcoef_part + nodes;           // [splits][k2][n_out][c]
  dim3 grid(k2, (c + kDwThreads - 1) / kDwThreads, splits);
  din_conv_bwd_dw_partial_kernel<<<grid, kDwThreads, 0, st>>>(x, dconv, part, b, t, n, c, kt, kn, ratio, n_out,
                                                              nodes_per_split, n_valid);
  DIN_CHECK_CUDA(cudaGetLastError());
  const int total = k2 * n_out * c;
  din_conv_bwd_dw_reduce_kernel<<<(total + 255) / 256 + 1, 256, 0, st>>>(part, dconv, dcoef_part, dw_tap, db_cat, dcoef,
                                                                        total, splits, static_cast<int>(nodes), n_out);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" long long din_dynamic_infer_bwd_ws_floats(int b, int t, int n, int c, int kt, int kn, int scale_factor) {
  if (b <= 0 || t <= 0 || n <= 0 || c <= 0 || kt <= 0 || kn <= 0) return -1;
  const long long nodes = static_cast<long long>(b) * t * n;
  const int k2 = kt * kn, n_out = (scale_factor ? 3 : 2) * k2;
  return nodes * (n_out + 1) + static_cast<long long>(din_bwd_splits(static_cast<int>(nodes), k2, c)) * k2 * n_out * c;
}
