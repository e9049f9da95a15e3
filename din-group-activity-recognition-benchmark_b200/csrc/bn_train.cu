// bn_train.cu — BatchNorm2d on BATCH statistics (module.train()), forward and backward, for a backbone that is trained
// without `cfg.set_bn_eval` (config.py:80 default False; scripts/train_collective_stage2_dynamic.py:12-16 trains
// ResNet-18 that way: model.train() at train_net_dynamic.py:161 leaves every BatchNorm on batch statistics, the batch
// being all B*T frames of the step, backbone.py:115-132 called from infer_model.py:317-319).
//
// The convolution runs WITHOUT the folded scale / shift and writes the raw z (fp16, NHWC); then
//   stats    : per-channel sum and sum of squares over all rows (fp32 partials, fp64 atomics)
//   finalize : mean, biased variance, invstd; scale = gamma * invstd, shift = beta - mean * scale; the running statistics
//              are updated in place as nn.BatchNorm2d does (momentum, unbiased variance)
//   apply    : y = [relu]( z * scale + shift [+ residual] )
// and backward, with g = dY already masked by the ReLU and xhat = (z - mean) * invstd:
//   reduce   : sg = sum g, sgx = sum g * xhat
//   apply    : dz = gamma * invstd * (g - sg / N - xhat * sgx / N);  d(beta) += sg / S, d(gamma) += sgx / S  (S = loss scale)
// All HBM-bound streaming kernels: 16-byte vectors, >= 16 rows per thread before any atomics.
#include <cstdint>

#include <algorithm>

#include "din_common.cuh"

namespace {

using namespace din;

// eight consecutive channels of z (fp16 or fp32) as floats
__device__ __forceinline__ void load8(const __half* p, float (&f)[8]) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  const __half* h = reinterpret_cast<const __half*>(&v);
#pragma unroll
  for (int e = 0; e < 8; ++e) f[e] = __half2float(h[e]);
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

struct RowMap {
  int octs, lanes_r, oct, rl;
};

__device__ __forceinline__ RowMap row_map(int c) {
  RowMap m;
  m.octs = c >> 3;
  m.lanes_r = 256 / m.octs;
  m.oct = threadIdx.x % m.octs;
  m.rl = threadIdx.x / m.octs;
  return m;
}

// reduce two 8-vectors per thread over the row lanes of the block, then one atomic per channel and quantity
template <typename Acc>
__device__ __forceinline__ void block_reduce_pairs(const float (&a)[8], const float (&b)[8], const RowMap& m, Acc* out_a,
                                                   Acc* out_b) {
  __shared__ float red[256][16 + 1];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    red[threadIdx.x][e] = a[e];
    red[threadIdx.x][8 + e] = b[e];
  }
  __syncthreads();
  if (threadIdx.x < m.octs) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float sa = 0.0f, sb = 0.0f;
      for (int l = 0; l < m.lanes_r; ++l) {
        sa += red[l * m.octs + threadIdx.x][e];
        sb += red[l * m.octs + threadIdx.x][8 + e];
      }
      atomicAdd(out_a + threadIdx.x * 8 + e, static_cast<Acc>(sa));
      atomicAdd(out_b + threadIdx.x * 8 + e, static_cast<Acc>(sb));
    }
  }
}

template <typename ZT>
__global__ void __launch_bounds__(256)
bn_stats_kernel(const ZT* __restrict__ z, long long rows, int c, double* __restrict__ sum, double* __restrict__ sumsq) {
  const RowMap m = row_map(c);
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (m.rl < m.lanes_r) {
    for (long long r = static_cast<long long>(blockIdx.x) * m.lanes_r + m.rl; r < rows;
         r += static_cast<long long>(gridDim.x) * m.lanes_r) {
      float f[8];
      load8(z + r * c + m.oct * 8, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s[e] += f[e];
        q[e] = fmaf(f[e], f[e], q[e]);
      }
    }
  }
  block_reduce_pairs<double>(s, q, m, sum, sumsq);
}

// one thread per channel
__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, long long count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, int c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const double n = static_cast<double>(count);
  const double mean = sum[i] / n;
  double var = sumsq[i] / n - mean * mean;                 // biased, as F.batch_norm normalises with
  if (var < 0.0) var = 0.0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float sc = gamma[i] * invstd;
  scale[i] = sc;
  shift[i] = beta[i] - static_cast<float>(mean) * sc;
  mean_out[i] = static_cast<float>(mean);
  invstd_out[i] = invstd;
  if (running_mean != nullptr) {
    const double unbiased = count > 1 ? var * n / (n - 1.0) : var;
    running_mean[i] = (1.0f - momentum) * running_mean[i] + momentum * static_cast<float>(mean);
    running_var[i] = (1.0f - momentum) * running_var[i] + momentum * static_cast<float>(unbiased);
  }
}

template <typename ZT>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const ZT* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                const __half* __restrict__ residual, __half* __restrict__ y, long long count8, int c8, int relu) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= count8) return;
  const int oc = static_cast<int>(i % c8);
  float zf[8];
  load8(z + i * 8, zf);
  uint4 rv = make_uint4(0, 0, 0, 0);
  if (residual != nullptr) rv = __ldg(reinterpret_cast<const uint4*>(residual) + i);
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale) + 2 * oc), s1 = __ldg(reinterpret_cast<const float4*>(scale) + 2 * oc + 1);
  const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift) + 2 * oc), h1 = __ldg(reinterpret_cast<const float4*>(shift) + 2 * oc + 1);
  const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
  const __half* hr = reinterpret_cast<const __half*>(&rv);
  uint4 o;
  __half* ho = reinterpret_cast<__half*>(&o);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float f = fmaf(zf[e], sc[e], sh[e]) + __half2float(hr[e]);
    if (relu) f = fmaxf(f, 0.0f);
    ho[e] = __float2half_rn(f);
  }
  reinterpret_cast<uint4*>(y)[i] = o;
}

template <typename ZT>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const __half* __restrict__ g, const ZT* __restrict__ z, const float* __restrict__ mean,
                     const float* __restrict__ invstd, long long rows, int c, float* __restrict__ sg,
                     float* __restrict__ sgx) {
  const RowMap m = row_map(c);
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (m.rl < m.lanes_r) {
    float mu[8], is[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      mu[e] = __ldg(mean + m.oct * 8 + e);
      is[e] = __ldg(invstd + m.oct * 8 + e);
    }
    for (long long r = static_cast<long long>(blockIdx.x) * m.lanes_r + m.rl; r < rows;
         r += static_cast<long long>(gridDim.x) * m.lanes_r) {
      const uint4 gv = __ldg(reinterpret_cast<const uint4*>(g + r * c) + m.oct);
      float zf[8];
      load8(z + r * c + m.oct * 8, zf);
      const __half* hg = reinterpret_cast<const __half*>(&gv);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float gf = __half2float(hg[e]);
        a[e] += gf;
        b[e] = fmaf(gf, (zf[e] - mu[e]) * is[e], b[e]);
      }
    }
  }
  block_reduce_pairs<float>(a, b, m, sg, sgx);
}

template <typename ZT>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const __half* __restrict__ g, const ZT* __restrict__ z, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ sg,
                    const float* __restrict__ sgx, float inv_count, __half* __restrict__ dz, long long count8, int c8,
                    float* __restrict__ dbeta, float* __restrict__ dgamma, const float* __restrict__ inv_scale) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (blockIdx.x == 0 && dbeta != nullptr) {               // parameter gradients, once per call
    const float scl = inv_scale ? __ldg(inv_scale) : 1.0f;
    for (int ch = threadIdx.x; ch < c8 * 8; ch += 256) {
      dbeta[ch] += sg[ch] * scl;
      dgamma[ch] += sgx[ch] * scl;
    }
  }
  if (i >= count8) return;
  const int oc = static_cast<int>(i % c8);
  const uint4 gv = __ldg(reinterpret_cast<const uint4*>(g) + i);
  float zf[8];
  load8(z + i * 8, zf);
  const __half* hg = reinterpret_cast<const __half*>(&gv);
  uint4 o;
  __half* ho = reinterpret_cast<__half*>(&o);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int ch = oc * 8 + e;
    const float is = __ldg(invstd + ch);
    const float xh = (zf[e] - __ldg(mean + ch)) * is;
    const float v = __ldg(gamma + ch) * is * (__half2float(hg[e]) - __ldg(sg + ch) * inv_count - xh * __ldg(sgx + ch) * inv_count);
    ho[e] = __float2half_rn(v);
  }
  reinterpret_cast<uint4*>(dz)[i] = o;
}

int reduce_grid(long long rows, int c) {
  const int sms = din_num_sms();
  int g = 8 * (sms > 0 ? sms : 148);
  const long long per = 256 / (c / 8);
  if (static_cast<long long>(g) * per * 16 > rows) g = static_cast<int>(std::max<long long>(1, rows / (per * 16)));
  return g;
}

}  // namespace

extern "C" int din_bn_stats(const void* z, int z_is_f32, long long rows, int c, double* sum, double* sumsq,
                            void* stream) {
  DIN_CHECK_ARG(z && sum && sumsq, "din_bn_stats: null pointer");
  DIN_CHECK_ARG(rows > 0 && c > 0 && c % 8 == 0 && c <= 2048, "din_bn_stats: bad shape rows=%lld c=%d", rows, c);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DIN_CHECK_CUDA(cudaMemsetAsync(sum, 0, sizeof(double) * c, st));
  DIN_CHECK_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(double) * c, st));
  if (z_is_f32) bn_stats_kernel<float><<<reduce_grid(rows, c), 256, 0, st>>>(static_cast<const float*>(z), rows, c, sum, sumsq);
  else bn_stats_kernel<__half><<<reduce_grid(rows, c), 256, 0, st>>>(static_cast<const __half*>(z), rows, c, sum, sumsq);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_bn_finalize_f32(const double* sum, const double* sumsq, long long count, const float* gamma,
                                   const float* beta, float eps, float momentum, float* running_mean,
                                   float* running_var, float* scale, float* shift, float* mean, float* invstd, int c,
                                   void* stream) {
  DIN_CHECK_ARG(sum && sumsq && gamma && beta && scale && shift && mean && invstd, "din_bn_finalize_f32: null pointer");
  DIN_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "din_bn_finalize_f32: running_mean / running_var");
  DIN_CHECK_ARG(count > 0 && c > 0, "din_bn_finalize_f32: bad shape count=%lld c=%d", count, c);
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      sum, sumsq, count, gamma, beta, eps, momentum, running_mean, running_var, scale, shift, mean, invstd, c);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_bn_apply(const void* z, int z_is_f32, const float* scale, const float* shift, const void* residual,
                            void* y, long long rows, int c, int relu, void* stream) {
  DIN_CHECK_ARG(z && scale && shift && y, "din_bn_apply: null pointer");
  DIN_CHECK_ARG(rows > 0 && c > 0 && c % 8 == 0, "din_bn_apply: bad shape rows=%lld c=%d", rows, c);
  const long long count8 = rows * (c / 8);
  DIN_CHECK_ARG((count8 + 255) / 256 <= INT32_MAX, "din_bn_apply: too large");
  const int grid = static_cast<int>((count8 + 255) / 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (z_is_f32)
    bn_apply_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(z), scale, shift,
                                                 static_cast<const __half*>(residual), static_cast<__half*>(y), count8,
                                                 c / 8, relu);
  else
    bn_apply_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(z), scale, shift,
                                                  static_cast<const __half*>(residual), static_cast<__half*>(y), count8,
                                                  c / 8, relu);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_bn_bwd(const void* g, const void* z, int z_is_f32, const float* mean, const float* invstd, const float* gamma,
                              float* sums, void* dz, float* dbeta, float* dgamma, const float* inv_scale, long long rows,
                              int c, void* stream) {
  DIN_CHECK_ARG(g && z && mean && invstd && gamma && sums && dz, "din_bn_bwd: null pointer");
  DIN_CHECK_ARG((dbeta == nullptr) == (dgamma == nullptr), "din_bn_bwd: dbeta / dgamma");
  DIN_CHECK_ARG(rows > 0 && c > 0 && c % 8 == 0 && c <= 2048, "din_bn_bwd: bad shape rows=%lld c=%d", rows, c);
  const long long count8 = rows * (c / 8);
  DIN_CHECK_ARG((count8 + 255) / 256 <= INT32_MAX, "din_bn_bwd: too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DIN_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * c, st));
  const __half* gh = static_cast<const __half*>(g);
  const int grid = static_cast<int>((count8 + 255) / 256);
  const float inv_count = 1.0f / static_cast<float>(rows);
  if (z_is_f32) {
    const float* zf = static_cast<const float*>(z);
    bn_bwd_reduce_kernel<float><<<reduce_grid(rows, c), 256, 0, st>>>(gh, zf, mean, invstd, rows, c, sums, sums + c);
    DIN_CHECK_CUDA(cudaGetLastError());
    bn_bwd_apply_kernel<float><<<grid, 256, 0, st>>>(gh, zf, mean, invstd, gamma, sums, sums + c, inv_count,
                                                     static_cast<__half*>(dz), count8, c / 8, dbeta, dgamma, inv_scale);
  } else {
    const __half* zh = static_cast<const __half*>(z);
    bn_bwd_reduce_kernel<__half><<<reduce_grid(rows, c), 256, 0, st>>>(gh, zh, mean, invstd, rows, c, sums, sums + c);
    DIN_CHECK_CUDA(cudaGetLastError());
    bn_bwd_apply_kernel<__half><<<grid, 256, 0, st>>>(gh, zh, mean, invstd, gamma, sums, sums + c, inv_count,
                                                      static_cast<__half*>(dz), count8, c / 8, dbeta, dgamma, inv_scale);
  }
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}
