// loss.cu — the step right after the hot path, kept on the device, and the small
// reduction the stage-1 base model needs.
//
// Reference call sites:
//   F.cross_entropy (+ class weights), argmax, correct count, ConfusionMeter.add, AverageMeter.update
//       train_net_dynamic.py:191-199,201-210,217  and  :258-292 (test loop); utils.py:193-264
//   actions_scores.reshape(B,T,N,-1).mean(dim=1)     base_model.py:138-139
//
// The reference pulls three scalars to the host per step (`.item()` x2 and ConfusionMeter's .cpu()); here the
// loss, the correct count, the confusion matrix and the epoch meters are produced by ONE launch and stay in
// device memory until the caller reads them (once per epoch).  The same launch writes d(loss)/d(logits), the
// seed of the backward pass.
#include <cfloat>
#include <cstdio>

#include "din_common.cuh"

namespace {

using namespace din;

constexpr int kCeThreads = 256;
constexpr int kCeMaxClasses = 64;

// One CTA.  Row i: m = max_j z_ij;  lse = m + log(sum_j exp(z_ij - m));  nll_i = lse - z_{i,y_i}.
// loss = sum_i w[y_i] nll_i / sum_i w[y_i]   (w == 1 without class weights: the plain mean) -- torch's
// 'mean' reduction (F.cross_entropy, weight=...).  argmax takes the FIRST maximum, as torch.argmax.
// Block reductions run in a fixed order: the result is deterministic.
__global__ void __launch_bounds__(kCeThreads)
ce_metrics_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                  const float* __restrict__ class_weight, float loss_scale, float* __restrict__ loss,
                  int* __restrict__ correct, int* __restrict__ conf, double* __restrict__ meters,
                  float* __restrict__ dlogits, int B, int A) {
  __shared__ float red_f[kCeThreads];
  __shared__ float red_w[kCeThreads];
  __shared__ int red_i[kCeThreads];
  __shared__ float wsum_s;
  float nll_w = 0.0f, w_sum = 0.0f;
  int n_correct = 0;
  for (int i = threadIdx.x; i < B; i += kCeThreads) {
    const float* z = logits + static_cast<size_t>(i) * A;
    const int y = static_cast<int>(labels[i]);
    float m = -FLT_MAX;
    int arg = 0;
    for (int j = 0; j < A; ++j) {
      const float v = z[j];
      if (v > m) { m = v; arg = j; }
    }
    float se = 0.0f;
    for (int j = 0; j < A; ++j) se += expf(z[j] - m);
    const float lse = logf(se);                              // log_softmax_j = (z_j - m) - lse, as torch
    // torch: ignore_index = -100 contributes nothing, any other label outside [0, A) is a device-side assert
    const bool y_ok = y >= 0 && y < A;
    if (!y_ok && y != -100) {
      printf("din_ce_metrics_f32: label %d of sample %d is outside [0, %d)\n", y, i, A);
      __trap();
    }
    const float w = y_ok ? (class_weight ? class_weight[y] : 1.0f) : 0.0f;
    if (y_ok) {
      nll_w += w * -((z[y] - m) - lse);
      w_sum += w;
      n_correct += (arg == y) ? 1 : 0;
      if (conf != nullptr) atomicAdd(conf + y * A + arg, 1);  // conf[target][predicted] (utils.py:257-264)
    }
  }
  red_f[threadIdx.x] = nll_w;
  red_w[threadIdx.x] = w_sum;
  red_i[threadIdx.x] = n_correct;
  __syncthreads();
  for (int s = kCeThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      red_f[threadIdx.x] += red_f[threadIdx.x + s];
      red_w[threadIdx.x] += red_w[threadIdx.x + s];
      red_i[threadIdx.x] += red_i[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float l = loss_scale * (red_f[0] / red_w[0]);
    wsum_s = red_w[0];
    if (loss != nullptr) *loss = l;
    if (correct != nullptr) *correct = red_i[0];
    if (meters != nullptr) {                                 // AverageMeter sums (utils.py AverageMeter.update)
      meters[0] += static_cast<double>(l) * B;               //   loss_meter.update(loss, batch_size)
      meters[1] += static_cast<double>(B);
      meters[2] += static_cast<double>(red_i[0]);            //   accuracy_meter: sum of correct
      meters[3] += 1.0;                                      //   steps
    }
  }
  if (dlogits == nullptr) return;
  __syncthreads();
  const float inv_w = loss_scale / wsum_s;
  for (int i = threadIdx.x; i < B; i += kCeThreads) {
    const float* z = logits + static_cast<size_t>(i) * A;
    float* dz = dlogits + static_cast<size_t>(i) * A;
    const int y = static_cast<int>(labels[i]);
    const bool y_ok = y >= 0 && y < A;
    const float w = y_ok ? (class_weight ? class_weight[y] : 1.0f) : 0.0f;
    float m = -FLT_MAX;
    for (int j = 0; j < A; ++j) m = fmaxf(m, z[j]);
    float se = 0.0f;
    for (int j = 0; j < A; ++j) se += expf(z[j] - m);
    const float g = w * inv_w;
    for (int j = 0; j < A; ++j) dz[j] = g * (expf(z[j] - m) / se - (j == y ? 1.0f : 0.0f));
  }
}

// y[o, i] = mean_a x[o, a, i]
__global__ void __launch_bounds__(256)
mean_axis_kernel(const float* __restrict__ x, float* __restrict__ y, int len, int inner, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= total) return;
  const long long o = idx / inner;
  const int i = static_cast<int>(idx - o * inner);
  const float* p = x + (o * len) * inner + i;
  float s = 0.0f;
  for (int a = 0; a < len; ++a) s += p[static_cast<size_t>(a) * inner];
  y[idx] = s / static_cast<float>(len);
}

}  // namespace

extern "C" int din_ce_metrics_f32(const float* logits, const int64_t* labels, const float* class_weight,
                                  float loss_scale, float* loss, int32_t* correct, int32_t* conf, double* meters,
                                  float* dlogits, int b, int a, void* stream) {
  DIN_CHECK_ARG(logits && labels, "din_ce_metrics_f32: null pointer");
  DIN_CHECK_ARG(b > 0 && a > 1 && a <= kCeMaxClasses, "din_ce_metrics_f32: bad shape b=%d a=%d (2 <= a <= %d)", b, a,
                kCeMaxClasses);
  ce_metrics_kernel<<<1, kCeThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, reinterpret_cast<const long long*>(labels), class_weight, loss_scale, loss, correct, conf, meters,
      dlogits, b, a);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_mean_axis_f32(const float* x, float* y, int outer, int len, int inner, void* stream) {
  DIN_CHECK_ARG(x && y, "din_mean_axis_f32: null pointer");
  DIN_CHECK_ARG(outer > 0 && len > 0 && inner > 0, "din_mean_axis_f32: bad shape outer=%d len=%d inner=%d", outer,
                len, inner);
  const long long total = static_cast<long long>(outer) * inner;
  const long long grid = (total + 255) / 256;
  DIN_CHECK_ARG(grid <= INT32_MAX, "din_mean_axis_f32: too large");
  mean_axis_kernel<<<static_cast<int>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, len, inner, total);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}
