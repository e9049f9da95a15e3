// flat.cu — gradient tensors <-> one flat fp32 buffer, many tensors per launch.
//
// The data-parallel training step ends with all-reduces over a flat gradient buffer (replaces the
// gather-on-GPU-0 of nn.DataParallel, train_net_dynamic.py:96).  The first version filled that buffer with one
// eager copy per parameter and emptied it with another (~90 + ~90 launches for VGG-16, all latency).  Here a whole
// bucket of tensors is scaled and packed by ONE launch: the job table travels in the kernel parameters, every block
// finds its job in a prefix table, float4 where source and destination are 16-byte aligned.
#include "din_common.cuh"

namespace {

constexpr int kFlatMaxJobs = 96;        // 96 * (8 + 8 + 8) + 97 * 4 bytes < the 4 KB kernel-parameter space
constexpr int kFlatThreads = 256;
constexpr long long kFlatElemsPerBlock = 8192;

struct FlatJobs {
  const float* src[kFlatMaxJobs];
  long long dst_off[kFlatMaxJobs];
  long long numel[kFlatMaxJobs];
  int block0[kFlatMaxJobs + 1];         // first block of each job; block0[n] = grid size
  int n;
};

__global__ void __launch_bounds__(kFlatThreads)
pack_flat_kernel(const __grid_constant__ FlatJobs jobs, float* __restrict__ flat, float scale) {
  int j = 0;
  while (j + 1 < jobs.n && static_cast<int>(blockIdx.x) >= jobs.block0[j + 1]) ++j;
  const long long first = static_cast<long long>(blockIdx.x - jobs.block0[j]) * kFlatElemsPerBlock;
  const long long n = jobs.numel[j];
  const long long count = n - first < kFlatElemsPerBlock ? n - first : kFlatElemsPerBlock;
  const float* s = jobs.src[j] + first;
  float* d = flat + jobs.dst_off[j] + first;
  if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
    const long long n4 = count >> 2;
    for (long long i = threadIdx.x; i < n4; i += kFlatThreads) {
      float4 v = __ldg(reinterpret_cast<const float4*>(s) + i);
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      reinterpret_cast<float4*>(d)[i] = v;
    }
    for (long long i = (n4 << 2) + threadIdx.x; i < count; i += kFlatThreads) d[i] = __ldg(s + i) * scale;
  } else {
    for (long long i = threadIdx.x; i < count; i += kFlatThreads) d[i] = __ldg(s + i) * scale;
  }
}

}  // namespace

extern "C" int din_pack_flat_f32(const DinFlatJob* jobs, int n_jobs, float* flat, float scale, void* stream) {
  DIN_CHECK_ARG(jobs && flat && n_jobs > 0, "din_pack_flat_f32: null pointer / no jobs");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int j0 = 0; j0 < n_jobs; j0 += kFlatMaxJobs) {
    FlatJobs t{};
    t.n = n_jobs - j0 < kFlatMaxJobs ? n_jobs - j0 : kFlatMaxJobs;
    long long blocks = 0;
    for (int i = 0; i < t.n; ++i) {
      const DinFlatJob& jb = jobs[j0 + i];
      DIN_CHECK_ARG(jb.src && jb.numel > 0 && jb.dst_offset >= 0, "din_pack_flat_f32: job %d: bad src / numel / offset",
                    j0 + i);
      DIN_CHECK_ARG((reinterpret_cast<uintptr_t>(jb.src) & 3) == 0, "din_pack_flat_f32: job %d: unaligned source", j0 + i);
      t.src[i] = static_cast<const float*>(jb.src);
      t.dst_off[i] = jb.dst_offset;
      t.numel[i] = jb.numel;
      t.block0[i] = static_cast<int>(blocks);
      blocks += (jb.numel + kFlatElemsPerBlock - 1) / kFlatElemsPerBlock;
      DIN_CHECK_ARG(blocks < (1ll << 30), "din_pack_flat_f32: too many elements in one call");
    }
    t.block0[t.n] = static_cast<int>(blocks);
    pack_flat_kernel<<<static_cast<int>(blocks), kFlatThreads, 0, st>>>(t, flat, scale);
    DIN_CHECK_CUDA(cudaGetLastError());
  }
  return DIN_OK;
}
