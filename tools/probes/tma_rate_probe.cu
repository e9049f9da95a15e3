// How fast does one SM's TMA unit deliver halo boxes?  Persistent CTAs (one per SM) stream the {64 ch, 10, 18} SWIZZLE_128B
// boxes of a 3x3 convolution's 16x8 tiles (and, for comparison, dense 2-D boxes of the same bytes) through a 4-stage ring with
// nothing consuming them: cycles per box and per 128-byte box row.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tma_rate_probe tools/probes/tma_rate_probe.cu
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

static PFN_cuTensorMapEncodeTiled_v12000 g_encode;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kStages = 4;
constexpr int kStageBytes = 24 * 1024;

// mode 0: 4-D halo boxes (coordinates of tile t of an [n, h, w, c] map), mode 1: 2-D boxes {64, rows} of the flat [pixels, c] view
__global__ void __launch_bounds__(128, 1)
stream_kernel(const __grid_constant__ CUtensorMap tm, int mode, int tiles, int tiles_x, int tiles_per_img, int box_bytes,
              int rows2d, long long* cycles, int channels_blocks) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  __shared__ uint64_t full[kStages];
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const long long t0 = clock64();
  int issued = 0, done = 0;
  auto issue = [&](int k) {
    const int s = k % kStages;
    const int t = blockIdx.x + k * gridDim.x;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(box_bytes * channels_blocks) : "memory");
    for (int cb = 0; cb < channels_blocks; ++cb) {
      uint8_t* dst = smem + s * kStageBytes * 1 + cb * 0;   // blocks of one tile overwrite each other (nothing reads them)
      if (mode == 0) {
        const int img = t / tiles_per_img, r = t - img * tiles_per_img, ty = r / tiles_x, tx = r - ty * tiles_x;
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(smem_u32(dst)), "l"((uint64_t)&tm), "r"(smem_u32(&full[s])), "r"(cb * 64), "r"(tx * 8 - 1), "r"(ty * 16 - 1), "r"(img) : "memory");
      } else {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(dst)), "l"((uint64_t)&tm), "r"(smem_u32(&full[s])), "r"(cb * 64), "r"(t * rows2d) : "memory");
      }
    }
  };
  const int mine = (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  for (; issued < mine && issued < kStages; ++issued) issue(issued);
  for (; done < mine; ++done) {
    const int s = done % kStages;
    const uint32_t parity = (done / kStages) & 1;
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2; selp.u32 %0, 1, 0, P; }" : "=r"(ok) : "r"(smem_u32(&full[s])), "r"(parity) : "memory");
    if (issued < mine) { issue(issued); ++issued; }
  }
  cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  const int n = 54, h = 180, w = 320;
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kStageBytes + 1024);
  for (int c : {64, 32, 128, 512}) {
    __half* x;
    const size_t elems = (size_t)n * h * w * c;
    cudaMalloc(&x, elems * 2);
    cudaMemset(x, 0, elems * 2);
    const int tiles_x = w / 8, tiles_per_img = tiles_x * ((h + 15) / 16), tiles = n * tiles_per_img;
    const int cblocks = (c + 63) / 64;
    for (int mode = 0; mode < 2; ++mode) {
      CUtensorMap tm;
      uint32_t es[4] = {1, 1, 1, 1};
      CUresult r;
      int box_bytes, rows;
      if (mode == 0) {
        uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)n};
        uint64_t st[3] = {(uint64_t)c * 2, (uint64_t)c * 2 * w, (uint64_t)c * 2 * w * h};
        uint32_t box[4] = {64, 10, 18, 1};
        r = g_encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        box_bytes = 180 * 128; rows = 180;
      } else {
        uint64_t dims[2] = {(uint64_t)c, (uint64_t)n * h * w};
        uint64_t st[1] = {(uint64_t)c * 2};
        uint32_t box[2] = {64, 128};                 // 128 pixels = the tile's own bytes, no halo
        r = g_encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, x, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        box_bytes = 128 * 128; rows = 128;
      }
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
      const int t2 = mode == 0 ? tiles : (int)((size_t)n * h * w / 128);
      float ms = 0;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        stream_kernel<<<148, 128, kStages * kStageBytes + 1024>>>(tm, mode, t2, tiles_x, tiles_per_img, box_bytes, 128, cyc, cblocks);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
      }
      long long hc[148];
      cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < 148; ++i) mx = hc[i] > mx ? hc[i] : mx;
      const double per_cta = (double)t2 / 148;
      const double real_bytes = (double)t2 * rows * (c < 64 ? c * 2 : 128) * cblocks;
      printf("c=%3d %s: %7.3f ms, %8.0f cycles per tile (%d box(es) of %d rows), %5.1f cycles per 128-byte box row, %6.0f GB/s fetched (err %s)\n",
             c, mode == 0 ? "halo 4-D {64,10,18}" : "dense 2-D {64,128} ", ms, mx / per_cta, cblocks, rows, mx / per_cta / (rows * cblocks),
             real_bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(x);
  }
  return 0;
}
