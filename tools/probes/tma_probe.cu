// Which TMA tile-load shapes does the B200 accept?  (probe for the fused conv1 kernel's image-patch loads)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tma_probe tools/probes/tma_probe.cu
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

static PFN_cuTensorMapEncodeTiled_v12000 g_encode;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK, bool CLUSTER>
__device__ void body(const CUtensorMap* m, float* out, int n_floats, int c0, int c1, int c2, int c3) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  float* dst = reinterpret_cast<float*>(smem + 1024);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n_floats * 4) : "memory");
    if constexpr (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
  }
  uint32_t ok = 0, spins = 0;
  while (!ok && ++spins < (1u << 22)) {
    asm volatile("{ .reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0; selp.u32 %0, 1, 0, P; }" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
  }
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < n_floats; i += blockDim.x) out[i] = ok ? dst[i] : -12345.0f;
}

template <int RANK> __global__ void k_plain(const __grid_constant__ CUtensorMap m, float* out, int n, int c0, int c1, int c2, int c3) {
  body<RANK, false>(&m, out, n, c0, c1, c2, c3);
}
template <int RANK> __global__ void __cluster_dims__(2, 1, 1) k_cluster(const __grid_constant__ CUtensorMap m, float* out, int n, int c0, int c1, int c2, int c3) {
  body<RANK, true>(&m, out, n, c0, c1, c2, c3);
}

static CUtensorMapSwizzle g_swz = CU_TENSOR_MAP_SWIZZLE_NONE;
static bool encode(CUtensorMap* tm, CUtensorMapDataType dt, int rank, void* base, const uint64_t* dims, const uint64_t* strides,
                   const uint32_t* box, CUtensorMapL2promotion l2) {
  uint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(tm, dt, rank, base, dims, strides + 1, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, g_swz,
                        l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("  encode failed: %d\n", (int)r);
  return r == CUDA_SUCCESS;
}

template <typename K> static void run(const char* what, K kern, const CUtensorMap& tm, int n_floats, int c0, int c1, int c2, int c3,
                                      const std::vector<float>& want) {
  float* out;
  cudaMalloc(&out, n_floats * 4);
  cudaMemset(out, 0, n_floats * 4);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  kern<<<2, 128, 32 * 1024>>>(tm, out, n_floats, c0, c1, c2, c3);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-60s -> %s\n", what, cudaGetErrorString(e)); exit(0); }
  std::vector<float> got(n_floats);
  cudaMemcpy(got.data(), out, n_floats * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < n_floats; ++i) bad += got[i] != want[i];
  printf("%-60s -> ok, %d of %d values differ from the expected box\n", what, bad, n_floats);
  cudaFree(out);
}

int main(int argc, char** argv) {
  int which = argc > 1 ? atoi(argv[1]) : 0;
  int x0 = argc > 2 ? atoi(argv[2]) : -2, y0 = argc > 3 ? atoi(argv[3]) : -2, c0 = argc > 4 ? atoi(argv[4]) : 3;
  if (argc > 5 && atoi(argv[5]) == 128) g_swz = CU_TENSOR_MAP_SWIZZLE_128B;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  const int W = 64, H = 48, C = 6;
  std::vector<float> img(W * H * C);
  for (size_t i = 0; i < img.size(); ++i) img[i] = (float)i;
  float* d;
  cudaMalloc(&d, img.size() * 4);
  cudaMemcpy(d, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
  auto expect = [&](int bw, int bh, int bc, int x0, int y0, int c0) {
    std::vector<float> v(bw * bh * bc);
    for (int c = 0; c < bc; ++c) for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) {
      int gx = x0 + x, gy = y0 + y, gc = c0 + c;
      v[(c * bh + y) * bw + x] = (gx < 0 || gx >= W || gy < 0 || gy >= H || gc < 0 || gc >= C) ? 0.0f : img[(gc * H + gy) * W + gx];
    }
    return v;
  };
  CUtensorMap tm;
  uint64_t dims3[3] = {W, H, C}, str3[3] = {4, W * 4, (uint64_t)W * H * 4};
  uint64_t dims4[4] = {W, H, C, 1}, str4[4] = {4, W * 4, (uint64_t)W * H * 4, (uint64_t)W * H * C * 4};
  struct Case { const char* name; int rank; uint32_t bw, bh, bc; CUtensorMapL2promotion l2; bool cluster; };
  Case cases[] = {
      {"3d box {16,8,2} L2_256B plain", 3, 16, 8, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, false},
      {"3d box {12,20,3} L2_256B plain", 3, 12, 20, 3, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, false},
      {"3d box {12,20,3} L2_NONE plain", 3, 12, 20, 3, CU_TENSOR_MAP_L2_PROMOTION_NONE, false},
      {"3d box {12,20,3} L2_128B plain", 3, 12, 20, 3, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, false},
      {"3d box {16,20,3} L2_256B plain", 3, 16, 20, 3, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, false},
      {"3d box {12,20,3} L2_NONE cluster", 3, 12, 20, 3, CU_TENSOR_MAP_L2_PROMOTION_NONE, true},
      {"3d box {16,20,3} L2_256B cluster", 3, 16, 20, 3, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, true},
      {"4d box {12,20,3,1} L2_256B plain", 4, 12, 20, 3, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, false},
      {"4d box {16,20,3,1} L2_256B cluster", 4, 16, 20, 3, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, true},
      {"3d box {32,8,2} L2_256B plain (128-byte rows)", 3, 32, 8, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, false},
  };
  const Case& cs = cases[which];
  uint32_t box[4] = {cs.bw, cs.bh, cs.bc, 1};
  if (!encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, cs.rank, d, cs.rank == 3 ? dims3 : dims4, cs.rank == 3 ? str3 : str4, box, cs.l2)) return 0;
  auto want = expect(cs.bw, cs.bh, cs.bc, x0, y0, c0);
  int n = cs.bw * cs.bh * cs.bc;
  printf("[x0=%d y0=%d c0=%d swizzle=%d] ", x0, y0, c0, (int)g_swz);
  if (cs.rank == 3) { if (cs.cluster) run(cs.name, k_cluster<3>, tm, n, x0, y0, c0, 0, want); else run(cs.name, k_plain<3>, tm, n, x0, y0, c0, 0, want); }
  else { if (cs.cluster) run(cs.name, k_cluster<4>, tm, n, x0, y0, c0, 0, want); else run(cs.name, k_plain<4>, tm, n, x0, y0, c0, 0, want); }
  return 0;
}
