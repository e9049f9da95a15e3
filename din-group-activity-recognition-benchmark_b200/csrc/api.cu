// api.cu — library-level entry points of libdin_sm100.so: versioning, thread-local error string,
// device queries and the tensor-map encoder shared by the TMA kernels.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include <cudaTypedefs.h>

#include "din_common.cuh"

namespace {
thread_local char g_err[512] = {0};

PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
std::once_flag g_encode_once;

void resolve_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) {
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
}
}  // namespace

int din_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int din_encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                    CUtensorMapSwizzle swizzle) {
  std::call_once(g_encode_once, resolve_encode);
  if (!g_encode) return din_set_error(DIN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  // strides_bytes[0] is the (implied) element stride; the driver takes rank-1 outer strides.
  CUresult r = g_encode(out, dtype, static_cast<cuuint32_t>(rank), base,
                        reinterpret_cast<const cuuint64_t*>(dims),
                        reinterpret_cast<const cuuint64_t*>(strides_bytes + 1),
                        reinterpret_cast<const cuuint32_t*>(box),
                        reinterpret_cast<const cuuint32_t*>(elem_strides), CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return din_set_error(DIN_ERR_CUDA,
                         "cuTensorMapEncodeTiled failed (CUresult %d): rank %d base %p dims [%llu %llu %llu %llu] "
                         "box [%u %u %u %u]",
                         static_cast<int>(r), rank, base, (unsigned long long)dims[0],
                         (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
                         (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], rank > 1 ? box[1] : 0,
                         rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  }
  return DIN_OK;
}

int din_num_sms() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev != cached_dev) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached_dev = dev;
    cached_sms = sms;
  }
  return cached_sms;
}

extern "C" {

// 2: + uint8 ingest, loss / metrics, backward entry points
// 3: + batched weight packing, dgrad with fused ReLU backward, ResNet-18 backward helpers, batch-statistics BatchNorm
int din_abi_version(void) { return 3; }

const char* din_last_error_string(void) { return g_err; }

int din_device_sm_count(void) {
  int n = din_num_sms();
  if (n <= 0) return din_set_error(DIN_ERR_CUDA, "no CUDA device available");
  return n;
}

}  // extern "C"
