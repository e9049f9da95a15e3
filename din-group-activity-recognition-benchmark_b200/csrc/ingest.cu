// ingest.cu — the loader's per-frame work on the device: JPEG decode + resize to the
// model's input size, producing the uint8 NHWC frames the stem kernels ingest (din_stem_conv_nhwc_u8 /
// din_conv3x3_stem_pair_nhwc_f16 with x_is_u8).
//
// Replaces, per frame, volleyball.py:237-244 / collective.py:181-186:
//     img = Image.open(path); img = transforms.functional.resize(img, image_size); img = np.array(img)
// i.e. libjpeg-turbo on one DataLoader worker + Pillow's ImagingResample (BILINEAR, antialiased when shrinking).
//
// Resize: Pillow's algorithm restated exactly (src/libImaging/Resample.c, 8 bits per channel): per axis a table of
// window bounds and 22-bit fixed-point coefficients (built on the host in double precision, as Pillow builds them),
// horizontal pass into a uint8 intermediate, vertical pass; every output = clip8((2^21 + sum pixel * coeff) >> 22).
// Pure integer work: the result is bit-identical to PIL (tests/test_ingest_gpu.py).  HBM-bound byte streaming: each
// thread produces one output pixel (3 bytes) from <= ksize neighbouring pixels of a row / column that its warp reads
// as one contiguous span.
//
// Decode: nvJPEG (library code, like cuBLAS for a plain GEMM) through its batched API, resolved with dlopen at the first
// call so that libdin_sm100.so itself carries no link-time dependency on libnvjpeg; frames whose size already equals the
// target are decoded straight into the output tensor (Volleyball: 1280 x 720 sources, no resize pass at all).
#include <dlfcn.h>
#include <nvjpeg.h>

#include <cmath>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "din_common.cuh"

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;       // Pillow: PRECISION_BITS

struct AxisTable {
  int ksize = 0;
  std::vector<int> bounds;     // [out][2]: first source index, count
  std::vector<int> coeff;      // [out][ksize] fixed point
};

// Pillow's precompute_coeffs (bilinear filter, support 1) + normalize_coeffs_8bpc for the full-image box.
AxisTable make_axis_table(int in_size, int out_size) {
  AxisTable t;
  const double scale = static_cast<double>(static_cast<float>(in_size) - 0.0f) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  t.ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  t.bounds.assign(static_cast<size_t>(out_size) * 2, 0);
  t.coeff.assign(static_cast<size_t>(out_size) * t.ksize, 0);
  std::vector<double> k(t.ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      const double w = a < 1.0 ? 1.0 - a : 0.0;
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      const double v = k[x] * (1 << kPrecisionBits);
      t.coeff[static_cast<size_t>(xx) * t.ksize + x] = v < 0 ? static_cast<int>(-0.5 + v) : static_cast<int>(0.5 + v);
    }
    t.bounds[2 * xx] = xmin;
    t.bounds[2 * xx + 1] = xmax;
  }
  return t;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// One pass of the separable resize over interleaved RGB bytes.  HORIZ: dst[y][xx] from src[y][xmin .. xmin+cnt);
// else: dst[yy][x] from src[ymin .. ymin+cnt)[x].  `bounds` / `coeff` index the output coordinate along the resized axis.
template <bool HORIZ>
__global__ void __launch_bounds__(256)
resample_pass_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n, int src_h, int src_w, int dst_h,
                     int dst_w, const int* __restrict__ bounds, const int* __restrict__ coeff, int ksize) {
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  const long long total = static_cast<long long>(n) * dst_h * dst_w;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % dst_w);
  const int y = static_cast<int>((idx / dst_w) % dst_h);
  const int img = static_cast<int>(idx / (static_cast<long long>(dst_w) * dst_h));
  const int o = HORIZ ? x : y;
  const int first = __ldg(bounds + 2 * o), cnt = __ldg(bounds + 2 * o + 1);
  const int* k = coeff + static_cast<size_t>(o) * ksize;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  const uint8_t* base = src + static_cast<size_t>(img) * src_h * src_w * 3;
  for (int i = 0; i < cnt; ++i) {
    const uint8_t* p = HORIZ ? base + (static_cast<size_t>(y) * src_w + first + i) * 3
                             : base + (static_cast<size_t>(first + i) * src_w + x) * 3;
    const int c = __ldg(k + i);
    s0 += static_cast<int>(p[0]) * c;
    s1 += static_cast<int>(p[1]) * c;
    s2 += static_cast<int>(p[2]) * c;
  }
  uint8_t* q = dst + idx * 3;
  q[0] = clip8(s0);
  q[1] = clip8(s1);
  q[2] = clip8(s2);
}

// device copies of the coefficient tables, cached per (in, out) size pair and device (a loader resizes every frame of a
// dataset between the same two sizes)
struct DeviceTable {
  int dev, in_size, out_size, ksize;
  int* bounds;
  int* coeff;
};
std::mutex g_tab_mutex;
std::vector<DeviceTable> g_tables;

int get_table(int in_size, int out_size, cudaStream_t st, DeviceTable* out) {
  int dev = 0;
  DIN_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_tab_mutex);
  for (const DeviceTable& t : g_tables)
    if (t.dev == dev && t.in_size == in_size && t.out_size == out_size) { *out = t; return DIN_OK; }
  const AxisTable h = make_axis_table(in_size, out_size);
  DeviceTable t{dev, in_size, out_size, h.ksize, nullptr, nullptr};
  DIN_CHECK_CUDA(cudaMalloc(&t.bounds, h.bounds.size() * sizeof(int)));
  DIN_CHECK_CUDA(cudaMalloc(&t.coeff, h.coeff.size() * sizeof(int)));
  // synchronous copies from pageable memory: once per size pair
  DIN_CHECK_CUDA(cudaMemcpy(t.bounds, h.bounds.data(), h.bounds.size() * sizeof(int), cudaMemcpyHostToDevice));
  DIN_CHECK_CUDA(cudaMemcpy(t.coeff, h.coeff.data(), h.coeff.size() * sizeof(int), cudaMemcpyHostToDevice));
  (void)st;
  g_tables.push_back(t);
  *out = t;
  return DIN_OK;
}

int resize_launch(const uint8_t* src, int n, int h, int w, uint8_t* dst, int oh, int ow, uint8_t* tmp, cudaStream_t st) {
  if (h == oh && w == ow) {                              // Pillow returns a copy
    if (src != dst) DIN_CHECK_CUDA(cudaMemcpyAsync(dst, src, static_cast<size_t>(n) * h * w * 3, cudaMemcpyDeviceToDevice, st));
    return DIN_OK;
  }
  const bool need_h = w != ow, need_v = h != oh;
  const uint8_t* cur = src;
  if (need_h) {
    DeviceTable t;
    int rc = get_table(w, ow, st, &t);
    if (rc != DIN_OK) return rc;
    uint8_t* out = need_v ? tmp : dst;
    const long long total = static_cast<long long>(n) * h * ow;
    resample_pass_kernel<true><<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(cur, out, n, h, w, h, ow, t.bounds,
                                                                                           t.coeff, t.ksize);
    DIN_CHECK_CUDA(cudaGetLastError());
    cur = out;
  }
  if (need_v) {
    DeviceTable t;
    int rc = get_table(h, oh, st, &t);
    if (rc != DIN_OK) return rc;
    const long long total = static_cast<long long>(n) * oh * ow;
    resample_pass_kernel<false><<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(cur, dst, n, h, ow, oh, ow, t.bounds,
                                                                                            t.coeff, t.ksize);
    DIN_CHECK_CUDA(cudaGetLastError());
  }
  return DIN_OK;
}

// ---- nvJPEG through dlopen ---------------------------------------------------------------------------------------------
struct NvJpegApi {
  void* lib = nullptr;
  decltype(&nvjpegCreateEx) CreateEx = nullptr;
  decltype(&nvjpegJpegStateCreate) JpegStateCreate = nullptr;
  decltype(&nvjpegGetImageInfo) GetImageInfo = nullptr;
  decltype(&nvjpegDecodeBatchedInitialize) DecodeBatchedInitialize = nullptr;
  decltype(&nvjpegDecodeBatched) DecodeBatched = nullptr;
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  int dev = -1, batch = -1, threads = -1, backend = -1;
};
std::mutex g_jpeg_mutex;
NvJpegApi g_jpeg;

int jpeg_api(NvJpegApi** out) {
  NvJpegApi& a = g_jpeg;
  if (a.lib == nullptr) {
    const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"};
    for (const char* nm : names) {
      a.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (a.lib) break;
    }
    if (a.lib == nullptr)
      return din_set_error(DIN_ERR_UNSUPPORTED, "din_jpeg_decode_resize_u8: libnvjpeg.so.12 not found (%s)", dlerror());
#define DIN_JPEG_SYM(field, sym)                                                                      \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, sym));                                 \
    if (a.field == nullptr) return din_set_error(DIN_ERR_UNSUPPORTED, "din_jpeg_decode_resize_u8: %s missing in libnvjpeg", sym)
    DIN_JPEG_SYM(CreateEx, "nvjpegCreateEx");
    DIN_JPEG_SYM(JpegStateCreate, "nvjpegJpegStateCreate");
    DIN_JPEG_SYM(GetImageInfo, "nvjpegGetImageInfo");
    DIN_JPEG_SYM(DecodeBatchedInitialize, "nvjpegDecodeBatchedInitialize");
    DIN_JPEG_SYM(DecodeBatched, "nvjpegDecodeBatched");
#undef DIN_JPEG_SYM
  }
  int dev = 0;
  DIN_CHECK_CUDA(cudaGetDevice(&dev));
  if (a.handle == nullptr || a.dev != dev) {
    // (one handle per process; a second device re-creates it -- loaders decode on one device)
    // backend: DIN_NVJPEG_BACKEND = 3 (hardware engine), 2 (GPU-assisted Huffman), 1 (CPU Huffman), 0 (nvJPEG's default);
    // unset: try the hardware engine, then the GPU-assisted backend, then the default
    const char* e = std::getenv("DIN_NVJPEG_BACKEND");
    const int order_auto[3] = {NVJPEG_BACKEND_HARDWARE, NVJPEG_BACKEND_GPU_HYBRID, NVJPEG_BACKEND_DEFAULT};
    const int forced = e ? std::atoi(e) : -1;
    nvjpegStatus_t s = NVJPEG_STATUS_NOT_INITIALIZED;
    for (int i = 0; i < 3; ++i) {
      const int be = forced >= 0 ? forced : order_auto[i];
      s = a.CreateEx(static_cast<nvjpegBackend_t>(be), nullptr, nullptr, 0, &a.handle);
      if (s == NVJPEG_STATUS_SUCCESS) { a.backend = be; break; }
      if (forced >= 0) break;
    }
    if (s != NVJPEG_STATUS_SUCCESS) return din_set_error(DIN_ERR_CUDA, "nvjpegCreateEx failed (status %d)", static_cast<int>(s));
    s = a.JpegStateCreate(a.handle, &a.state);
    if (s != NVJPEG_STATUS_SUCCESS) return din_set_error(DIN_ERR_CUDA, "nvjpegJpegStateCreate failed (status %d)", static_cast<int>(s));
    a.dev = dev;
    a.batch = -1;
  }
  *out = &a;
  return DIN_OK;
}

}  // namespace

extern "C" int din_resize_bilinear_u8(const uint8_t* src, int n, int h, int w, uint8_t* dst, int oh, int ow, uint8_t* tmp,
                                      void* stream) {
  DIN_CHECK_ARG(src && dst, "din_resize_bilinear_u8: null pointer");
  DIN_CHECK_ARG(n > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "din_resize_bilinear_u8: bad shape n=%d %dx%d -> %dx%d", n, h, w, oh,
                ow);
  DIN_CHECK_ARG(!(h != oh && w != ow) || tmp != nullptr,
                "din_resize_bilinear_u8: both axes change: tmp (n * h * ow * 3 bytes) is required");
  return resize_launch(src, n, h, w, dst, oh, ow, tmp, static_cast<cudaStream_t>(stream));
}

extern "C" int din_jpeg_backend(void) {
  std::lock_guard<std::mutex> lock(g_jpeg_mutex);
  return g_jpeg.backend;
}

extern "C" int din_jpeg_image_info(const unsigned char* jpeg, size_t nbytes, int* h, int* w) {
  DIN_CHECK_ARG(jpeg && nbytes > 0 && h && w, "din_jpeg_image_info: null pointer");
  std::lock_guard<std::mutex> lock(g_jpeg_mutex);
  NvJpegApi* a = nullptr;
  int rc = jpeg_api(&a);
  if (rc != DIN_OK) return rc;
  int ncomp = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
  nvjpegChromaSubsampling_t sub;
  const nvjpegStatus_t s = a->GetImageInfo(a->handle, jpeg, nbytes, &ncomp, &sub, ws, hs);
  if (s != NVJPEG_STATUS_SUCCESS)
    return din_set_error(DIN_ERR_INVALID_ARG, "din_jpeg_image_info: not a decodable JPEG stream (nvjpeg status %d)", static_cast<int>(s));
  *h = hs[0];
  *w = ws[0];
  return DIN_OK;
}

extern "C" int din_jpeg_decode_resize_u8(const unsigned char* const* jpeg, const size_t* nbytes, int n, uint8_t* frames,
                                         int out_h, int out_w, uint8_t* workspace, size_t workspace_bytes, int cpu_threads,
                                         void* stream) {
  const char* who = "din_jpeg_decode_resize_u8";
  DIN_CHECK_ARG(jpeg && nbytes && frames, "%s: null pointer", who);
  DIN_CHECK_ARG(n > 0 && out_h > 0 && out_w > 0, "%s: bad shape n=%d -> %dx%d", who, n, out_h, out_w);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::lock_guard<std::mutex> lock(g_jpeg_mutex);
  NvJpegApi* a = nullptr;
  int rc = jpeg_api(&a);
  if (rc != DIN_OK) return rc;
  if (cpu_threads < 1) cpu_threads = 1;
  // sizes, and where each frame is decoded: straight into `frames` when no resize is needed, else into the workspace
  std::vector<int> hs(n), ws(n);
  std::vector<nvjpegImage_t> dst(n);
  std::vector<size_t> ws_off(n, 0);
  size_t need = 0;
  const size_t frame_bytes = static_cast<size_t>(out_h) * out_w * 3;
  for (int i = 0; i < n; ++i) {
    int ncomp = 0, w4[NVJPEG_MAX_COMPONENT], h4[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t sub;
    const nvjpegStatus_t s = a->GetImageInfo(a->handle, jpeg[i], nbytes[i], &ncomp, &sub, w4, h4);
    if (s != NVJPEG_STATUS_SUCCESS)
      return din_set_error(DIN_ERR_INVALID_ARG, "%s: frame %d is not a decodable JPEG stream (nvjpeg status %d)", who, i,
                           static_cast<int>(s));
    hs[i] = h4[0];
    ws[i] = w4[0];
    if (hs[i] != out_h || ws[i] != out_w) {
      ws_off[i] = need;
      // decoded frame + the horizontal pass's intermediate, 256-byte aligned
      need += ((static_cast<size_t>(hs[i]) * ws[i] * 3 + 255) & ~static_cast<size_t>(255)) +
              ((static_cast<size_t>(hs[i]) * out_w * 3 + 255) & ~static_cast<size_t>(255));
    }
  }
  DIN_CHECK_ARG(need == 0 || (workspace != nullptr && workspace_bytes >= need),
                "%s: frames need resizing: workspace of %zu bytes required (got %zu)", who, need, workspace_bytes);
  for (int i = 0; i < n; ++i) {
    for (int c = 0; c < NVJPEG_MAX_COMPONENT; ++c) { dst[i].channel[c] = nullptr; dst[i].pitch[c] = 0; }
    const bool direct = hs[i] == out_h && ws[i] == out_w;
    dst[i].channel[0] = direct ? frames + static_cast<size_t>(i) * frame_bytes : workspace + ws_off[i];
    dst[i].pitch[0] = static_cast<size_t>(ws[i]) * 3;
  }
  if (a->batch != n || a->threads != cpu_threads) {
    const nvjpegStatus_t s = a->DecodeBatchedInitialize(a->handle, a->state, n, cpu_threads, NVJPEG_OUTPUT_RGBI);
    if (s != NVJPEG_STATUS_SUCCESS)
      return din_set_error(DIN_ERR_CUDA, "%s: nvjpegDecodeBatchedInitialize failed (status %d)", who, static_cast<int>(s));
    a->batch = n;
    a->threads = cpu_threads;
  }
  const nvjpegStatus_t s = a->DecodeBatched(a->handle, a->state, jpeg, nbytes, dst.data(), st);
  if (s != NVJPEG_STATUS_SUCCESS)
    return din_set_error(DIN_ERR_CUDA, "%s: nvjpegDecodeBatched failed (status %d)", who, static_cast<int>(s));
  for (int i = 0; i < n; ++i) {
    if (hs[i] == out_h && ws[i] == out_w) continue;
    uint8_t* raw = workspace + ws_off[i];
    uint8_t* tmp = raw + ((static_cast<size_t>(hs[i]) * ws[i] * 3 + 255) & ~static_cast<size_t>(255));
    rc = resize_launch(raw, 1, hs[i], ws[i], frames + static_cast<size_t>(i) * frame_bytes, out_h, out_w, tmp, st);
    if (rc != DIN_OK) return rc;
  }
  return DIN_OK;
}
