// stem_tc.cu — the 3-channel stem convolution on the tensor cores.
//
// Replaces prep_images (utils.py:8-19) + the first conv of each backbone (VGG-16 features.0 3x3 s1 p1,
// ResNet-18 conv1 7x7 s2 p3 + BN, Inception-v3 Conv2d_1a_3x3 s2 p0 + BN; backbone.py:88-99,115-132,44).
//
// The layer is HBM-bound (VGG: 11 MB fp32 in, 118 MB fp16 out per frame, 3.2 GFLOP) but as a CUDA-core
// direct conv it cost 14.5 ms of a 61 ms step.  Here K = 3*kh*kw (27 / 147) is padded to a multiple of 16
// and each CTA turns a strip of 128 output pixels into one M=128 UMMA tile:
//   * 128 threads = 128 pixels: each thread gathers its K input values from the NCHW fp32 image (coalesced
//     along x), applies prep_images with the reference's three roundings, converts to fp16 and writes its
//     row of the A tile straight into the canonical no-swizzle K-major UMMA layout (8x16-byte core
//     matrices; LBO = 128 B between K chunks, SBO = K_pad*16 B between 8-row groups);
//   * the weights (BN scale folded) are laid out the same way once per CTA;
//   * one thread issues K_pad/16 tcgen05.mma (M=128, N=c_out, fp32 accumulate in TMEM);
//   * the same 128 threads read the accumulator back (tcgen05.ld), add bias, ReLU, convert, transpose
//     through shared memory and write NHWC fp16 with full-sector stores.
// Persistent CTAs (several per SM, 64 TMEM columns each) hide the per-tile latency chain.
#include "din_common.cuh"

namespace {

using namespace din;

constexpr int kStemThreads = 128;
constexpr int kStemMaxK = 160;           // 7x7x3 = 147 -> 160
constexpr int kEpiPitch = 80;            // bytes per pixel row in the store-transpose scratch

struct StemParams {
  const float* x; const float* w; const float* bias; __half* y;
  int n, h, w_in, oh, ow, c_out, kh, kw, stride, pad, relu, prep;
  int k_real, k_pad;
  int strips_per_row;     // ceil(ow / 128)
  int patch_w;            // input columns one strip needs: 127*stride + kw
  int patch_pitch;        // patch_w rounded up to odd (bank spread for strided reads)
  int num_tiles;          // n * oh * strips_per_row
};

// canonical no-swizzle K-major layout: element (row r, k) of an [rows x k_pad] operand
__device__ __forceinline__ uint32_t canon_off(int r, int kc /*16-byte chunk*/, int sbo_bytes) {
  return static_cast<uint32_t>((r >> 3) * sbo_bytes + kc * 128 + (r & 7) * 16);
}

__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1u) << 46;   // descriptor version (Blackwell); layout type 0 = no swizzle
  return d;
}

__device__ __forceinline__ void tmem_ld_32x32b_x32_stem(uint32_t taddr, uint32_t (&v)[32]) {
  tmem_ld_32x32b_x32(taddr, v);
}

template <int COUT>  // 64 or 32
__global__ void __launch_bounds__(kStemThreads)
stem_tc_kernel(const StemParams p) {
  constexpr int kTmemCols = COUT < 32 ? 32 : COUT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const int sbo = p.k_pad * 16;                       // bytes between 8-row groups
  uint8_t* a_s = smem;                                // 16 groups
  uint8_t* b_s = a_s + 16 * sbo;                      // COUT/8 groups
  uint8_t* scratch = b_s + (COUT / 8) * sbo;          // 4 warps x 32 x 80 B
  int* koff = reinterpret_cast<int*>(scratch + 4 * 32 * kEpiPitch);   // [k_pad] offset into the patch, or -1
  float* bias_s = reinterpret_cast<float*>(koff + kStemMaxK);         // [COUT]
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(bias_s + 64);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(mma_bar + 1);
  float* patch = reinterpret_cast<float*>(tmem_ptr_smem + 4);         // [3][kh][patch_pitch] prep'd input rows

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup: tap table, weights in UMMA layout, bias, barrier, TMEM
  for (int k = tid; k < p.k_pad; k += kStemThreads) {
    int v = -1;
    if (k < p.k_real) {
      const int c = k / (p.kh * p.kw);
      const int r = k - c * p.kh * p.kw;
      const int ky = r / p.kw, kx = r - ky * p.kw;
      v = (c * p.kh + ky) * p.patch_pitch + kx;
    }
    koff[k] = v;
  }
  for (int i = tid; i < COUT * (p.k_pad / 8); i += kStemThreads) {
    const int o = i / (p.k_pad / 8), kc = i - o * (p.k_pad / 8);
    __align__(16) __half hv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kc * 8 + e;
      // OIHW [c_out][3][kh][kw] flattened == [c_out][k] with k = (c*kh + ky)*kw + kx
      hv[e] = __float2half_rn(k < p.k_real ? __ldg(p.w + static_cast<size_t>(o) * p.k_real + k) : 0.0f);
    }
    *reinterpret_cast<uint4*>(b_s + canon_off(o, kc, sbo)) = *reinterpret_cast<const uint4*>(hv);
  }
  if (tid < COUT) bias_s[tid] = p.bias ? __ldg(p.bias + tid) : 0.0f;
  if (tid == 0) {
    mbar_init(mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<kTmemCols>(tmem_ptr_smem);
  fence_proxy_async_smem();       // generic-proxy smem writes (b_s) -> visible to the tensor core (async proxy)
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);   // warp-uniform for UTCHMMA / LDTM
  const uint32_t idesc = umma_idesc_f16_f32(128, COUT);
  const size_t plane = static_cast<size_t>(p.h) * p.w_in;

  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const int strip = tile % p.strips_per_row;
    const int row = tile / p.strips_per_row;
    const int oy = row % p.oh;
    const int img = row / p.oh;
    const int ox = strip * 128 + tid;
    const float* xi = p.x + static_cast<size_t>(img) * 3 * plane;
    const int iy0 = oy * p.stride - p.pad;
    (void)ox;

    // ---- stage the strip's input patch: coalesced along x, prep_images applied once per input element,
    //      zero where the convolution pads
    {
      const int rows = 3 * p.kh;
      const int gx0 = strip * 128 * p.stride - p.pad;
      for (int i = tid; i < rows * p.patch_w; i += kStemThreads) {
        const int rr = i / p.patch_w, px = i - rr * p.patch_w;
        const int c = rr / p.kh, ky = rr - c * p.kh;
        const int gy = iy0 + ky, gx = gx0 + px;
        float v = 0.0f;
        if (gy >= 0 && gy < p.h && gx >= 0 && gx < p.w_in) {
          v = __ldg(xi + c * plane + static_cast<size_t>(gy) * p.w_in + gx);
          // prep_images (utils.py:14-17): (x/255 - 0.5)*2; the product by 1/255 differs from the division by
          // at most 1 ulp(fp32), far below the fp16 rounding applied next
          if (p.prep) v = (v * (1.0f / 255.0f) - 0.5f) * 2.0f;
        }
        patch[rr * p.patch_pitch + px] = v;
      }
    }
    __syncthreads();
    // ---- im2col row of this thread's pixel -> A tile (canonical UMMA layout)
    {
      const float* prow = patch + tid * p.stride;
      for (int kc = 0; kc < p.k_pad / 8; ++kc) {
        __align__(16) __half hv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int ko = koff[kc * 8 + e];
          hv[e] = __float2half_rn(ko >= 0 ? prow[ko] : 0.0f);
        }
        *reinterpret_cast<uint4*>(a_s + canon_off(tid, kc, sbo)) = *reinterpret_cast<const uint4*>(hv);
      }
    }
    fence_proxy_async_smem();
    __syncthreads();

    // ---- MMA: warp 0 walks the loop uniformly, one elected lane issues K_pad/16 instructions; each consumes
    //      two adjacent 16-byte K chunks
    if (warp == 0) {
      tc_fence_after_sync();
      const uint32_t a_addr = smem_u32(a_s), b_addr = smem_u32(b_s);
      if (elect_one()) {
        for (int ks = 0; ks < p.k_pad / 16; ++ks) {
          const uint64_t ad = desc_noswz(a_addr + ks * 256, 128, sbo);
          const uint64_t bd = desc_noswz(b_addr + ks * 256, 128, sbo);
          umma_f16_ss(tmem_base, ad, bd, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(mma_bar, phase);
    phase ^= 1u;
    tc_fence_after_sync();

    // ---- epilogue: TMEM -> bias/ReLU -> fp16 -> transpose -> NHWC stores (64 contiguous bytes / 4 lanes)
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    uint8_t* sc = scratch + warp * 32 * kEpiPitch;
    const int unit = lane & 3;
#pragma unroll 1
    for (int c0 = 0; c0 < COUT; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32_stem(taddr + c0, v);
      tmem_ld_wait();
      __half2 hh[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float f0 = __uint_as_float(v[2 * j]) + bias_s[c0 + 2 * j];
        float f1 = __uint_as_float(v[2 * j + 1]) + bias_s[c0 + 2 * j + 1];
        if (p.relu) { f0 = fmaxf(f0, 0.0f); f1 = fmaxf(f1, 0.0f); }
        hh[j] = __floats2half2_rn(f0, f1);
      }
      uint4* wr = reinterpret_cast<uint4*>(sc + lane * kEpiPitch);
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) wr[u4] = *reinterpret_cast<uint4*>(&hh[4 * u4]);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int src = (lane >> 2) + 8 * k;               // pixel slot within this warp's 32 pixels
        const int sx = strip * 128 + warp * 32 + src;
        if (sx < p.ow) {
          const uint4 o = *reinterpret_cast<const uint4*>(sc + src * kEpiPitch + unit * 16);
          __half* yp = p.y + ((static_cast<size_t>(img) * p.oh + oy) * p.ow + sx) * COUT + c0 + unit * 8;
          *reinterpret_cast<uint4*>(yp) = o;
        }
      }
      __syncwarp();
    }
    tc_fence_before_sync();
    __syncthreads();     // every warp has drained TMEM and a_s may be rebuilt
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace

int din_stem_tc_launch(const float* x, const float* w, const float* bias, void* y, int n, int h, int w_in,
                       int c_out, int kh, int kw, int stride, int pad, int relu, int prep, cudaStream_t st) {
  StemParams p{};
  p.x = x; p.w = w; p.bias = bias; p.y = static_cast<__half*>(y);
  p.n = n; p.h = h; p.w_in = w_in;
  p.oh = (h + 2 * pad - kh) / stride + 1;
  p.ow = (w_in + 2 * pad - kw) / stride + 1;
  p.c_out = c_out; p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad; p.relu = relu; p.prep = prep;
  p.k_real = 3 * kh * kw;
  p.k_pad = (p.k_real + 15) / 16 * 16;
  p.strips_per_row = (p.ow + 127) / 128;
  p.patch_w = 127 * stride + kw;
  p.patch_pitch = p.patch_w | 1;
  const long long tiles = static_cast<long long>(n) * p.oh * p.strips_per_row;
  if (tiles >= INT32_MAX) return din_set_error(DIN_ERR_INVALID_ARG, "din_stem_conv_nchw_f32: too many tiles");
  p.num_tiles = static_cast<int>(tiles);
  const int sbo = p.k_pad * 16;
  const size_t smem = 128 + static_cast<size_t>(16 + c_out / 8) * sbo + 4 * 32 * kEpiPitch + kStemMaxK * 4 + 64 * 4 + 32 +
                      static_cast<size_t>(3) * kh * p.patch_pitch * 4;
  const int sms = din_num_sms();
  int per_sm = static_cast<int>((200 * 1024) / smem);
  const int tmem_limit = 512 / (c_out < 32 ? 32 : c_out);
  if (per_sm > tmem_limit) per_sm = tmem_limit;
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  long long grid = static_cast<long long>(sms > 0 ? sms : 148) * per_sm;
  if (grid > tiles) grid = tiles;
  if (c_out == 64) {
    DIN_CHECK_CUDA(cudaFuncSetAttribute(stem_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    stem_tc_kernel<64><<<static_cast<int>(grid), kStemThreads, smem, st>>>(p);
  } else {
    DIN_CHECK_CUDA(cudaFuncSetAttribute(stem_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    stem_tc_kernel<32><<<static_cast<int>(grid), kStemThreads, smem, st>>>(p);
  }
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}
