// stem_tc.cu — the 3-channel stem convolution on the tensor cores.
//
// Replaces prep_images (utils.py:8-19) + the first conv of each backbone (VGG-16 features.0 3x3 s1 p1,
// ResNet-18 conv1 7x7 s2 p3 + BN, Inception-v3 Conv2d_1a_3x3 s2 p0 + BN; backbone.py:88-99,115-132,44).
//
// The layer is HBM-bound (VGG: 11 MB fp32 in, 118 MB fp16 out per frame, 3.2 GFLOP) but as a CUDA-core
// direct conv it cost 14.5 ms of a 61 ms step.  Here K = 3*kh*kw (27 / 147) is padded to a multiple of 16
// and each CTA turns a strip of 128 output pixels of one image row into one M=128 UMMA tile:
//   * the strip's input patch (3 x kh rows x (127*stride + kw) columns) is staged in shared memory with
//     coalesced loads, all in flight at once, prep_images applied once per input element, the value that
//     preps to 0 where the convolution pads;
//   * 128 threads = 128 pixels: each thread converts its K patch values to fp16 and writes its row of the
//     A tile straight into the canonical no-swizzle K-major UMMA layout (8 x 16-byte core matrices;
//     LBO = 128 B between K chunks, SBO = K_pad*16 B between 8-row groups); filter geometry is a template
//     parameter so every offset is an immediate (runtime divisions made the first version 5x slower);
//   * the weights (BN scale folded) are laid out the same way once per CTA;
//   * one elected lane issues K_pad/16 tcgen05.mma (M=128, N=c_out, fp32 accumulate in TMEM);
//   * the same 128 threads read the accumulator back (tcgen05.ld), add bias, ReLU, convert, transpose
//     through shared memory and write NHWC fp16 with full-sector stores.
// Persistent CTAs (several per SM, <= 64 TMEM columns each) hide the per-tile latency chain.
//
// Where it stands (tools/stem_ab.py, B200): the layer writes 10.7x the bytes it reads, and a WRITE-ONLY stream on
// this part reaches 3.95 TB/s (torch fill of the same 1.9 GB; the 6.5 TB/s figure is a copy, half reads) -- the
// kernel runs at 3.0-3.2 TB/s = 77-82 % of that.  Two restructurings were built, measured and dropped: (i) loading
// the next tile's patch into registers early (+8 registers cost a resident CTA: -7 %); (ii) a warp-specialised
// version (producer / epilogue warps, double-buffered A tile and accumulator, 256-bit per-lane stores instead of
// the shared-memory transpose): -9 % -- per-lane 32-byte stores to 32 different lines are slower than the
// transposed 64-byte-contiguous ones.
#include <cstdlib>

#include "din_common.cuh"

namespace {

using namespace din;

constexpr int kStemThreads = 128;
constexpr int kEpiPitch = 80;            // bytes per pixel row in the store-transpose scratch

struct StemParams {
  int pool_ph, pool_pw;   // stem_s2d_ws_kernel<U8, POOL = true>: extent of the MaxPool2d(3, 2, 1) output that y holds
  int pool_strips;        //   ceil(ow / 128): strips of 128 conv = 64 pooled pixels (the first one shared with the strip to the left)
  const void* x;          // fp32 NCHW [n,3,h,w]  or (U8) uint8 NHWC [n,h,w,3]
  const float* w; const float* bias; __half* y;
  int n, h, w_in, oh, ow, pad, relu, prep;
  int strips_per_row;     // ceil(ow / 128)
  int num_tiles;          // n * oh * strips_per_row
  uint32_t fd_mul, fd_shr;  // fast division by strips_per_row
  uint32_t fo_mul, fo_shr;  // fast division by oh
};

template <int COUT, int KH, int KW, int STRIDE>
struct StemCfg {
  static constexpr int kReal = 3 * KH * KW;
  static constexpr int kBiasK = kReal;                       // K columns kReal, kReal+1: A = 1, B = bias hi / lo
  static constexpr int kPad = (kReal + 2 + 15) / 16 * 16;
  static constexpr int kSbo = kPad * 16;                      // bytes between 8-row groups
  static constexpr int kPatchW = 127 * STRIDE + KW;
  static constexpr int kPitch = kPatchW | 1;                  // odd pitch: bank spread for strided reads
  static constexpr int kRows = 3 * KH;
  static constexpr int kLoadsPerRow = (kPatchW + kStemThreads - 1) / kStemThreads;
  static constexpr int kTmemCols = COUT < 32 ? 32 : COUT;
  static constexpr size_t kSmem = 128 + static_cast<size_t>(16 + COUT / 8) * kSbo + 4 * 32 * kEpiPitch + 64 * 4 + 32 +
                                  static_cast<size_t>(kRows) * kPitch * 4;
};

// Input patch of one tile held in registers between the global loads and the shared-memory staging (all loads
// in flight at once).
//   fp32 NCHW: one float per (channel, filter row, column slot);  uint8 NHWC: the pixel's 3 bytes packed in
//   one word per (filter row, column slot), bit 24 set = padding.
template <int COUT, int KH, int KW, int STRIDE, bool U8>
struct PatchRegs {
  using Cfg = StemCfg<COUT, KH, KW, STRIDE>;
  static constexpr int kN = U8 ? KH * Cfg::kLoadsPerRow : Cfg::kRows * Cfg::kLoadsPerRow;
  uint32_t v[kN];
  uint32_t valid;     // fp32 path: bit (ky * kLoadsPerRow + u) = that (row, column slot) is inside the image
};

struct TileCoord { int img, oy, strip; };

__device__ __forceinline__ TileCoord decode_tile(const StemParams& p, int tile) {
  const int row = (__umulhi(tile, p.fd_mul) + tile) >> p.fd_shr;            // tile / strips_per_row
  TileCoord t;
  t.strip = tile - row * p.strips_per_row;
  t.img = (__umulhi(row, p.fo_mul) + row) >> p.fo_shr;                      // row / oh
  t.oy = row - t.img * p.oh;
  return t;
}

template <int COUT, int KH, int KW, int STRIDE, bool U8>
__device__ __forceinline__ void load_patch(const StemParams& p, const TileCoord& t, int tid,
                                           PatchRegs<COUT, KH, KW, STRIDE, U8>& r) {
  // Every load is a PREDICATED load with no arithmetic on its result in this function: the compiler then issues
  // all of them back to back (one exposed memory latency per tile).  Doing prep_images right at the load put
  // each LDG in its own branch and serialised 18 latencies: the kernel ran 60 % slower (measured).
  using Cfg = StemCfg<COUT, KH, KW, STRIDE>;
  const int iy0 = t.oy * STRIDE - p.pad;
  const int gx0 = t.strip * 128 * STRIDE - p.pad;
  const size_t plane = static_cast<size_t>(p.h) * p.w_in;
  r.valid = 0u;
  if constexpr (U8) {
    const uint8_t* xi = static_cast<const uint8_t*>(p.x) + static_cast<size_t>(t.img) * 3 * plane;
#pragma unroll
    for (int ky = 0; ky < KH; ++ky) {
      const int gy = iy0 + ky;
      const bool row_ok = gy >= 0 && gy < p.h;
      const uint8_t* src = xi + static_cast<size_t>(row_ok ? gy : 0) * p.w_in * 3;
#pragma unroll
      for (int u = 0; u < Cfg::kLoadsPerRow; ++u) {
        const int px = tid + u * kStemThreads;
        const int gx = gx0 + px;
        const bool ok = row_ok && px < Cfg::kPatchW && gx >= 0 && gx < p.w_in;
        const uint8_t* q = src + static_cast<size_t>(ok ? gx : 0) * 3;
        const uint32_t b0 = ok ? __ldg(q) : 0u, b1 = ok ? __ldg(q + 1) : 0u, b2 = ok ? __ldg(q + 2) : 0u;
        r.v[ky * Cfg::kLoadsPerRow + u] = b0 | (b1 << 8) | (b2 << 16) | (ok ? 0u : 0x01000000u);
      }
    }
  } else {
    const float* xi = static_cast<const float*>(p.x) + static_cast<size_t>(t.img) * 3 * plane;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int ky = 0; ky < KH; ++ky) {
        const int gy = iy0 + ky;
        const bool row_ok = gy >= 0 && gy < p.h;
        const float* src = xi + c * plane + static_cast<size_t>(row_ok ? gy : 0) * p.w_in;
#pragma unroll
        for (int u = 0; u < Cfg::kLoadsPerRow; ++u) {
          const int px = tid + u * kStemThreads;
          const int gx = gx0 + px;
          const bool ok = row_ok && px < Cfg::kPatchW && gx >= 0 && gx < p.w_in;
          // padding: the raw value that preps to EXACTLY 0 under prep_value's FMA is not representable, so the
          // pad is a plain 0 here and stage_patch skips prep for it (prep == 0: a 0 is a 0 anyway)
          r.v[(c * KH + ky) * Cfg::kLoadsPerRow + u] = __float_as_uint(ok ? __ldg(src + gx) : 0.0f);
          if (c == 0) r.valid |= (ok ? 1u : 0u) << (ky * Cfg::kLoadsPerRow + u);
        }
      }
    }
  }
}

// registers -> shared-memory patch [3*KH][kPitch] fp32, prep_images applied once per input element
// (prep_value).  uint8 pixels convert exactly, so both ingest paths agree bit for bit on equal pixel values.
template <int COUT, int KH, int KW, int STRIDE, bool U8>
__device__ __forceinline__ void stage_patch(const StemParams& p, int tid, float* patch,
                                            const PatchRegs<COUT, KH, KW, STRIDE, U8>& r) {
  using Cfg = StemCfg<COUT, KH, KW, STRIDE>;
  if constexpr (U8) {
#pragma unroll
    for (int ky = 0; ky < KH; ++ky) {
#pragma unroll
      for (int u = 0; u < Cfg::kLoadsPerRow; ++u) {
        const int px = tid + u * kStemThreads;
        if (px < Cfg::kPatchW) {
          const uint32_t w = r.v[ky * Cfg::kLoadsPerRow + u];
          const bool is_pad = (w >> 24) != 0;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float f = static_cast<float>((w >> (8 * c)) & 0xFFu);
            patch[(c * KH + ky) * Cfg::kPitch + px] = is_pad ? 0.0f : prep_value(f, p.prep != 0);
          }
        }
      }
    }
  } else {
#pragma unroll
    for (int rr = 0; rr < Cfg::kRows; ++rr) {
#pragma unroll
      for (int u = 0; u < Cfg::kLoadsPerRow; ++u) {
        const int px = tid + u * kStemThreads;
        const float f = __uint_as_float(r.v[rr * Cfg::kLoadsPerRow + u]);
        const bool ok = (r.valid >> ((rr % KH) * Cfg::kLoadsPerRow + u)) & 1u;      // same for the 3 channels
        if (px < Cfg::kPatchW) patch[rr * Cfg::kPitch + px] = ok ? prep_value(f, p.prep != 0) : 0.0f;
      }
    }
  }
}

template <int COUT, int KH, int KW, int STRIDE, bool U8>
__global__ void __launch_bounds__(kStemThreads)
stem_tc_kernel(const StemParams p) {
  using Cfg = StemCfg<COUT, KH, KW, STRIDE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem(smem_raw, 128);
  uint8_t* a_s = smem;                                      // 16 row groups
  uint8_t* b_s = a_s + 16 * Cfg::kSbo;                      // COUT/8 row groups
  uint8_t* scratch = b_s + (COUT / 8) * Cfg::kSbo;          // 4 warps x 32 x 80 B
  float* bias_s = reinterpret_cast<float*>(scratch + 4 * 32 * kEpiPitch);   // [COUT]
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(bias_s + 64);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(mma_bar + 1);
  float* patch = reinterpret_cast<float*>(tmem_ptr_smem + 4);               // [3*KH][kPitch] prep'd input rows

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup: weights in UMMA layout, bias, barrier, TMEM
  for (int i = tid; i < COUT * (Cfg::kPad / 8); i += kStemThreads) {
    const int o = i / (Cfg::kPad / 8), kc = i % (Cfg::kPad / 8);
    __align__(16) __half hv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kc * 8 + e;
      // OIHW [c_out][3][kh][kw] flattened == [c_out][k] with k = (c*kh + ky)*kw + kx.  The bias rides in two
      // extra K columns (hi + lo fp16 parts, multiplied by A = 1): added exactly, in the fp32 accumulator.
      float wv = 0.0f;
      if (k < Cfg::kReal) {
        wv = __ldg(p.w + static_cast<size_t>(o) * Cfg::kReal + k);
      } else if (p.bias != nullptr && k <= Cfg::kBiasK + 1) {
        const float bv = __ldg(p.bias + o);
        const float bh = __half2float(__float2half_rn(bv));
        wv = (k == Cfg::kBiasK) ? bh : bv - bh;
      }
      hv[e] = __float2half_rn(wv);
    }
    *reinterpret_cast<uint4*>(b_s + canon_off(o, kc, Cfg::kSbo)) = *reinterpret_cast<const uint4*>(hv);
  }
  if (tid == 0) {
    mbar_init(mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem);
  fence_proxy_async_smem();       // generic-proxy smem writes (b_s) -> visible to the tensor core (async proxy)
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);   // warp-uniform for UTCHMMA / LDTM
  constexpr uint32_t idesc = umma_idesc_f16_f32(128, COUT);

  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const TileCoord cur = decode_tile(p, tile);
    const int img = cur.img, oy = cur.oy, strip = cur.strip;
    {
      // all loads of the patch are issued before the first use.  (Prefetching the NEXT tile's patch into
      // registers was measured: +8 registers cost a resident CTA per SM and the kernel ran 7 % slower.)
      PatchRegs<COUT, KH, KW, STRIDE, U8> regs;
      load_patch<COUT, KH, KW, STRIDE, U8>(p, cur, tid, regs);
      stage_patch<COUT, KH, KW, STRIDE, U8>(p, tid, patch, regs);
    }
    __syncthreads();
    // ---- im2col row of this thread's pixel -> A tile (canonical UMMA layout); all offsets are immediates
    {
      const float* prow = patch + tid * STRIDE;
#pragma unroll
      for (int kc = 0; kc < Cfg::kPad / 8; ++kc) {
        __align__(16) __half2 hv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float f[2];
#pragma unroll
          for (int z = 0; z < 2; ++z) {
            const int k = kc * 8 + 2 * e + z;
            if (k < Cfg::kReal) {
              const int c = k / (KH * KW), r = k % (KH * KW);
              f[z] = prow[(c * KH + r / KW) * Cfg::kPitch + r % KW];
            } else {
              f[z] = (k <= Cfg::kBiasK + 1) ? 1.0f : 0.0f;       // the two bias columns
            }
          }
          hv[e] = __floats2half2_rn(f[0], f[1]);
        }
        *reinterpret_cast<uint4*>(a_s + canon_off(tid, kc, Cfg::kSbo)) = *reinterpret_cast<const uint4*>(hv);
      }
    }
    fence_proxy_async_smem();
    __syncthreads();

    // ---- MMA: warp 0, one elected lane issues K_pad/16 instructions; each consumes two adjacent 16-byte
    //      K chunks (descriptors are warp-uniform)
    if (warp == 0) {
      tc_fence_after_sync();
      const uint32_t a_addr = smem_u32(a_s), b_addr = smem_u32(b_s);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < Cfg::kPad / 16; ++ks) {
          const uint64_t ad = desc_noswz(a_addr + ks * 256, 128, Cfg::kSbo);
          const uint64_t bd = desc_noswz(b_addr + ks * 256, 128, Cfg::kSbo);
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
    __half* yrow = p.y + ((static_cast<size_t>(img) * p.oh + oy) * p.ow) * COUT;
#pragma unroll
    for (int c0 = 0; c0 < COUT; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr + c0, v);
      tmem_ld_wait();
      uint32_t hh[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        hh[j] = pack_half2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]), p.relu != 0);
      uint4* wr = reinterpret_cast<uint4*>(sc + lane * kEpiPitch);
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) wr[u4] = make_uint4(hh[4 * u4], hh[4 * u4 + 1], hh[4 * u4 + 2], hh[4 * u4 + 3]);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int src = (lane >> 2) + 8 * k;               // pixel slot within this warp's 32 pixels
        const int sx = strip * 128 + warp * 32 + src;
        if (sx < p.ow) {
          const uint4 o = *reinterpret_cast<const uint4*>(sc + src * kEpiPitch + unit * 16);
          *reinterpret_cast<uint4*>(yrow + static_cast<size_t>(sx) * COUT + c0 + unit * 8) = o;
        }
      }
      __syncwarp();
    }
    tc_fence_before_sync();
    __syncthreads();     // every warp has drained TMEM; a_s and the patch may be rebuilt
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// Wide variant for the large-K stems (ResNet-18 7x7 s2: K = 147 -> 160): 256 threads per CTA.  The 128-thread kernel
// spends its time in the im2col build (147 shared loads + 74 conversions + 20 stores per thread, 2 CTAs = 8 warps per
// SM: latency-bound, 0.8 TB/s of output).  Here two threads share a pixel (each builds half of the K chunks), the
// patch staging is spread over 256 threads, and the eight warps split the epilogue by (TMEM lane quadrant, 32-column
// chunk) -- twice the warps per SM for the same shared memory.
// ------------------------------------------------------------------------------------------------
// chunks [KC0, KC1) of pixel `pix`'s im2col row -> A tile (canonical no-swizzle layout); all offsets immediates
template <int COUT, int KH, int KW, int STRIDE, int KC0, int KC1>
__device__ __forceinline__ void build_im2col_chunks(const float* prow, uint8_t* a_s, int pix) {
  using Cfg = StemCfg<COUT, KH, KW, STRIDE>;
#pragma unroll
  for (int kc = KC0; kc < KC1; ++kc) {
    __align__(16) __half2 hv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float f[2];
#pragma unroll
      for (int z = 0; z < 2; ++z) {
        const int k = kc * 8 + 2 * e + z;
        if (k < Cfg::kReal) {
          const int c = k / (KH * KW), r = k % (KH * KW);
          f[z] = prow[(c * KH + r / KW) * Cfg::kPitch + r % KW];
        } else {
          f[z] = (k <= Cfg::kBiasK + 1) ? 1.0f : 0.0f;
        }
      }
      hv[e] = __floats2half2_rn(f[0], f[1]);
    }
    *reinterpret_cast<uint4*>(a_s + canon_off(pix, kc, Cfg::kSbo)) = *reinterpret_cast<const uint4*>(hv);
  }
}

constexpr int kWideThreads = 256;

template <int COUT, int KH, int KW, int STRIDE, bool U8>
__global__ void __launch_bounds__(kWideThreads)
stem_tc_wide_kernel(const StemParams p) {
  using Cfg = StemCfg<COUT, KH, KW, STRIDE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem(smem_raw, 128);
  uint8_t* a_s = smem;
  uint8_t* b_s = a_s + 16 * Cfg::kSbo;
  uint8_t* scratch = b_s + (COUT / 8) * Cfg::kSbo;          // 8 warps x 32 x 80 B
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(scratch + 8 * 32 * kEpiPitch);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(mma_bar + 1);
  float* patch = reinterpret_cast<float*>(tmem_ptr_smem + 4);               // [3*KH][kPitch]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < COUT * (Cfg::kPad / 8); i += kWideThreads) {
    const int o = i / (Cfg::kPad / 8), kc = i % (Cfg::kPad / 8);
    __align__(16) __half hv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kc * 8 + e;
      float wv = 0.0f;
      if (k < Cfg::kReal) {
        wv = __ldg(p.w + static_cast<size_t>(o) * Cfg::kReal + k);
      } else if (p.bias != nullptr && k <= Cfg::kBiasK + 1) {
        const float bv = __ldg(p.bias + o);
        const float bh = __half2float(__float2half_rn(bv));
        wv = (k == Cfg::kBiasK) ? bh : bv - bh;
      }
      hv[e] = __float2half_rn(wv);
    }
    *reinterpret_cast<uint4*>(b_s + canon_off(o, kc, Cfg::kSbo)) = *reinterpret_cast<const uint4*>(hv);
  }
  if (tid == 0) {
    mbar_init(mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  constexpr uint32_t idesc = umma_idesc_f16_f32(128, COUT);
  const size_t plane = static_cast<size_t>(p.h) * p.w_in;
  constexpr int kElems = (U8 ? KH : Cfg::kRows) * Cfg::kPatchW;             // loads per tile (a uint8 load = 3 channels)
  constexpr int kIter = (kElems + kWideThreads - 1) / kWideThreads;

  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const TileCoord cur = decode_tile(p, tile);
    const int iy0 = cur.oy * STRIDE - p.pad;
    const int gx0 = cur.strip * 128 * STRIDE - p.pad;
    // ---- stage the patch: predicated loads first (all in flight), prep + store afterwards
    {
      uint32_t v[kIter];
      bool ok[kIter];
#pragma unroll
      for (int u = 0; u < kIter; ++u) {
        const int i = tid + u * kWideThreads;
        const int r = i / Cfg::kPatchW, px = i - r * Cfg::kPatchW;          // r: (c, ky) row, or ky for uint8
        const int ky = U8 ? r : r % KH;
        const int gy = iy0 + ky, gx = gx0 + px;
        ok[u] = i < kElems && gy >= 0 && gy < p.h && gx >= 0 && gx < p.w_in;
        const int cy = min(max(gy, 0), p.h - 1), cx = min(max(gx, 0), p.w_in - 1);
        if constexpr (U8) {
          const uint8_t* q = static_cast<const uint8_t*>(p.x) + (static_cast<size_t>(cur.img) * plane + static_cast<size_t>(cy) * p.w_in + cx) * 3;
          v[u] = static_cast<uint32_t>(__ldg(q)) | (static_cast<uint32_t>(__ldg(q + 1)) << 8) | (static_cast<uint32_t>(__ldg(q + 2)) << 16);
        } else {
          const int c = r / KH;
          v[u] = __float_as_uint(__ldg(static_cast<const float*>(p.x) + (static_cast<size_t>(cur.img) * 3 + min(c, 2)) * plane +
                                       static_cast<size_t>(cy) * p.w_in + cx));
        }
      }
#pragma unroll
      for (int u = 0; u < kIter; ++u) {
        const int i = tid + u * kWideThreads;
        if (i < kElems) {
          const int r = i / Cfg::kPatchW, px = i - r * Cfg::kPatchW;
          if constexpr (U8) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
              patch[(c * KH + r) * Cfg::kPitch + px] =
                  ok[u] ? prep_value(static_cast<float>((v[u] >> (8 * c)) & 0xFFu), p.prep != 0) : 0.0f;
          } else {
            patch[r * Cfg::kPitch + px] = ok[u] ? prep_value(__uint_as_float(v[u]), p.prep != 0) : 0.0f;
          }
        }
      }
    }
    __syncthreads();
    // ---- im2col: threads t and t + 128 share pixel t & 127 and build half of its K chunks each (two statically
    //      unrolled instantiations, so every patch offset stays an immediate)
    {
      const int pix = tid & 127;
      constexpr int kChunks = Cfg::kPad / 8, kHalfChunks = (kChunks + 1) / 2;
      if (tid < 128) build_im2col_chunks<COUT, KH, KW, STRIDE, 0, kHalfChunks>(patch + pix * STRIDE, a_s, pix);
      else build_im2col_chunks<COUT, KH, KW, STRIDE, kHalfChunks, kChunks>(patch + pix * STRIDE, a_s, pix);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after_sync();
      const uint32_t a_addr = smem_u32(a_s), b_addr = smem_u32(b_s);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < Cfg::kPad / 16; ++ks) {
          const uint64_t ad = desc_noswz(a_addr + ks * 256, 128, Cfg::kSbo);
          const uint64_t bd = desc_noswz(b_addr + ks * 256, 128, Cfg::kSbo);
          umma_f16_ss(tmem_base, ad, bd, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(mma_bar, phase);
    phase ^= 1u;
    tc_fence_after_sync();
    // ---- epilogue: warp = (lane quadrant q, 32-column chunk)
    {
      const int q = warp & 3, chunk = warp >> 2;
      const int c0 = chunk * 32;
      if (c0 < COUT) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        uint8_t* sc = scratch + warp * 32 * kEpiPitch;
        const int unit = lane & 3;
        __half* yrow = p.y + ((static_cast<size_t>(cur.img) * p.oh + cur.oy) * p.ow) * COUT;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        tmem_ld_wait();
        uint32_t hh[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          hh[j] = pack_half2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]), p.relu != 0);
        uint4* wr = reinterpret_cast<uint4*>(sc + lane * kEpiPitch);
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) wr[u4] = make_uint4(hh[4 * u4], hh[4 * u4 + 1], hh[4 * u4 + 2], hh[4 * u4 + 3]);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int src = (lane >> 2) + 8 * k;
          const int sx = cur.strip * 128 + q * 32 + src;
          if (sx < p.ow) {
            const uint4 o = *reinterpret_cast<const uint4*>(sc + src * kEpiPitch + unit * 16);
            *reinterpret_cast<uint4*>(yrow + static_cast<size_t>(sx) * COUT + c0 + unit * 8) = o;
          }
        }
      }
    }
    tc_fence_before_sync();
    __syncthreads();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// ResNet-18's 7x7 stride-2 pad-3 stem as an IMPLICIT GEMM (no im2col): space-to-depth turns it into a 4x4 stride-1
// convolution over 12 channels,
//     out(y, x) = sum_{ty,tx in 0..3} sum_{c,dy,dx} w[c][2ty+dy][2tx+dx] * S[y+ty][x+tx][(c,dy,dx)],
//     S[Y][X][(c,dy,dx)] = prep(in[c][2Y+dy-3][2X+dx-3])          (w index 7 = 0),
// so ONE 16-half vector per half-resolution pixel (12 values + two bias ones + 2 zeros) serves all 16 taps: tap
// (ty, tx)'s A operand is the same shared memory seen through a no-swizzle K-major descriptor whose start is shifted by
// (ty * pitch + tx) rows (rows 16 bytes apart in each of the two K-chunk planes: SBO = 128, LBO = the plane size).
// Per 128-pixel strip: 4 x 131 vectors of 12 values instead of 128 im2col rows of 147 -- the wide im2col kernel above is
// bound by exactly that build (147 shared loads + 74 conversions + 20 stores per pixel: 8.8-9.5 ms per 320 frames at
// 720p, 1.1 TB/s of output) --, 16 MMAs of K = 16 instead of 10.
// ------------------------------------------------------------------------------------------------
constexpr int kS2dPatchW = 127 * 2 + 8;          // 262 input columns per strip
constexpr int kS2dPitch = 264;                   // even: the (dx = 0, 1) pairs are 8-byte aligned
constexpr int kS2dRows = 24;                     // 3 channels x 8 input rows
constexpr int kS2dVecs = 131;                    // half-resolution pixels per patch row (128 + 3)
constexpr int kS2dVPitch = 136;                  // rows between consecutive ty in the vector array
constexpr int kS2dVPlane = 4 * kS2dVPitch * 16;  // bytes of one K-chunk plane
constexpr int kS2dWTap = 2048;                   // bytes of one tap's weight tile (2 planes x 64 rows x 16 B)
constexpr size_t kS2dSmem = 128 + 2 * kS2dVPlane + 16 * kS2dWTap + 8 * 32 * kEpiPitch + 64 +
                            static_cast<size_t>(kS2dRows) * kS2dPitch * 4;

template <bool U8>
__global__ void __launch_bounds__(kWideThreads, 2)
stem_s2d_kernel(const StemParams p) {
  constexpr int COUT = 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem(smem_raw, 128);
  uint8_t* v_s = smem;                                       // vectors: 2 planes x [4 * kS2dVPitch] x 16 B
  uint8_t* w_s = v_s + 2 * kS2dVPlane;                       // 16 taps x (2 planes x 64 rows x 16 B)
  uint8_t* scratch = w_s + 16 * kS2dWTap;                    // 8 warps x 32 x 80 B
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(scratch + 8 * 32 * kEpiPitch);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(mma_bar + 1);
  float* patch = reinterpret_cast<float*>(tmem_ptr_smem + 14);              // [24][kS2dPitch], 8-byte aligned
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time: weights in tap-major [tap][plane][row][8 halfs]; bias (hi, lo) rides in columns 12, 13 of tap (0,0)
  for (int i = tid; i < 16 * 2 * COUT; i += kWideThreads) {
    const int o = i & 63, plane = (i >> 6) & 1, tap = i >> 7;
    const int ty = tap >> 2, tx = tap & 3;
    __align__(16) __half hv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = plane * 8 + e;
      float wv = 0.0f;
      if (k < 12) {
        const int c = k >> 2, ky = 2 * ty + ((k >> 1) & 1), kx = 2 * tx + (k & 1);
        if (ky < 7 && kx < 7) wv = __ldg(p.w + ((static_cast<size_t>(o) * 3 + c) * 7 + ky) * 7 + kx);
      } else if (k <= 13 && tap == 0 && p.bias != nullptr) {
        const float bv = __ldg(p.bias + o);
        const float bh = __half2float(__float2half_rn(bv));
        wv = (k == 12) ? bh : bv - bh;
      }
      hv[e] = __float2half_rn(wv);
    }
    *reinterpret_cast<uint4*>(w_s + tap * kS2dWTap + plane * 1024 + o * 16) = *reinterpret_cast<const uint4*>(hv);
  }
  for (int i = tid; i < 2 * kS2dVPlane / 16; i += kWideThreads) reinterpret_cast<uint4*>(v_s)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<64>(tmem_ptr_smem);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  constexpr uint32_t idesc = umma_idesc_f16_f32(128, COUT);
  const size_t plane_px = static_cast<size_t>(p.h) * p.w_in;
  constexpr int kElems = (U8 ? 8 : kS2dRows) * kS2dPatchW;
  constexpr int kIter = (kElems + kWideThreads - 1) / kWideThreads;

  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const TileCoord cur = decode_tile(p, tile);
    const int iy0 = cur.oy * 2 - 3;
    const int gx0 = cur.strip * 128 * 2 - 3;
    // ---- stage the 8 input rows x 262 columns x 3 channels (prep_images applied, 0 where the convolution pads)
    {
      uint32_t v[kIter];
      bool ok[kIter];
#pragma unroll
      for (int u = 0; u < kIter; ++u) {
        const int i = tid + u * kWideThreads;
        const int r = i / kS2dPatchW, px = i - r * kS2dPatchW;
        const int ky = U8 ? r : (r & 7);
        const int gy = iy0 + ky, gx = gx0 + px;
        ok[u] = i < kElems && gy >= 0 && gy < p.h && gx >= 0 && gx < p.w_in;
        const int cy = min(max(gy, 0), p.h - 1), cx = min(max(gx, 0), p.w_in - 1);
        if constexpr (U8) {
          const uint8_t* q = static_cast<const uint8_t*>(p.x) + (static_cast<size_t>(cur.img) * plane_px + static_cast<size_t>(cy) * p.w_in + cx) * 3;
          v[u] = static_cast<uint32_t>(__ldg(q)) | (static_cast<uint32_t>(__ldg(q + 1)) << 8) | (static_cast<uint32_t>(__ldg(q + 2)) << 16);
        } else {
          const int c = min(r >> 3, 2);
          v[u] = __float_as_uint(__ldg(static_cast<const float*>(p.x) + (static_cast<size_t>(cur.img) * 3 + c) * plane_px +
                                       static_cast<size_t>(cy) * p.w_in + cx));
        }
      }
#pragma unroll
      for (int u = 0; u < kIter; ++u) {
        const int i = tid + u * kWideThreads;
        if (i < kElems) {
          const int r = i / kS2dPatchW, px = i - r * kS2dPatchW;
          if constexpr (U8) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
              patch[(c * 8 + r) * kS2dPitch + px] =
                  ok[u] ? prep_value(static_cast<float>((v[u] >> (8 * c)) & 0xFFu), p.prep != 0) : 0.0f;
          } else {
            patch[r * kS2dPitch + px] = ok[u] ? prep_value(__uint_as_float(v[u]), p.prep != 0) : 0.0f;
          }
        }
      }
    }
    __syncthreads();
    // ---- one 16-half vector per half-resolution pixel (ty, X): (c, dy, dx) values, then the two bias ones
    for (int v = tid; v < 4 * kS2dVecs; v += kWideThreads) {
      const int ty = v / kS2dVecs, X = v - ty * kS2dVecs;
      uint32_t h[8];
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const float2 f = *reinterpret_cast<const float2*>(patch + (c * 8 + 2 * ty + dy) * kS2dPitch + 2 * X);
          h[c * 2 + dy] = pack_half2(f.x, f.y, false);
        }
      h[6] = 0x3C003C00u;                                    // columns 12, 13 = 1 (bias hi / lo live in tap (0,0)'s weights)
      h[7] = 0u;
      uint8_t* dst = v_s + (ty * kS2dVPitch + X) * 16;
      *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(dst + kS2dVPlane) = make_uint4(h[4], h[5], h[6], h[7]);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after_sync();
      const uint32_t v_addr = smem_u32(v_s), w_addr = smem_u32(w_s);
      if (elect_one()) {
#pragma unroll
        for (int tap = 0; tap < 16; ++tap) {
          const uint64_t ad = desc_noswz(v_addr + ((tap >> 2) * kS2dVPitch + (tap & 3)) * 16, kS2dVPlane, 128);
          const uint64_t bd = desc_noswz(w_addr + tap * kS2dWTap, 1024, 128);
          umma_f16_ss(tmem_base, ad, bd, idesc, tap > 0 ? 1u : 0u);
        }
        umma_commit(mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(mma_bar, phase);
    phase ^= 1u;
    tc_fence_after_sync();
    // ---- epilogue: warp = (lane quadrant q, 32-column chunk), as stem_tc_wide_kernel
    {
      const int q = warp & 3, chunk = warp >> 2;
      const int c0 = chunk * 32;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      uint8_t* sc = scratch + warp * 32 * kEpiPitch;
      const int unit = lane & 3;
      __half* yrow = p.y + ((static_cast<size_t>(cur.img) * p.oh + cur.oy) * p.ow) * COUT;
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr + c0, v);
      tmem_ld_wait();
      uint32_t hh[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        hh[j] = pack_half2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]), p.relu != 0);
      uint4* wr = reinterpret_cast<uint4*>(sc + lane * kEpiPitch);
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) wr[u4] = make_uint4(hh[4 * u4], hh[4 * u4 + 1], hh[4 * u4 + 2], hh[4 * u4 + 3]);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int src = (lane >> 2) + 8 * k;
        const int sx = cur.strip * 128 + q * 32 + src;
        if (sx < p.ow) {
          const uint4 o = *reinterpret_cast<const uint4*>(sc + src * kEpiPitch + unit * 16);
          *reinterpret_cast<uint4*>(yrow + static_cast<size_t>(sx) * COUT + c0 + unit * 8) = o;
        }
      }
    }
    tc_fence_before_sync();
    __syncthreads();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<64>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// The same implicit GEMM, warp-specialised and pipelined (the kernel above runs load -> build -> MMA -> store one after
// the other inside a CTA and was no faster than the im2col kernel: 10.5 vs 9.7 ms per 320 frames -- both are bound by
// that serial chain of latencies, not by the im2col build as round 1 assumed):
//   warp 0        TMA: box {264, 8, 3} of the fp32 NCHW image ({832 B, 8} of the uint8 NHWC frame) per tile, 3 tiles
//                 ahead (the box starts on a 16-byte boundary left of the strip: see tools/probes/tma_probe.cu);
//   builders      two groups of eight warps (six with the pool fused: 192 threads walk the 524 vectors in three trips,
//                 exactly as 256 would) build the 4 x 131 half-resolution vectors (prep_images, 0 where the convolution
//                 pads) of alternate tiles, each group into its own vector buffer: a tile costs a group ~1200 cycles of
//                 building plus ~850 of fence.proxy.async / barrier hand-off (per-role cycle counters), which one group
//                 alone cannot hide;
//   warp 1        16 tcgen05.mma (K = 16 each) per tile into one of two TMEM accumulators;
//   epilogue      TMEM -> ReLU -> fp16 -> transpose -> NHWC stores: four warps, one per TMEM lane quadrant; with the pool
//                 fused eight (warp e reads lane quadrant e & 3 and the 32 output channels of half e >> 2: the four-warp
//                 pool epilogue was the longest role, ~1650 of a tile's ~2100 cycles).
// One persistent CTA per SM (164 KB of shared memory, 218 KB with the pool's ring).
// ------------------------------------------------------------------------------------------------
constexpr int kWsThreads = 32 * 22;
// builder warps per GROUP (the first group builds the even tiles of a CTA, the second the odd ones, each into its own
// vector buffer) and epilogue warps: 8 + 8 + 4 without the pool (two epilogue warps per pixel would write its 128-byte
// line as two 64-byte halves at different times: 1.25 -> 1.47 ms), 6 + 6 + 8 with it
template <bool POOL> struct WsRoles {
  static constexpr int kBuilders = POOL ? 6 : 8;
  static constexpr int kEpiWarp0 = 2 + 2 * kBuilders;
  static constexpr int kEpiWarps = POOL ? 8 : 4;
  static_assert(kEpiWarp0 + kEpiWarps == 22 && (kEpiWarp0 & 3) == 2, "warp roles");
};
constexpr int kWsEpiWarps = 8;                   // scratch areas
constexpr int kWsPatchStages = 3;
// fp32: the 264 input columns 256*strip - 4 .. + 259 of a patch row exceed TMA's 256-element box limit, so every tile is
// two boxes: A = columns [0, 136) and B = columns [132, 264) of that range (both start on 16-byte boundaries); the
// (dx = 0, 1) pair of vector X sits at columns 2X, 2X + 1 (8-byte aligned): X <= 65 reads A, X >= 66 reads B.
constexpr int kWsBoxA = 136, kWsBoxB = 132, kWsBoxBStart = 132, kWsSplitX = 66;
constexpr int kWsBlockA = 3 * 8 * kWsBoxA * 4;   // 13056 bytes
constexpr int kWsBoxU8 = 832;                    // uint8: bytes per patch row (pixels 256*strip - 16 .. + 261), loaded as
                                                 // 208 32-bit elements (a box is at most 256 elements wide)
constexpr int kWsPatchBytes = kWsBlockA + 3 * 8 * kWsBoxB * 4;           // 25728 (uint8: 8 * 832 = 6656)
constexpr int kWsPatchStage = 26112;
constexpr size_t kWsSmem = 1024 + kWsPatchStages * kWsPatchStage + 2 * 2 * kS2dVPlane + 16 * kS2dWTap + kWsEpiWarps * 32 * kEpiPitch + 256;
// POOL: resnet18.maxpool (MaxPool2d(3, 2, 1), backbone.py:115-132) fused into the stem.  The kernel is bound by its OUTPUT
// writes (3.2 GB per 107 frames at 720p, 3.1 TB/s of the part's ~3.95 TB/s write-only rate), and the pool throws three
// quarters of them away: here a CTA owns a strip of 128 convolution pixels (conv x = 128 * strip + r, the geometry of the
// kernel without the pool) and walks the rows of one image top to bottom, its epilogue warps keep the last three ReLU'd
// convolution rows in a shared memory ring and emit one pooled row for every second convolution row.  Pooled pixel
// 64 * strip + jl covers ring columns 2 jl - 1 .. 2 jl + 1: jl = 1 .. 63 are complete inside the strip; jl = 0 also needs
// the LAST column of the strip to the left, so for strip > 0 both strips fold their share into y with a 16-byte
// red.global.max (REDG.E.MAX.F16x8) over a pixel that stem_pool_zero_kernel cleared before the launch -- exact: the maximum
// is associative and every ReLU output is >= 0.  (First version: strips of 126 convolution = 63 complete pooled pixels, one
// column overlapping: 1280-wide frames needed SIX strips for 5.08 strips' worth of pixels, 17 % of all tiles nearly empty.)
// Padding: a ReLU output is >= 0 and every window holds a real pixel, so positions outside the image enter the maximum as 0.
constexpr int kWsRingPitch = 144;                // bytes per pixel in the ring: 128 + 16 (bank spread for 16-byte stores)
constexpr int kWsRingRow = 128 * kWsRingPitch;
constexpr size_t kWsPoolSmem = kWsSmem + 3 * kWsRingRow;

// iteration order of every role: POOL = false: tiles blockIdx.x, + gridDim.x, ... decoded as (img, oy, strip);
// POOL = true: tasks (img, strip) = blockIdx.x, + gridDim.x, ..., each walked over oy = 0 .. oh - 1
template <bool POOL>
struct WsIter {
  const StemParams& p;
  int tile, oy_, task;
  __device__ __forceinline__ WsIter(const StemParams& pp) : p(pp), tile(blockIdx.x), oy_(0), task(blockIdx.x) {}
  __device__ __forceinline__ bool valid() const { return POOL ? task < p.n * p.pool_strips : tile < p.num_tiles; }
  __device__ __forceinline__ TileCoord coord() const {
    if constexpr (POOL) {
      TileCoord t;
      t.img = task / p.pool_strips;
      t.strip = task - t.img * p.pool_strips;
      t.oy = oy_;
      return t;
    } else {
      return decode_tile(p, tile);
    }
  }
  __device__ __forceinline__ void next() {
    if constexpr (POOL) {
      if (++oy_ == p.oh) { oy_ = 0; task += gridDim.x; }
    } else {
      tile += gridDim.x;
    }
  }
};

template <bool U8, bool POOL = false>
__global__ void __launch_bounds__(kWsThreads, 1)
stem_s2d_ws_kernel(const __grid_constant__ CUtensorMap tmap_img, const __grid_constant__ CUtensorMap tmap_b,
                   const StemParams p) {
  constexpr int COUT = 64;
  constexpr int kStripStride = 128;                             // convolution pixels between strips
  constexpr int kXShift = 0;
  constexpr int kWsBuilders = WsRoles<POOL>::kBuilders, kWsEpiWarp0 = WsRoles<POOL>::kEpiWarp0;
  constexpr int kNumEpi = WsRoles<POOL>::kEpiWarps;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem(smem_raw, 1024);
  uint8_t* patch_s = smem;                                           // kWsPatchStages x raw patch
  uint8_t* v_s = patch_s + kWsPatchStages * kWsPatchStage;           // 2 x (2 planes x [4 * kS2dVPitch] x 16 B)
  uint8_t* w_s = v_s + 2 * 2 * kS2dVPlane;                           // 16 taps x 2 KB
  uint8_t* scratch = w_s + 16 * kS2dWTap;                            // 8 epilogue warps x 32 x 80 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + kWsEpiWarps * 32 * kEpiPitch);
  uint64_t* patch_full = bars;                    // [3]
  uint64_t* patch_empty = patch_full + 3;         // [3]  the builder warps of one group
  uint64_t* v_full = patch_empty + 3;             // [2]  the builder warps of one group
  uint64_t* v_empty = v_full + 2;                 // [2]  MMA commit
  uint64_t* acc_full = v_empty + 2;               // [2]  MMA commit
  uint64_t* acc_empty = acc_full + 2;             // [2]  8 epilogue warps
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);
  uint8_t* ring = reinterpret_cast<uint8_t*>(bars) + 256;          // POOL: 3 convolution rows x 128 pixels x kWsRingPitch
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time: weights (as stem_s2d_kernel), zeroed vector buffers, barriers, TMEM
  for (int i = tid; i < 16 * 2 * COUT; i += kWsThreads) {
    const int o = i & 63, plane = (i >> 6) & 1, tap = i >> 7;
    const int ty = tap >> 2, tx = tap & 3;
    __align__(16) __half hv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = plane * 8 + e;
      float wv = 0.0f;
      if (k < 12) {
        // columns use the space-to-depth phase -4 (pairs start on even input columns = 8-byte aligned in the patch):
        // S[Y][X][(c,dy,dx)] = in[c][2Y+dy-3][2X+dx-4], so kx = 2*tx + dx - 1 (kx = -1 and ky = 7: zero weights)
        const int c = k >> 2, ky = 2 * ty + ((k >> 1) & 1), kx = 2 * tx + (k & 1) - 1;
        if (ky < 7 && kx >= 0 && kx < 7) wv = __ldg(p.w + ((static_cast<size_t>(o) * 3 + c) * 7 + ky) * 7 + kx);
      } else if (k <= 13 && tap == 0 && p.bias != nullptr) {
        const float bv = __ldg(p.bias + o);
        const float bh = __half2float(__float2half_rn(bv));
        wv = (k == 12) ? bh : bv - bh;
      }
      hv[e] = __float2half_rn(wv);
    }
    *reinterpret_cast<uint4*>(w_s + tap * kS2dWTap + plane * 1024 + o * 16) = *reinterpret_cast<const uint4*>(hv);
  }
  for (int i = tid; i < 2 * 2 * kS2dVPlane / 16; i += kWsThreads) reinterpret_cast<uint4*>(v_s)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    tma_prefetch_desc(&tmap_img);
    for (int s = 0; s < 3; ++s) { mbar_init(&patch_full[s], 1); mbar_init(&patch_empty[s], kWsBuilders); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&v_full[s], kWsBuilders); mbar_init(&v_empty[s], 1);
      mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kNumEpi);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_ptr_smem);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int j = 0;
      for (WsIter<POOL> it(p); it.valid(); it.next(), ++j) {
        const TileCoord cur = it.coord();
        const int ps = j % kWsPatchStages;
        // input column of (X = 0, dx = 0); the box starts at the 16-byte boundary at or left of it
        const int gx0 = (cur.strip * kStripStride - kXShift) * 2 - 4;
        mbar_wait_relaxed(&patch_empty[ps], ((j / kWsPatchStages) & 1) ^ 1);
        if constexpr (U8) {
          const int p0 = gx0 - 12;                                         // first pixel of the box: a multiple of 16
          mbar_arrive_expect_tx(&patch_full[ps], 8u * kWsBoxU8);
          tma_load_3d(patch_s + ps * kWsPatchStage, &tmap_img, &patch_full[ps], (p0 * 3) / 4, cur.oy * 2 - 3, cur.img);
        } else {
          const int c0 = gx0;                                              // 256 * strip - 4: a multiple of 4 columns
          mbar_arrive_expect_tx(&patch_full[ps], static_cast<uint32_t>(kWsPatchBytes));
          tma_load_3d(patch_s + ps * kWsPatchStage, &tmap_img, &patch_full[ps], c0, cur.oy * 2 - 3, cur.img * 3);
          tma_load_3d(patch_s + ps * kWsPatchStage + kWsBlockA, &tmap_b, &patch_full[ps], c0 + kWsBoxBStart, cur.oy * 2 - 3,
                      cur.img * 3);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_f16_f32(128, COUT);
    const uint32_t v_addr = smem_u32(v_s), w_addr = smem_u32(w_s);
    int j = 0;
    for (WsIter<POOL> it(p); it.valid(); it.next(), ++j) {
      const int b = j & 1;
      const uint32_t use = static_cast<uint32_t>(j >> 1);
      mbar_wait(&v_full[b], use & 1u);
      mbar_wait(&acc_empty[b], (use & 1u) ^ 1u);
      tc_fence_after_sync();
      if (leader) {
#pragma unroll
        for (int tap = 0; tap < 16; ++tap) {
          const uint64_t ad = desc_noswz(v_addr + b * (2 * kS2dVPlane) + ((tap >> 2) * kS2dVPitch + (tap & 3)) * 16,
                                         kS2dVPlane, 128);
          const uint64_t bd = desc_noswz(w_addr + tap * kS2dWTap, 1024, 128);
          umma_f16_ss(tmem_base + static_cast<uint32_t>(b * COUT), ad, bd, idesc, tap > 0 ? 1u : 0u);
        }
        umma_commit(&v_empty[b]);
        umma_commit(&acc_full[b]);
      }
    }
  } else if (warp < 2 + 2 * kWsBuilders) {
    // ------------------------------------------------------------------ builders: one vector per half-res pixel
    const int grp = (warp - 2) / kWsBuilders;                         // tiles j with (j & 1) == grp, vector buffer grp
    const int bt = tid - 64 - grp * 32 * kWsBuilders;                 // 0..255 inside the group
    int j = 0;
    for (WsIter<POOL> it(p); it.valid(); it.next(), ++j) {
      if ((j & 1) != grp) continue;
      const TileCoord cur = it.coord();
      const int ps = j % kWsPatchStages, b = j & 1;
      const uint32_t use = static_cast<uint32_t>(j >> 1);
      mbar_wait_relaxed(&patch_full[ps], (j / kWsPatchStages) & 1);
      mbar_wait_relaxed(&v_empty[b], (use & 1u) ^ 1u);
      const uint8_t* patch = patch_s + ps * kWsPatchStage;
      uint8_t* vb = v_s + b * (2 * kS2dVPlane);
      // input coordinates of (ty = 0, dy = 0) / (X = 0, dx = 0)
      const int iy0 = cur.oy * 2 - 3, gx0 = (cur.strip * kStripStride - kXShift) * 2 - 4;
      // offset of input column gx0 inside the loaded box (see the producer)
      const int boff = U8 ? 12 : 0;
      for (int v = bt; v < 4 * kS2dVecs; v += 32 * kWsBuilders) {
        const int ty = v / kS2dVecs, X = v - ty * kS2dVecs;
        const bool in_a = X < kWsSplitX;
        (void)in_a;
        bool rok[2], cok[2];
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          rok[d] = (iy0 + 2 * ty + d >= 0) && (iy0 + 2 * ty + d < p.h);
          cok[d] = (gx0 + 2 * X + d >= 0) && (gx0 + 2 * X + d < p.w_in);
        }
        uint32_t h[8];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int dy = 0; dy < 2; ++dy) {
            float raw[2];
            if constexpr (U8) {
              const uint8_t* q = patch + (2 * ty + dy) * kWsBoxU8 + (boff + 2 * X) * 3 + c;
              raw[0] = static_cast<float>(q[0]);
              raw[1] = static_cast<float>(q[3]);
            } else {
              // one aligned 8-byte load per (c, dy): lanes read consecutive pairs, no bank conflicts (with the pairs on
              // odd columns -- phase -3 -- these were 4-byte loads at stride 2: 790 of the tile's 2430 shared-memory
              // wavefronts were their conflicts)
              const float* row = in_a ? reinterpret_cast<const float*>(patch) + (c * 8 + 2 * ty + dy) * kWsBoxA + 2 * X + boff
                                      : reinterpret_cast<const float*>(patch + kWsBlockA) + (c * 8 + 2 * ty + dy) * kWsBoxB +
                                            2 * X + boff - kWsBoxBStart;
              const float2 pr = *reinterpret_cast<const float2*>(row);
              raw[0] = pr.x;
              raw[1] = pr.y;
            }
            float f[2];
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) f[dx] = (rok[dy] && cok[dx]) ? prep_value(raw[dx], p.prep != 0) : 0.0f;
            h[c * 2 + dy] = pack_half2(f[0], f[1], false);
          }
        h[6] = 0x3C003C00u;
        h[7] = 0u;
        uint8_t* dst = vb + (ty * kS2dVPitch + X) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(dst + kS2dVPlane) = make_uint4(h[4], h[5], h[6], h[7]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&patch_empty[ps]);
        mbar_arrive(&v_full[b]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: warp e = lane quadrant e
    const int q = warp & 3;                                           // warps 14..21 -> quadrants 2, 3, 0, 1, 2, 3, 0, 1
    const int c0 = ((warp - kWsEpiWarp0) >> 2) * 32;                  // POOL: this warp's 32 output channels
    uint8_t* sc = scratch + (warp - kWsEpiWarp0) * 32 * kEpiPitch;
    const int unit = lane & 3;
    int j = 0;
    for (WsIter<POOL> it(p); it.valid(); it.next(), ++j) {
      const TileCoord cur = it.coord();
      const int b = j & 1;
      mbar_wait_relaxed(&acc_full[b], static_cast<uint32_t>(j >> 1) & 1u);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(b * COUT);
      if constexpr (POOL) {
        // this lane's convolution pixel -> the ring (zeros outside the image), then every second row one pooled row
        const int r = q * 32 + lane;
        const int cx = cur.strip * 128 + r;
        const bool inside = cx < p.ow;
        uint8_t* dst = ring + (cur.oy % 3) * kWsRingRow + r * kWsRingPitch;
        {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int u4 = 0; u4 < 4; ++u4) {
            uint32_t hh[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
              hh[k] = inside ? pack_half2(__uint_as_float(v[8 * u4 + 2 * k]), __uint_as_float(v[8 * u4 + 2 * k + 1]), true) : 0u;
            *reinterpret_cast<uint4*>(dst + c0 * 2 + u4 * 16) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[b]);
        const bool last_row = cur.oy == p.oh - 1;
        if ((cur.oy & 1) || last_row) {
          const int py = cur.oy >> 1;                                 // rows 2 py - 1 .. 2 py + 1 (clipped to the image)
          asm volatile("bar.sync 1, 256;" ::: "memory");              // the eight epilogue warps: all ring rows written
          const int et = (warp - kWsEpiWarp0) * 32 + lane;            // 0..255
          __half* yrow = p.y + ((static_cast<size_t>(cur.img) * p.pool_ph + py) * p.pool_pw) * COUT;
          // rows clipped INTO the image (a repeated row leaves the maximum unchanged): nine independent loads in flight
          int slot[3];
#pragma unroll
          for (int d = 0; d < 3; ++d) slot[d] = min(max(2 * py - 1 + d, 0), p.oh - 1) % 3;
          // items 0 .. 511: pooled pixel jl = i >> 3 of the strip, 8 channels vv; items 512 .. 519: the strip's last column
          // (r = 127) folded into pooled pixel 0 of the strip to the right
          for (int i = et; i < 65 * 8; i += 32 * kNumEpi) {
            const int jl = i >> 3, vv = i & 7;
            const int px = cur.strip * 64 + jl;
            if (px < p.pool_pw) {
              const int r0 = jl == 64 ? 127 : max(2 * jl - 1, 0), r1 = jl == 64 ? 127 : 2 * jl + 1;
              __half2 m[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) m[e] = __float2half2_rn(0.0f);
#pragma unroll
              for (int d = 0; d < 3; ++d) {
                const uint8_t* rr = ring + slot[d] * kWsRingRow + vv * 16;
#pragma unroll
                for (int dxp = 0; dxp < 3; ++dxp) {
                  const int r = min(r0 + dxp, r1);                    // a repeated column leaves the maximum unchanged
                  const uint4 t = *reinterpret_cast<const uint4*>(rr + r * kWsRingPitch);
                  const __half2* th = reinterpret_cast<const __half2*>(&t);
#pragma unroll
                  for (int e = 0; e < 4; ++e) m[e] = __hmax2(m[e], th[e]);
                }
              }
              __half* dst = yrow + static_cast<size_t>(px) * COUT + vv * 8;
              const uint4 mv = *reinterpret_cast<const uint4*>(m);
              if (jl == 64 || (jl == 0 && cur.strip > 0))
                asm volatile("red.global.max.noftz.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(mv.x), "r"(mv.y), "r"(mv.z),
                             "r"(mv.w)
                             : "memory");
              else
                *reinterpret_cast<uint4*>(dst) = mv;
            }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");              // ring rows may be overwritten by the next tiles
        }
      } else {
      __half* yrow = p.y + ((static_cast<size_t>(cur.img) * p.oh + cur.oy) * p.ow) * COUT;
#pragma unroll
      for (int c0 = 0; c0 < COUT; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        tmem_ld_wait();
        uint32_t hh[16];
#pragma unroll
        for (int k = 0; k < 16; ++k)
          hh[k] = pack_half2(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1]), p.relu != 0);
        uint4* wr = reinterpret_cast<uint4*>(sc + lane * kEpiPitch);
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) wr[u4] = make_uint4(hh[4 * u4], hh[4 * u4 + 1], hh[4 * u4 + 2], hh[4 * u4 + 3]);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int src = (lane >> 2) + 8 * k;
          const int sx = cur.strip * 128 + q * 32 + src;
          if (sx < p.ow) {
            const uint4 o = *reinterpret_cast<const uint4*>(sc + src * kEpiPitch + unit * 16);
            *reinterpret_cast<uint4*>(yrow + static_cast<size_t>(sx) * COUT + c0 + unit * 8) = o;
          }
        }
        __syncwarp();
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<128>(tmem_base);
  }
}

// pooled pixels 64, 128, ... of every row are reduced into by two strips (red.global.max): cleared first
__global__ void stem_pool_zero_kernel(__half* __restrict__ y, long long rows, int pool_pw, int shared_px) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // (row, shared pixel, 16-byte unit)
  if (i >= rows * shared_px * 8) return;
  const int vv = static_cast<int>(i & 7);
  const long long t = i >> 3;
  const int k = static_cast<int>(t % shared_px);
  const long long row = t / shared_px;
  reinterpret_cast<uint4*>(y + (row * pool_pw + 64 * (k + 1)) * 64)[vv] = make_uint4(0, 0, 0, 0);
}

template <bool U8, bool POOL = false>
int launch_s2d_ws(StemParams& p, cudaStream_t st) {
  CUtensorMap timg, timg_b;
  if (U8) {
    // the uint8 frame as 32-bit elements (w % 16 == 0: rows are multiples of 48 bytes)
    const uint64_t row = static_cast<uint64_t>(p.w_in) * 3;
    const uint64_t dims[3] = {row / 4, static_cast<uint64_t>(p.h), static_cast<uint64_t>(p.n)};
    const uint64_t strides[3] = {4, row, row * p.h};
    const uint32_t box[3] = {kWsBoxU8 / 4, 8, 1};
    const uint32_t es[3] = {1, 1, 1};
    int rc = din_encode_tmap(&timg, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(p.x), dims, strides, box, es,
                             CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != DIN_OK) return rc;
    timg_b = timg;
  } else {
    const uint64_t dims[3] = {static_cast<uint64_t>(p.w_in), static_cast<uint64_t>(p.h), static_cast<uint64_t>(p.n) * 3};
    const uint64_t strides[3] = {4, static_cast<uint64_t>(p.w_in) * 4, static_cast<uint64_t>(p.w_in) * 4 * p.h};
    const uint32_t es[3] = {1, 1, 1};
    const uint32_t box_a[3] = {kWsBoxA, 8, 3}, box_b[3] = {kWsBoxB, 8, 3};
    int rc = din_encode_tmap(&timg, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(p.x), dims, strides, box_a, es,
                             CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != DIN_OK) return rc;
    rc = din_encode_tmap(&timg_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(p.x), dims, strides, box_b, es,
                         CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != DIN_OK) return rc;
  }
  const int sms = din_num_sms();
  long long grid = sms > 0 ? sms : 148;
  if constexpr (POOL) {
    const long long tasks = static_cast<long long>(p.n) * p.pool_strips;
    if (grid > tasks) grid = tasks;
    const int shared_px = (p.pool_pw - 1) / 64;                       // pooled pixels 64 k, k >= 1
    if (shared_px > 0) {
      const long long rows = static_cast<long long>(p.n) * p.pool_ph, items = rows * shared_px * 8;
      stem_pool_zero_kernel<<<static_cast<unsigned>((items + 255) / 256), 256, 0, st>>>(p.y, rows, p.pool_pw, shared_px);
    }
    DIN_OPT_IN_SMEM((stem_s2d_ws_kernel<U8, true>), kWsPoolSmem);
    stem_s2d_ws_kernel<U8, true><<<static_cast<int>(grid), kWsThreads, kWsPoolSmem, st>>>(timg, timg_b, p);
  } else {
    if (grid > p.num_tiles) grid = p.num_tiles;
    DIN_OPT_IN_SMEM((stem_s2d_ws_kernel<U8, false>), kWsSmem);
    stem_s2d_ws_kernel<U8, false><<<static_cast<int>(grid), kWsThreads, kWsSmem, st>>>(timg, timg_b, p);
  }
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

template <bool U8>
int launch_s2d(StemParams& p, cudaStream_t st) {
  const int sms = din_num_sms();
  int per_sm = static_cast<int>((224 * 1024) / (kS2dSmem + 1024));
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  long long grid = static_cast<long long>(sms > 0 ? sms : 148) * per_sm;
  if (grid > p.num_tiles) grid = p.num_tiles;
  DIN_OPT_IN_SMEM(stem_s2d_kernel<U8>, kS2dSmem);
  stem_s2d_kernel<U8><<<static_cast<int>(grid), kWideThreads, kS2dSmem, st>>>(p);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

void fastdiv(uint32_t d, uint32_t* mul, uint32_t* shr) {
  uint32_t l = 0;
  while ((1u << l) < d) ++l;
  *shr = l;
  *mul = static_cast<uint32_t>(((static_cast<uint64_t>(1) << 32) * ((static_cast<uint64_t>(1) << l) - d)) / d + 1);
}

template <int COUT, int KH, int KW, int STRIDE, bool U8>
int launch_wide(StemParams& p, cudaStream_t st) {
  using Cfg = StemCfg<COUT, KH, KW, STRIDE>;
  constexpr size_t kSmemWide = Cfg::kSmem + 4 * 32 * kEpiPitch;      // eight epilogue scratch areas instead of four
  const int sms = din_num_sms();
  int per_sm = static_cast<int>((224 * 1024) / (kSmemWide + 1024));   // 227 KB per SM, 1 KB reserved per CTA
  const int tmem_limit = 512 / Cfg::kTmemCols;
  if (per_sm > tmem_limit) per_sm = tmem_limit;
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  long long grid = static_cast<long long>(sms > 0 ? sms : 148) * per_sm;
  if (grid > p.num_tiles) grid = p.num_tiles;
  DIN_OPT_IN_SMEM((stem_tc_wide_kernel<COUT, KH, KW, STRIDE, U8>), kSmemWide);
  stem_tc_wide_kernel<COUT, KH, KW, STRIDE, U8><<<static_cast<int>(grid), kWideThreads, kSmemWide, st>>>(p);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

template <int COUT, int KH, int KW, int STRIDE, bool U8>
int launch(StemParams& p, cudaStream_t st) {
  if constexpr (KH * KW >= 25) {                       // large-K stems: the 256-thread kernel (DIN_STEM_WIDE=0: A/B)
    const char* e = std::getenv("DIN_STEM_WIDE");
    if (!(e && e[0] == '0')) return launch_wide<COUT, KH, KW, STRIDE, U8>(p, st);
  }
  using Cfg = StemCfg<COUT, KH, KW, STRIDE>;
  const int sms = din_num_sms();
  int per_sm = static_cast<int>((200 * 1024) / Cfg::kSmem);
  const int tmem_limit = 512 / Cfg::kTmemCols;
  if (per_sm > tmem_limit) per_sm = tmem_limit;
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  long long grid = static_cast<long long>(sms > 0 ? sms : 148) * per_sm;
  if (grid > p.num_tiles) grid = p.num_tiles;
  DIN_OPT_IN_SMEM((stem_tc_kernel<COUT, KH, KW, STRIDE, U8>), Cfg::kSmem);
  stem_tc_kernel<COUT, KH, KW, STRIDE, U8><<<static_cast<int>(grid), kStemThreads, Cfg::kSmem, st>>>(p);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

template <bool U8>
int dispatch(StemParams& p, int c_out, int kh, int kw, int stride, cudaStream_t st) {
  if (c_out == 64 && kh == 3 && kw == 3 && stride == 1) return launch<64, 3, 3, 1, U8>(p, st);   // VGG-16
  if (c_out == 64 && kh == 7 && kw == 7 && stride == 2) {                                         // ResNet-18
    // DIN_STEM_S2D=0: the im2col kernel; =1: the serial implicit-GEMM kernel; default: the pipelined one (A/B knobs)
    const char* e = std::getenv("DIN_STEM_S2D");
    const bool tma_ok = (reinterpret_cast<uintptr_t>(p.x) & 15) == 0 && (U8 ? p.w_in % 16 == 0 : p.w_in % 4 == 0);
    if (p.pad == 3 && !(e && e[0] == '0')) {
      if (tma_ok && !(e && e[0] == '1')) return launch_s2d_ws<U8>(p, st);
      return launch_s2d<U8>(p, st);
    }
    return launch<64, 7, 7, 2, U8>(p, st);
  }
  if (c_out == 32 && kh == 3 && kw == 3 && stride == 2) return launch<32, 3, 3, 2, U8>(p, st);   // Inception-v3
  return DIN_ERR_UNSUPPORTED;
}

// ================================================================================================
// Stem weight / bias gradient on the tensor cores (VGG-16 features.0: 64 x 3 x 3x3, stride 1, pad 1):
//     dW[co][k] += inv_scale * sum_pixels dZ[p][co] * im2col(prep(x))[p][k],  k = (c*3 + ky)*3 + kx;   db = column k = 27
// The same GEMM shape as conv_wgrad_tcgen05.cu (K = the pixel index, both operands MN-major), per 128-pixel strip:
//   A = dZ strip [128 pixels][64 channels] fp16, one TMA box in the 128B-swizzled layout; the upper half of M = 128
//       reads a zero-filled region (LBO points at it), so no M = 64 data path is needed;
//   B = the im2col rows of the strip [128 pixels][32] fp16 (27 taps, a ones column for the bias, 4 zero columns),
//       written by the 128 threads in the canonical no-swizzle MN-major layout from the staged input patch --
//       the forward stem's load_patch / stage_patch;
//   D = [128 x 32] fp32 in TMEM, accumulated over ALL strips of a persistent CTA, flushed once with atomics.
// Replaces the CUDA-core stem_wgrad_kernel (12 % of a training step in the first launch list).
// ================================================================================================
constexpr int kSwgABytes = 128 * 128;            // dZ strip tile

template <int KH, int KW, int STRIDE>
struct SwgCfg {
  using Cfg = StemCfg<64, KH, KW, STRIDE>;
  static constexpr int kTaps = 3 * KH * KW;                          // + 1 ones column = bias gradient
  static constexpr int kN = (kTaps + 1 + 15) / 16 * 16;              // MMA N (multiple of 16): 32 / 160
  static constexpr int kChunks = kN / 8;                             // 16-byte chunks per im2col row
  static constexpr int kLbo = kChunks * 128;                         // bytes between 8-pixel groups
  static constexpr int kBBytes = 128 * kN * 2;
  static constexpr int kTmemCols = kN <= 32 ? 32 : (kN <= 64 ? 64 : (kN <= 128 ? 128 : 256));
  static constexpr size_t kSmem = 1024 + 2 * kSwgABytes + kBBytes + 64 + static_cast<size_t>(Cfg::kRows) * Cfg::kPitch * 4;
};

__device__ __forceinline__ uint64_t desc_mn_sw128_stem(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1u) << 46;
  d |= static_cast<uint64_t>(2u) << 61;          // SWIZZLE_128B
  return d;
}

template <int KH, int KW, int STRIDE, bool U8>
__global__ void __launch_bounds__(kStemThreads)
stem_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dz, const StemParams p, float* __restrict__ dw,
                     float* __restrict__ db, const float* __restrict__ inv_scale, const int c_out) {
  using Cfg = StemCfg<64, KH, KW, STRIDE>;
  using SC = SwgCfg<KH, KW, STRIDE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem(smem_raw, 1024);
  uint8_t* a_s = smem;                                      // [128 pixel rows][128 B], TMA, SWIZZLE_128B
  uint8_t* z_s = a_s + kSwgABytes;                          // zeros: the upper 64 rows of M
  uint8_t* b_s = z_s + kSwgABytes;                          // im2col, no-swizzle MN-major
  uint64_t* tma_bar = reinterpret_cast<uint64_t*>(b_s + SC::kBBytes);
  uint64_t* mma_bar = tma_bar + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(mma_bar + 1);
  float* patch = reinterpret_cast<float*>(tmem_ptr_smem + 4);               // [3*KH][kPitch]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < kSwgABytes / 16; i += kStemThreads) reinterpret_cast<uint4*>(z_s)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    tma_prefetch_desc(&tmap_dz);
    mbar_init(tma_bar, 1);
    mbar_init(mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<SC::kTmemCols>(tmem_ptr_smem);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  // kind::f16, fp32 accumulate, A and B MN-major (bits 15, 16), M = 128, N = kN
  constexpr uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(SC::kN >> 3) << 17) |
                             (static_cast<uint32_t>(128 >> 4) << 24);

  uint32_t phase = 0;
  bool any = false;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const TileCoord cur = decode_tile(p, tile);
    if (tid == 0) {                                          // the previous strip's MMAs have retired (mma_bar wait)
      mbar_arrive_expect_tx(tma_bar, kSwgABytes);
      tma_load_4d(a_s, &tmap_dz, tma_bar, 0, cur.strip * 128, cur.oy, cur.img);
    }
    {
      PatchRegs<64, KH, KW, STRIDE, U8> regs;
      load_patch<64, KH, KW, STRIDE, U8>(p, cur, tid, regs);
      stage_patch<64, KH, KW, STRIDE, U8>(p, tid, patch, regs);
    }
    __syncthreads();
    {
      // im2col row of pixel `tid`: kN fp16 = kChunks 16-byte chunks; chunk j of pixel p lives at
      // (p / 8) * kLbo + j * 128 + (p % 8) * 16   (8-pixel groups kLbo apart = LBO, 8-column chunks 128 B apart = SBO)
      const float* prow = patch + tid * STRIDE;
      uint8_t* dst = b_s + (tid >> 3) * SC::kLbo + (tid & 7) * 16;
#pragma unroll
      for (int j = 0; j < SC::kChunks; ++j) {
        __align__(16) __half2 hv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float f[2];
#pragma unroll
          for (int z = 0; z < 2; ++z) {
            const int k = j * 8 + 2 * e + z;                 // k = (c*KH + ky)*KW + kx (OIHW flatten); kTaps = bias
            f[z] = (k < SC::kTaps) ? prow[(k / KW) * Cfg::kPitch + (k % KW)] : (k == SC::kTaps ? 1.0f : 0.0f);
          }
          hv[e] = __floats2half2_rn(f[0], f[1]);
        }
        *reinterpret_cast<uint4*>(dst + j * 128) = *reinterpret_cast<const uint4*>(hv);
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      mbar_wait(tma_bar, phase);
      tc_fence_after_sync();
      const uint32_t a_addr = smem_u32(a_s), b_addr = smem_u32(b_s);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {                    // 16 pixels per MMA
          const uint64_t ad = desc_mn_sw128_stem(a_addr + ks * 2048, kSwgABytes, 1024);
          const uint64_t bd = desc_noswz(b_addr + ks * 2 * SC::kLbo, SC::kLbo, 128);
          umma_f16_ss(tmem_base, ad, bd, idesc, (any || ks > 0) ? 1u : 0u);
        }
        umma_commit(mma_bar);
      }
      __syncwarp();
    }
    any = true;
    mbar_wait(mma_bar, phase);                               // a_s, b_s and the patch may be rebuilt
    phase ^= 1u;
    tc_fence_after_sync();
  }

  // ---- flush: TMEM lanes 0..63 = output channels, columns = taps (kTaps = bias)
  if (any && warp < 2 && warp * 32 < c_out) {               // c_out = 32 (Inception-v3): rows 32.. were zero-filled by TMA
    const float scl = inv_scale ? __ldg(inv_scale) : 1.0f;
    const int co = warp * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < SC::kN; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int k = c0 + j;
        if (k < SC::kTaps) atomicAdd(dw + co * SC::kTaps + k, __uint_as_float(v[j]) * scl);
        else if (k == SC::kTaps && db != nullptr) atomicAdd(db + co, __uint_as_float(v[j]) * scl);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc<SC::kTmemCols>(tmem_base);
  }
}

template <int KH, int KW, int STRIDE, bool U8>
int launch_stem_wgrad(const CUtensorMap& tdz, StemParams& p, float* dw, float* dbias, const float* inv_scale, int c_out,
                      cudaStream_t st) {
  using SC = SwgCfg<KH, KW, STRIDE>;
  const int sms = din_num_sms();
  int per_sm = static_cast<int>((224 * 1024) / (SC::kSmem + 1024));
  if (per_sm > 512 / SC::kTmemCols) per_sm = 512 / SC::kTmemCols;
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  long long grid = static_cast<long long>(sms > 0 ? sms : 148) * per_sm;
  if (grid > p.num_tiles) grid = p.num_tiles;
  DIN_OPT_IN_SMEM((stem_wgrad_tc_kernel<KH, KW, STRIDE, U8>), SC::kSmem);
  stem_wgrad_tc_kernel<KH, KW, STRIDE, U8><<<static_cast<int>(grid), kStemThreads, SC::kSmem, st>>>(tdz, p, dw, dbias,
                                                                                                  inv_scale, c_out);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

}  // namespace

// x_is_u8 == 0: x is fp32 NCHW [n,3,h,w];  != 0: x is uint8 NHWC [n,h,w,3] (the decoded frame as the loader
// holds it before `img.transpose(2,0,1)` / `.float()`, volleyball.py:239-243,270).
// Returns DIN_ERR_UNSUPPORTED for filter geometries without a tensor-core instantiation (the fp32 caller then
// uses the generic CUDA-core kernel in stem_pool.cu).
int din_stem_tc_launch(const void* x, int x_is_u8, const float* w, const float* bias, void* y, int n, int h,
                       int w_in, int c_out, int kh, int kw, int stride, int pad, int relu, int prep,
                       cudaStream_t st) {
  StemParams p{};
  p.x = x; p.w = w; p.bias = bias; p.y = static_cast<__half*>(y);
  p.n = n; p.h = h; p.w_in = w_in;
  p.oh = (h + 2 * pad - kh) / stride + 1;
  p.ow = (w_in + 2 * pad - kw) / stride + 1;
  p.pad = pad; p.relu = relu; p.prep = prep;
  p.strips_per_row = (p.ow + 127) / 128;
  const long long tiles = static_cast<long long>(n) * p.oh * p.strips_per_row;
  if (tiles >= INT32_MAX) return din_set_error(DIN_ERR_INVALID_ARG, "din_stem_conv: too many tiles");
  p.num_tiles = static_cast<int>(tiles);
  fastdiv(static_cast<uint32_t>(p.strips_per_row), &p.fd_mul, &p.fd_shr);
  fastdiv(static_cast<uint32_t>(p.oh), &p.fo_mul, &p.fo_shr);
  return x_is_u8 ? dispatch<true>(p, c_out, kh, kw, stride, st) : dispatch<false>(p, c_out, kh, kw, stride, st);
}

// ResNet-18's conv1 + bn1 (folded) + relu + maxpool (backbone.py:115-132) in one launch: stem_s2d_ws_kernel<U8, POOL = true>.
// y: [n, ph, pw, 64] with ph = (oh - 1) / 2 + 1.  DIN_ERR_UNSUPPORTED when the image rows do not meet the TMA alignment (the
// caller then runs the stem and the pool separately).
int din_stem_pool_tc_launch(const void* x, int x_is_u8, const float* w, const float* bias, void* y, int n, int h, int w_in,
                            int prep, cudaStream_t st) {
  StemParams p{};
  p.x = x; p.w = w; p.bias = bias; p.y = static_cast<__half*>(y);
  p.n = n; p.h = h; p.w_in = w_in;
  p.oh = (h + 6 - 7) / 2 + 1;
  p.ow = (w_in + 6 - 7) / 2 + 1;
  p.pad = 3; p.relu = 1; p.prep = prep;
  p.pool_ph = (p.oh - 1) / 2 + 1;
  p.pool_pw = (p.ow - 1) / 2 + 1;
  p.pool_strips = (p.ow + 127) / 128;
  p.strips_per_row = p.pool_strips;
  const long long tiles = static_cast<long long>(n) * p.oh * p.pool_strips;
  if (tiles >= INT32_MAX) return din_set_error(DIN_ERR_INVALID_ARG, "din_stem_conv_pool: too many tiles");
  p.num_tiles = static_cast<int>(tiles);
  const bool tma_ok = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (x_is_u8 ? w_in % 16 == 0 : w_in % 4 == 0);
  if (!tma_ok) return DIN_ERR_UNSUPPORTED;
  return x_is_u8 ? launch_s2d_ws<true, true>(p, st) : launch_s2d_ws<false, true>(p, st);
}

// Tensor-core stem weight gradient (see stem_wgrad_tc_kernel); arguments validated by din_stem_wgrad.
// Geometries: VGG-16 (3x3, stride 1, pad 1) and ResNet-18 (7x7, stride 2, pad 3) with 64 output channels, Inception-v3
// (3x3, stride 2, pad 0) with 32.
int din_stem_wgrad_tc_launch(const void* x, int x_is_u8, const void* dz, float* dw, float* dbias, const float* inv_scale,
                             int n, int h, int w_in, int c_out, int kh, int stride, int pad, int prep, cudaStream_t st) {
  StemParams p{};
  p.x = x; p.n = n; p.h = h; p.w_in = w_in; p.pad = pad; p.prep = prep;
  p.oh = (h + 2 * pad - kh) / stride + 1;
  p.ow = (w_in + 2 * pad - kh) / stride + 1;
  p.strips_per_row = (p.ow + 127) / 128;
  const long long tiles = static_cast<long long>(n) * p.oh * p.strips_per_row;
  if (tiles >= INT32_MAX) return din_set_error(DIN_ERR_INVALID_ARG, "din_stem_wgrad: too many tiles");
  p.num_tiles = static_cast<int>(tiles);
  fastdiv(static_cast<uint32_t>(p.strips_per_row), &p.fd_mul, &p.fd_shr);
  fastdiv(static_cast<uint32_t>(p.oh), &p.fo_mul, &p.fo_shr);
  CUtensorMap tdz;
  {
    // c_out = 32: the 64-channel box is half out of bounds -> zero-filled rows 32..63 of M
    const uint64_t cs = 2ull * c_out;
    const uint64_t dims[4] = {static_cast<uint64_t>(c_out), static_cast<uint64_t>(p.ow), static_cast<uint64_t>(p.oh),
                              static_cast<uint64_t>(n)};
    const uint64_t strides[4] = {2, cs, cs * p.ow, cs * p.ow * p.oh};
    const uint32_t box[4] = {64, 128, 1, 1};
    const uint32_t es[4] = {1, 1, 1, 1};
    int rc = din_encode_tmap(&tdz, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(dz), dims, strides, box, es,
                             CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != DIN_OK) return rc;
  }
  if (kh == 3 && stride == 1)
    return x_is_u8 ? launch_stem_wgrad<3, 3, 1, true>(tdz, p, dw, dbias, inv_scale, c_out, st)
                   : launch_stem_wgrad<3, 3, 1, false>(tdz, p, dw, dbias, inv_scale, c_out, st);
  if (kh == 7 && stride == 2)
    return x_is_u8 ? launch_stem_wgrad<7, 7, 2, true>(tdz, p, dw, dbias, inv_scale, c_out, st)
                   : launch_stem_wgrad<7, 7, 2, false>(tdz, p, dw, dbias, inv_scale, c_out, st);
  if (kh == 3 && stride == 2)                                  // Inception-v3 Conv2d_1a_3x3 (32 channels, pad 0)
    return x_is_u8 ? launch_stem_wgrad<3, 3, 2, true>(tdz, p, dw, dbias, inv_scale, c_out, st)
                   : launch_stem_wgrad<3, 3, 2, false>(tdz, p, dw, dbias, inv_scale, c_out, st);
  return din_set_error(DIN_ERR_UNSUPPORTED, "din_stem_wgrad: unsupported stem geometry %dx%d stride %d", kh, kh, stride);
}
