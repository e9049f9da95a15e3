// conv_tcgen05.cu — implicit-GEMM convolution / dense GEMM on the 5th-gen tensor cores.
//
// Replaces every Conv2d(+BN)(+ReLU) of the truncated torchvision backbones the reference wraps
// (backbone/backbone.py:10-132; cuDNN implicit GEMM there) and nn.Linear fc_emb_1
// (infer_model.py:50,184; cuBLAS sgemm there).
//
// Formulation.  Output tile = 128 output pixels (a th x tw patch of one image, th*tw = 128) x BN
// output channels.  K runs over (filter tap, 64-channel block).  For one K step
//   A = the th x tw x 64ch input patch shifted by the tap offset.  Activations are NHWC fp16, so this
//       is ONE 4-D TMA box {64ch, tw, th, 1}; out-of-image elements are zero-filled by the TMA unit,
//       which *is* the convolution's zero padding; stride-2 convs use the tensor map's element
//       strides.  With 64 fp16 = 128 B per pixel and SWIZZLE_128B the box lands exactly in the
//       K-major SW128 layout tcgen05.mma consumes: no im2col buffer ever exists.
//   B = BN rows x 64 columns of the packed weight [c_out][tap][c_in] (K-major): one 2-D TMA box.
// A persistent, warp-specialised CTA per SM: warp 0 = TMA producer, warp 1 = MMA issuer (one lane),
// warps 2..5 = epilogue (TMEM -> registers -> bias / residual / ReLU -> global).  The fp32
// accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the main loop of
// tile i+1.
#include "din_common.cuh"

namespace {

using namespace din;

struct ConvKParams {
  int oh, ow;
  int c_out, y_c_stride;
  int tiles_x, tiles_per_img;
  int n_tiles_n;
  int num_tiles;
  int th, tw, tw_log2;
  int kh, kw, stride, pad_h, pad_w;
  int n_cblk;  // c_in / 64
  int c_in;
  int relu, out_f32;
  const float* bias;
  const __half* residual;
  void* y;
};

constexpr int kBM = 128;
constexpr int kBK = 64;                    // fp16 elements per K step = one 128-byte swizzle row
constexpr int kABytes = kBM * kBK * 2;     // 16 KB
constexpr int kNumThreads = 192;

template <int BN>
struct ConvCfg {
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (192 * 1024) / kStageBytes;  // 4 / 6 / 8 stages for BN = 256 / 128 / 64
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const ConvKParams p) {
  using Cfg = ConvCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int k_iters = p.kh * p.kw * p.n_cblk;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one lane)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles_n;
        const int mt = tile / p.n_tiles_n;
        const int img = mt / p.tiles_per_img;
        const int r = mt - img * p.tiles_per_img;
        const int tyi = r / p.tiles_x;
        const int txi = r - tyi * p.tiles_x;
        const int ix0 = txi * p.tw * p.stride - p.pad_w;
        const int iy0 = tyi * p.th * p.stride - p.pad_h;
        const int n0 = nt * BN;
        for (int ky = 0; ky < p.kh; ++ky) {
          for (int kx = 0; kx < p.kw; ++kx) {
            const int kbase = (ky * p.kw + kx) * p.c_in;
            for (int cb = 0; cb < p.n_cblk; ++cb) {
              mbar_wait(&empty_bar[stage], phase ^ 1u);
              uint8_t* sa = smem + stage * Cfg::kStageBytes;
              uint8_t* sb = sa + kABytes;
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
              tma_load_4d(sa, &tmap_a, &full_bar[stage], cb * kBK, ix0 + kx, iy0 + ky, img);
              tma_load_2d(sb, &tmap_b, &full_bar[stage], kbase + cb * kBK, n0);
              if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (lane 0 issues)
    constexpr uint32_t idesc = umma_idesc_f16_f32(kBM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int kit = 0; kit < k_iters; ++kit) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // advance 16 fp16 = 32 bytes along K inside the 128-byte swizzle row: +2 in 16-byte units
            umma_f16_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kit | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                       // frees the smem slot when the MMAs retire
          if (kit == k_iters - 1) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps, 128 lanes)
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int m = q * 32 + lane;
    const int ty = m >> p.tw_log2;
    const int tx = m & (p.tw - 1);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int nt = tile % p.n_tiles_n;
      const int mt = tile / p.n_tiles_n;
      const int img = mt / p.tiles_per_img;
      const int r = mt - img * p.tiles_per_img;
      const int tyi = r / p.tiles_x;
      const int txi = r - tyi * p.tiles_x;
      const int oy = tyi * p.th + ty;
      const int ox = txi * p.tw + tx;
      const bool valid = (oy < p.oh) && (ox < p.ow);
      const int n0 = nt * BN;
      const size_t pix = (static_cast<size_t>(img) * p.oh + oy) * p.ow + ox;

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        tmem_ld_wait();
        const int col0 = n0 + c0;
        if (valid && col0 < p.c_out) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (col0 + j < p.c_out) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
              }
            }
          }
          const size_t off = pix * p.y_c_stride + col0;
          if (p.residual != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (col0 + j < p.c_out) {
                const uint4 rr = __ldg(reinterpret_cast<const uint4*>(p.residual + off + j));
                const __half2* h2 = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 t2 = __half22float2(h2[e]);
                  f[j + 2 * e] += t2.x;
                  f[j + 2 * e + 1] += t2.y;
                }
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
          }
          if (p.out_f32) {
            float* yp = reinterpret_cast<float*>(p.y) + off;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (col0 + j < p.c_out) {
                *reinterpret_cast<float4*>(yp + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
              }
            }
          } else {
            __half* yp = reinterpret_cast<__half*>(p.y) + off;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (col0 + j < p.c_out) {
                uint4 o;
                __half2* h2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(f[j + 2 * e], f[j + 2 * e + 1]);
                *reinterpret_cast<uint4*>(yp + j) = o;
              }
            }
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int BN>
int launch_conv(const CUtensorMap& ta, const CUtensorMap& tb, const ConvKParams& p, int grid, cudaStream_t st) {
  using Cfg = ConvCfg<BN>;
  static thread_local int attr_dev = -1;
  int dev = 0;
  DIN_CHECK_CUDA(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    DIN_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytes));
    attr_dev = dev;
  }
  conv_igemm_kernel<BN><<<grid, kNumThreads, Cfg::kSmemBytes, st>>>(ta, tb, p);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

__global__ void pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                   __half* __restrict__ out, int c_out, int c_in, int c_in_p, int taps) {
  const size_t total = static_cast<size_t>(c_out) * taps * c_in_p;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % c_in_p);
    const size_t t2 = i / c_in_p;
    const int tap = static_cast<int>(t2 % taps);
    const int o = static_cast<int>(t2 / taps);
    float v = 0.0f;
    if (c < c_in) {
      v = w[(static_cast<size_t>(o) * c_in + c) * taps + tap];
      if (scale != nullptr) v *= scale[o];
    }
    out[i] = __float2half_rn(v);
  }
}

}  // namespace

extern "C" int din_pack_conv_weight_f16(const float* w_oihw, const float* scale, void* w_packed, int c_out,
                                        int c_in, int c_in_padded, int kh, int kw, void* stream) {
  DIN_CHECK_ARG(w_oihw && w_packed, "din_pack_conv_weight_f16: null pointer");
  DIN_CHECK_ARG(c_out > 0 && c_in > 0 && c_in_padded >= c_in && kh > 0 && kw > 0,
                "din_pack_conv_weight_f16: bad shape c_out=%d c_in=%d c_in_padded=%d k=%dx%d", c_out, c_in,
                c_in_padded, kh, kw);
  const size_t total = static_cast<size_t>(c_out) * kh * kw * c_in_padded;
  const int block = 256;
  const int grid = static_cast<int>(std::min<size_t>((total + block - 1) / block, 148 * 8));
  pack_weight_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      w_oihw, scale, static_cast<__half*>(w_packed), c_out, c_in, c_in_padded, kh * kw);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

extern "C" int din_conv2d_nhwc_f16(const DinConvDesc* d, const void* x, const void* w_packed, const float* bias,
                                   const void* residual, void* y, void* stream) {
  DIN_CHECK_ARG(d && x && w_packed && y, "din_conv2d_nhwc_f16: null pointer");
  DIN_CHECK_ARG(d->n > 0 && d->h > 0 && d->w > 0, "din_conv2d_nhwc_f16: bad extent n=%d h=%d w=%d", d->n, d->h,
                d->w);
  DIN_CHECK_ARG(d->c_in > 0 && d->c_in % kBK == 0, "din_conv2d_nhwc_f16: c_in=%d must be a multiple of 64",
                d->c_in);
  DIN_CHECK_ARG(d->x_c_stride >= d->c_in && d->x_c_stride % 8 == 0,
                "din_conv2d_nhwc_f16: x_c_stride=%d must be >= c_in and a multiple of 8", d->x_c_stride);
  DIN_CHECK_ARG(d->c_out > 0 && d->c_out % 8 == 0, "din_conv2d_nhwc_f16: c_out=%d must be a multiple of 8",
                d->c_out);
  DIN_CHECK_ARG(d->y_c_stride >= d->c_out && d->y_c_stride % 8 == 0,
                "din_conv2d_nhwc_f16: y_c_stride=%d must be >= c_out and a multiple of 8", d->y_c_stride);
  DIN_CHECK_ARG(d->kh >= 1 && d->kw >= 1 && d->kh * d->kw <= 49, "din_conv2d_nhwc_f16: bad filter %dx%d", d->kh,
                d->kw);
  DIN_CHECK_ARG(d->stride == 1 || d->stride == 2, "din_conv2d_nhwc_f16: stride=%d unsupported", d->stride);
  DIN_CHECK_ARG(d->pad_h >= 0 && d->pad_w >= 0, "din_conv2d_nhwc_f16: negative padding");
  DIN_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(residual) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
                "din_conv2d_nhwc_f16: pointers must be 16-byte aligned");
  const int oh = (d->h + 2 * d->pad_h - d->kh) / d->stride + 1;
  const int ow = (d->w + 2 * d->pad_w - d->kw) / d->stride + 1;
  DIN_CHECK_ARG(oh > 0 && ow > 0, "din_conv2d_nhwc_f16: empty output %dx%d", oh, ow);

  // output tile geometry: th x tw = 128 pixels, fewest tiles wins, wider rows break ties
  int best_tw = 128, best_tiles = INT32_MAX;
  for (int tw = 128; tw >= 8; tw >>= 1) {
    const int th = 128 / tw;
    const int tiles = ((ow + tw - 1) / tw) * ((oh + th - 1) / th);
    if (tiles < best_tiles) { best_tiles = tiles; best_tw = tw; }
  }
  ConvKParams p{};
  p.oh = oh; p.ow = ow;
  p.c_out = d->c_out; p.y_c_stride = d->y_c_stride;
  p.tw = best_tw; p.th = 128 / best_tw;
  p.tw_log2 = 0;
  while ((1 << p.tw_log2) < p.tw) ++p.tw_log2;
  p.tiles_x = (ow + p.tw - 1) / p.tw;
  p.tiles_per_img = best_tiles;
  const int bn = d->c_out > 128 ? 256 : (d->c_out > 64 ? 128 : 64);
  p.n_tiles_n = (d->c_out + bn - 1) / bn;
  const long long total_tiles = static_cast<long long>(d->n) * p.tiles_per_img * p.n_tiles_n;
  DIN_CHECK_ARG(total_tiles < INT32_MAX, "din_conv2d_nhwc_f16: too many tiles");
  p.num_tiles = static_cast<int>(total_tiles);
  p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad_h = d->pad_h; p.pad_w = d->pad_w;
  p.c_in = d->c_in; p.n_cblk = d->c_in / kBK;
  p.relu = d->relu; p.out_f32 = d->out_f32;
  p.bias = bias; p.residual = static_cast<const __half*>(residual); p.y = y;

  CUtensorMap ta, tb;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(d->c_in), static_cast<uint64_t>(d->w),
                              static_cast<uint64_t>(d->h), static_cast<uint64_t>(d->n)};
    const uint64_t cs = static_cast<uint64_t>(d->x_c_stride) * 2;
    const uint64_t strides[4] = {2, cs, cs * d->w, cs * d->w * d->h};
    const uint32_t box[4] = {static_cast<uint32_t>(kBK), static_cast<uint32_t>((p.tw - 1) * d->stride + 1),
                             static_cast<uint32_t>((p.th - 1) * d->stride + 1), 1};
    const uint32_t es[4] = {1, static_cast<uint32_t>(d->stride), static_cast<uint32_t>(d->stride), 1};
    int rc = din_encode_tmap(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, box,
                             es, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != DIN_OK) return rc;
  }
  {
    const uint64_t ktot = static_cast<uint64_t>(d->kh) * d->kw * d->c_in;
    const uint64_t dims[2] = {ktot, static_cast<uint64_t>(d->c_out)};
    const uint64_t strides[2] = {2, ktot * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(kBK), static_cast<uint32_t>(bn)};
    const uint32_t es[2] = {1, 1};
    int rc = din_encode_tmap(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(w_packed), dims, strides,
                             box, es, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != DIN_OK) return rc;
  }
  const int sms = din_num_sms();
  DIN_CHECK_ARG(sms > 0, "din_conv2d_nhwc_f16: no CUDA device");
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (bn) {
    case 256: return launch_conv<256>(ta, tb, p, grid, st);
    case 128: return launch_conv<128>(ta, tb, p, grid, st);
    default: return launch_conv<64>(ta, tb, p, grid, st);
  }
}
