// conv_wgrad_tcgen05.cu — weight gradient of the backbones' stride-1 convolutions (1x1, 3x3, 5x5, 1x7, 7x1; any padding)
// on the tcgen05 tensor cores.
//
// What autograd computes for nn.Conv2d.weight in `total_loss.backward()` (train_net_dynamic.py:220-224) when the
// backbone is trained (scripts/train_volleyball_stage2_dynamic.py:12 `cfg.train_backbone = True`; VGG-16
// features.*, backbone.py:88-99):
//     dW[co][ky][kx][ci] = sum over (img, y, x) of  dZ[img, y, x, co] * X[img, y + ky - pad, x + kx - pad, ci]
// i.e. a GEMM whose K dimension is the PIXEL index: D[co, (kx, ci)] += dZ^T[co, pixel] * X_shifted[pixel, ci].
//
// Both operands are NHWC fp16 activations, so the pixel index is the slow dimension of both tiles: they are
// *MN-major* UMMA operands (instruction-descriptor bits 15/16), and the very same TMA boxes the forward kernel
// loads (64 channels = one 128-byte swizzle row per pixel) are consumed without any transpose:
//   * A = dZ tile: 16 x 8 output pixels x 128 output channels = two boxes {64 c, 8 w, 16 h}: K = 128 pixel rows,
//     8-row groups 1024 B apart (SBO), the two 64-channel halves 16 KB apart (LBO);
//   * B = the input halo rows of ONE filter row ky: {64 c, 8 + KW - 1 w, 16 h} per 64 input channels; tap kx is the
//     same shared memory seen through a descriptor shifted by kx rows, with SBO = the halo pitch (8 + KW - 1 rows):
//     the hardware applies the 128B swizzle on absolute address bits (probed for the forward kernel);
//   * accumulators: KW taps x NCI input channels fp32 TMEM columns (3 x 128, 5 x 64, 7 x 64, 1 x 128), accumulated over
//     ALL pixel tiles the CTA owns (no per-tile epilogue); TMA zero-fill is the padding, nulls ragged tiles and fills a
//     partial last channel block (Inception-v3: c_in = 32, 48, 80, 96, 160, 288; the tensor map carries the real c_in).
// Work unit = (128 output channels, NCI input channels, filter row ky); the pixel tiles of a unit are split over
// several CTAs so that the grid fills the 148 SMs about twice; each CTA finishes with fp32 vector atomics into dW
// (scaled by 1/loss-scale).  One producer warp, one MMA warp (a single elected lane issues), four epilogue warps.
#include <cstdlib>

#include "din_common.cuh"

namespace {

using namespace din;

struct WgradParams {
  int c_in, c_out, kh, kw;
  int pad_h, pad_w;
  int tiles_x, tiles_per_img, num_tiles;   // 16 x 8 tiles of the dZ grid, all images
  int n_ci_blk;                            // c_in / NCI
  int tiles_per_split, splits;
  float* dw;                               // [c_out][kh][kw][c_in] fp32
  const float* inv_scale;                  // device scalar (1 / loss scale) or nullptr
};

constexpr int kWgThreads = 192;
constexpr int kWgABytes = 128 * 128;       // one 64-channel half of the dZ tile: 128 pixel rows x 128 B
constexpr int wg_b_bytes(int kw) { return 16 * (8 + kw - 1) * 128; }   // one 64-channel block of the halo rows of a filter row
constexpr int kWgMaxStages = 6;

// shared-memory descriptor for an MN-major SWIZZLE_128B operand (canonical layout ((8,8,m),(8,k)) in 16-byte
// units: 64 contiguous MN elements per 128-byte row, 8 K rows per group): LBO = bytes between 64-element MN
// chunks, SBO = bytes between 8-row K groups.
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1u) << 46;
  d |= static_cast<uint64_t>(2u) << 61;
  return d;
}

// kind::f16 instruction descriptor: fp16 A/B, fp32 D, BOTH operands MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t idesc_f16_f32_mn(int m, int n) {
  return (1u << 4) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// PAIR64 (c_out <= 64, KW = 3): the upper 64 rows of M, otherwise zero-filled, carry a second copy of the dZ tile shifted
// by one pixel to the left: row 64 + co of the MMA for tap kx then accumulates sum_p dZ[p - 1, co] * X[p + kx] -- tap kx + 1
// over the pixel set shifted by one.  Two MMA groups (kx = 0, 2) instead of three produce all three taps (and a discarded
// "tap 3"); the pixel tiles run over ceil((ow + 1) / 8) columns so that the shifted windows cover every pixel once.
template <int NCI, int KW, bool PAIR64 = false>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dz, const __grid_constant__ CUtensorMap tmap_x,
                  const WgradParams p, const int n_stages) {
  constexpr int kWgBBytes = wg_b_bytes(KW);
  constexpr int kPitchBytes = (8 + KW - 1) * 128;         // one halo row
  constexpr int kStageBytes = 2 * kWgABytes + (NCI / 64) * kWgBBytes;
  constexpr int kTmemCols = (KW * NCI <= 64) ? 64 : (KW * NCI <= 128) ? 128 : (KW * NCI <= 256) ? 256 : 512;
  static_assert(KW * NCI <= 512, "accumulators exceed TMEM");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem(smem_raw, 1024);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + n_stages * kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = full + kWgMaxStages;
  uint64_t* acc_full = empty + kWgMaxStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- which unit / which slice of the pixel tiles
  const int unit = blockIdx.x / p.splits;
  const int split = blockIdx.x - unit * p.splits;
  const int ky = unit % p.kh;
  const int blk = unit / p.kh;
  const int ci0 = (blk % p.n_ci_blk) * NCI;
  const int co0 = (blk / p.n_ci_blk) * 128;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(p.num_tiles, t_begin + p.tiles_per_split);
  if (t_begin >= t_end) return;                       // uniform over the CTA

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_dz);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < n_stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_ptr_smem);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);

  if (warp == 0) {
    // ------------------------------------------------------------------ producer (one lane)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int img = tile / p.tiles_per_img;
        const int r = tile - img * p.tiles_per_img;
        const int tyi = r / p.tiles_x, txi = r - tyi * p.tiles_x;
        const int x0 = txi * 8, y0 = tyi * 16;
        mbar_wait(&empty[stage], phase ^ 1u);
        uint8_t* sa = smem + stage * kStageBytes;
        uint8_t* sb = sa + 2 * kWgABytes;
        mbar_arrive_expect_tx(&full[stage], static_cast<uint32_t>(kStageBytes));
        tma_load_4d(sa, &tmap_dz, &full[stage], co0, x0, y0, img);
        if constexpr (PAIR64) tma_load_4d(sa + kWgABytes, &tmap_dz, &full[stage], co0, x0 - 1, y0, img);
        else tma_load_4d(sa + kWgABytes, &tmap_dz, &full[stage], co0 + 64, x0, y0, img);   // beyond c_out: zero fill
#pragma unroll
        for (int c = 0; c < NCI / 64; ++c)
          tma_load_4d(sb + c * kWgBBytes, &tmap_x, &full[stage], ci0 + c * 64, x0 - p.pad_w, y0 + ky - p.pad_h, img);
        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, one lane issues)
    const bool leader = elect_one();
    constexpr uint32_t idesc = idesc_f16_f32_mn(128, NCI);
    const uint32_t smem0 = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accum = 0;
    for (int tile = t_begin; tile < t_end; ++tile) {
      mbar_wait(&full[stage], phase);
      tc_fence_after_sync();
      const uint32_t sa = smem0 + stage * kStageBytes;
      const uint32_t sb = sa + 2 * kWgABytes;
      if (leader) {
#pragma unroll
        for (int kx = 0; kx < KW; kx += (PAIR64 ? 2 : 1)) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {            // 16 pixels (two 8-pixel tile rows) per MMA
            const uint64_t ad = desc_mn_sw128(sa + ks * 2048, kWgABytes, 1024);
            const uint64_t bd = desc_mn_sw128(sb + kx * 128 + ks * 2 * kPitchBytes, kWgBBytes, kPitchBytes);
            umma_f16_ss(tmem_base + kx * NCI, ad, bd, idesc, ks == 0 ? accum : 1u);
          }
        }
        umma_commit(&empty[stage]);
      }
      accum = 1;
      if (++stage == n_stages) { stage = 0; phase ^= 1u; }
    }
    if (leader) umma_commit(acc_full);
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> fp32 atomics into dW
    const int q = warp & 3;                            // TMEM lane quadrant this warp may read
    mbar_wait(acc_full, 0);
    tc_fence_after_sync();
    const float scl = p.inv_scale ? __ldg(p.inv_scale) : 1.0f;
    // PAIR64: TMEM lanes 64..127 (quadrants 2, 3) hold tap kx + 1 of channel lane - 64
    const int co = PAIR64 ? co0 + (q & 1) * 32 + lane : co0 + q * 32 + lane;
    const int tap_shift = PAIR64 ? (q >> 1) : 0;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
    for (int kx = 0; kx < KW; kx += (PAIR64 ? 2 : 1)) {
#pragma unroll 1
      for (int c0 = 0; c0 < NCI; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + kx * NCI + c0, v);
        tmem_ld_wait();
        if (co < p.c_out && kx + tap_shift < KW) {
          float* dst = p.dw + ((static_cast<size_t>(co) * p.kh + ky) * p.kw + kx + tap_shift) * p.c_in + ci0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (ci0 + c0 + j < p.c_in)                  // partial last channel block (c_in is a multiple of 8)
              atomicAdd(reinterpret_cast<float4*>(dst + j),
                      make_float4(__uint_as_float(v[j]) * scl, __uint_as_float(v[j + 1]) * scl,
                                  __uint_as_float(v[j + 2]) * scl, __uint_as_float(v[j + 3]) * scl));
          }
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2, M = 256) for c_out > 128 and c_in % 128 == 0.  ncu on the one-CTA kernel: 78 % tensor
// pipe, because a 128 x 128 x 16 MMA with two MN-major operands reads 8 KB of shared memory in its 64 cycles -- the whole
// 128 B/clk of the SM's shared-memory pipe -- while TMA writes the next stage into the same memory.  Here the two CTAs of
// a cluster take the two 128-channel halves of a 256-channel dZ block (M = 256) against the SAME input halo: each CTA
// stages only ITS 64 of the 128 input channels (B is split along N across the pair), so per MMA a CTA reads 4 + 2 KB, and
// per pixel tile it loads 52 KB instead of 72 KB.  Both CTAs run the producer for their own boxes (completion bytes land
// on the leader's barrier), the leader's MMA warp issues, tcgen05.commit multicasts to both CTAs' barriers, every CTA
// drains its own 128 TMEM lanes (= its output channels, all NCI x KW columns).
// ------------------------------------------------------------------------------------------------------------------
template <int KW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWgThreads, 1)
conv_wgrad_2cta_kernel(const __grid_constant__ CUtensorMap tmap_dz, const __grid_constant__ CUtensorMap tmap_x,
                       const WgradParams p, const int n_stages) {
  constexpr int NCI = 128;
  constexpr int kWgBBytes = wg_b_bytes(KW);               // this CTA's 64 input channels of the halo rows
  constexpr int kPitchBytes = (8 + KW - 1) * 128;
  constexpr int kStageBytes = 2 * kWgABytes + kWgBBytes;
  constexpr int kTmemCols = (KW * NCI <= 128) ? 128 : (KW * NCI <= 256) ? 256 : 512;
  static_assert(KW * NCI <= 512, "accumulators exceed TMEM");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem(smem_raw, 1024);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + n_stages * kStageBytes);
  uint64_t* full = bars;                                  // used in the leader CTA only
  uint64_t* empty = full + kWgMaxStages;
  uint64_t* acc_full = empty + kWgMaxStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());
  const int pair = blockIdx.x >> 1;
  const int unit = pair / p.splits;
  const int split = pair - unit * p.splits;
  const int ky = unit % p.kh;
  const int blk = unit / p.kh;
  const int ci0 = (blk % p.n_ci_blk) * NCI;
  const int co0 = (blk / p.n_ci_blk) * 256 + rank * 128;  // this CTA's 128 output channels
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(p.num_tiles, t_begin + p.tiles_per_split);
  if (t_begin >= t_end) return;                           // uniform over the pair

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_dz);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < n_stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<kTmemCols>(tmem_ptr_smem);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                                     // the peer's barriers exist before anything signals them
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        const int img = tile / p.tiles_per_img;
        const int r = tile - img * p.tiles_per_img;
        const int tyi = r / p.tiles_x, txi = r - tyi * p.tiles_x;
        const int x0 = txi * 8, y0 = tyi * 16;
        mbar_wait(&empty[stage], phase ^ 1u);             // own barrier: the leader's commit is multicast to both CTAs
        uint8_t* sa = smem + stage * kStageBytes;
        uint8_t* sb = sa + 2 * kWgABytes;
        if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2u * static_cast<uint32_t>(kStageBytes));
        tma_load_4d_2cta(sa, &tmap_dz, &full[stage], co0, x0, y0, img);
        tma_load_4d_2cta(sa + kWgABytes, &tmap_dz, &full[stage], co0 + 64, x0, y0, img);
        tma_load_4d_2cta(sb, &tmap_x, &full[stage], ci0 + rank * 64, x0 - p.pad_w, y0 + ky - p.pad_h, img);
        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const bool leader = elect_one();
      constexpr uint32_t idesc = idesc_f16_f32_mn(256, NCI);
      const uint32_t smem0 = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accum = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        mbar_wait(&full[stage], phase);
        tc_fence_after_sync();
        const uint32_t sa = smem0 + stage * kStageBytes;
        const uint32_t sb = sa + 2 * kWgABytes;
        if (leader) {
#pragma unroll
          for (int kx = 0; kx < KW; ++kx) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t ad = desc_mn_sw128(sa + ks * 2048, kWgABytes, 1024);
              const uint64_t bd = desc_mn_sw128(sb + kx * 128 + ks * 2 * kPitchBytes, kWgBBytes, kPitchBytes);
              umma_f16_ss_2cta(tmem_base + kx * NCI, ad, bd, idesc, ks == 0 ? accum : 1u);
            }
          }
          umma_commit_2cta(&empty[stage]);
        }
        accum = 1;
        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
      }
      if (leader) umma_commit_2cta(acc_full);
    }
  } else {
    const int q = warp & 3;
    mbar_wait(acc_full, 0);
    tc_fence_after_sync();
    const float scl = p.inv_scale ? __ldg(p.inv_scale) : 1.0f;
    const int co = co0 + q * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
    for (int kx = 0; kx < KW; ++kx) {
#pragma unroll 1
      for (int c0 = 0; c0 < NCI; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + kx * NCI + c0, v);
        tmem_ld_wait();
        if (co < p.c_out) {
          float* dst = p.dw + ((static_cast<size_t>(co) * p.kh + ky) * p.kw + kx) * p.c_in + ci0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            atomicAdd(reinterpret_cast<float4*>(dst + j),
                      make_float4(__uint_as_float(v[j]) * scl, __uint_as_float(v[j + 1]) * scl,
                                  __uint_as_float(v[j + 2]) * scl, __uint_as_float(v[j + 3]) * scl));
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                                     // nobody leaves while the peer may still signal it
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc_2cta<kTmemCols>(tmem_base);
  }
}

// db[c] += inv_scale * sum over pixels of dz[p][c]     (bias gradient; fp16 in, fp32 atomics out)
__global__ void __launch_bounds__(256)
colsum_f16_kernel(const __half* __restrict__ dz, float* __restrict__ db, long long rows, int c, int c_stride,
                  const float* __restrict__ inv_scale) {
  __shared__ float red[256][8 + 1];
  const int octs = c >> 3;                              // channel octets
  const int lanes_r = 256 / octs;                       // row lanes per CTA (octs <= 256)
  const int oct = threadIdx.x % octs, rl = threadIdx.x / octs;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (rl < lanes_r) {
    for (long long r = static_cast<long long>(blockIdx.x) * lanes_r + rl; r < rows;
         r += static_cast<long long>(gridDim.x) * lanes_r) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(dz + r * c_stride) + oct);
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h[e]);
        acc[2 * e] += f.x;
        acc[2 * e + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[threadIdx.x][e] = acc[e];
  __syncthreads();
  if (threadIdx.x < octs) {
    const float scl = inv_scale ? __ldg(inv_scale) : 1.0f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float s = 0.0f;
      for (int l = 0; l < lanes_r; ++l) s += red[l * octs + threadIdx.x][e];
      atomicAdd(db + threadIdx.x * 8 + e, s * scl);
    }
  }
}

template <int NCI, int KW, bool PAIR64 = false>
int launch_wgrad(const CUtensorMap& tdz, const CUtensorMap& tx, const WgradParams& p, int grid, cudaStream_t st) {
  constexpr int kStageBytes = 2 * kWgABytes + (NCI / 64) * wg_b_bytes(KW);
  int n_stages = (220 * 1024) / kStageBytes;
  if (n_stages > 4) n_stages = 4;
  const size_t smem = static_cast<size_t>(n_stages) * kStageBytes + 1024 + (2 * kWgMaxStages + 1) * 8 + 16;
  DIN_OPT_IN_SMEM((conv_wgrad_kernel<NCI, KW, PAIR64>), smem);
  conv_wgrad_kernel<NCI, KW, PAIR64><<<grid, kWgThreads, smem, st>>>(tdz, tx, p, n_stages);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

template <int KW>
int launch_wgrad_2cta(const CUtensorMap& tdz, const CUtensorMap& tx, const WgradParams& p, int grid, cudaStream_t st) {
  constexpr int kStageBytes = 2 * kWgABytes + wg_b_bytes(KW);
  int n_stages = (216 * 1024) / kStageBytes;
  if (n_stages > 4) n_stages = 4;
  const size_t smem = static_cast<size_t>(n_stages) * kStageBytes + 1024 + (2 * kWgMaxStages + 1) * 8 + 16;
  DIN_OPT_IN_SMEM((conv_wgrad_2cta_kernel<KW>), smem);
  conv_wgrad_2cta_kernel<KW><<<grid, kWgThreads, smem, st>>>(tdz, tx, p, n_stages);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}

}  // namespace

extern "C" int din_conv2d_wgrad_nhwc_f16(const void* x, const void* dz, float* dw, float* dbias,
                                         const float* inv_scale, int n, int h, int w, int c_in, int x_c_stride,
                                         int c_out, int dz_c_stride, int kh, int kw, int pad_h, int pad_w,
                                         void* stream) {
  DIN_CHECK_ARG(x && dz && dw, "din_conv2d_wgrad_nhwc_f16: null pointer");
  DIN_CHECK_ARG(n > 0 && h > 0 && w > 0, "din_conv2d_wgrad_nhwc_f16: bad extent n=%d h=%d w=%d", n, h, w);
  DIN_CHECK_ARG((kw == 1 || kw == 3 || kw == 5 || kw == 7) && kh >= 1 && kh <= 7,
                "din_conv2d_wgrad_nhwc_f16: stride-1 filters with kw in {1,3,5,7} and kh <= 7 (got %dx%d)", kh, kw);
  DIN_CHECK_ARG(pad_h >= 0 && pad_h < kh + 2 && pad_w >= 0 && pad_w < kw + 2, "din_conv2d_wgrad_nhwc_f16: bad padding");
  DIN_CHECK_ARG(c_in > 0 && c_in % 8 == 0 && x_c_stride >= c_in && x_c_stride % 8 == 0,
                "din_conv2d_wgrad_nhwc_f16: c_in=%d must be a multiple of 8 (x_c_stride=%d)", c_in, x_c_stride);
  DIN_CHECK_ARG(c_out > 0 && c_out % 8 == 0 && dz_c_stride >= c_out && dz_c_stride % 8 == 0,
                "din_conv2d_wgrad_nhwc_f16: c_out=%d must be a multiple of 8 (dz_c_stride=%d)", c_out, dz_c_stride);
  DIN_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(dw)) & 15) == 0,
                "din_conv2d_wgrad_nhwc_f16: pointers must be 16-byte aligned");
  const int oh = h + 2 * pad_h - kh + 1, ow = w + 2 * pad_w - kw + 1;
  DIN_CHECK_ARG(oh > 0 && ow > 0, "din_conv2d_wgrad_nhwc_f16: empty output");
  const int sms = din_num_sms();
  DIN_CHECK_ARG(sms > 0, "din_conv2d_wgrad_nhwc_f16: no CUDA device");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  const int nci = (c_in % 128 == 0 && kw <= 3) ? 128 : 64;        // kw * nci fp32 accumulator columns <= 512
  // CTA pair (M = 256 output channels per MMA): see conv_wgrad_2cta_kernel.  DIN_WGRAD_2CTA=0 keeps the one-CTA kernel (A/B).
  bool two_cta = nci == 128 && c_out > 128 && (kw == 1 || kw == 3);
  {
    const char* e = std::getenv("DIN_WGRAD_2CTA");
    if (e && e[0] == '0') two_cta = false;
  }
  WgradParams p{};
  p.c_in = c_in; p.c_out = c_out; p.kh = kh; p.kw = kw; p.pad_h = pad_h; p.pad_w = pad_w;
  p.tiles_x = (ow + 7) / 8;
  p.tiles_per_img = p.tiles_x * ((oh + 15) / 16);
  const long long tiles = static_cast<long long>(n) * p.tiles_per_img;
  DIN_CHECK_ARG(tiles < INT32_MAX, "din_conv2d_wgrad_nhwc_f16: too many tiles");
  p.num_tiles = static_cast<int>(tiles);
  p.n_ci_blk = (c_in + nci - 1) / nci;
  // c_out <= 64 with a 3x3 filter: two taps share one MMA (see PAIR64).  DIN_WGRAD_PAIR64=0: the zero-filled upper half (A/B).
  bool pair64 = !two_cta && c_out <= 64 && kw == 3;
  {
    const char* e = std::getenv("DIN_WGRAD_PAIR64");
    if (e && e[0] == '0') pair64 = false;
  }
  if (pair64) {                                             // the shifted windows need one more pixel column of tiles
    p.tiles_x = (ow + 1 + 7) / 8;
    p.tiles_per_img = p.tiles_x * ((oh + 15) / 16);
    p.num_tiles = static_cast<int>(static_cast<long long>(n) * p.tiles_per_img);
  }
  const int co_blk = two_cta ? 256 : 128;
  const int units = ((c_out + co_blk - 1) / co_blk) * p.n_ci_blk * kh;     // (pairs of) CTAs before the pixel split
  int splits = ((two_cta ? sms : 2 * sms) + units - 1) / units;
  if (splits > p.num_tiles) splits = p.num_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.num_tiles + splits - 1) / splits;
  p.splits = (p.num_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.dw = dw; p.inv_scale = inv_scale;

  CUtensorMap tdz, tx;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(c_out), static_cast<uint64_t>(ow), static_cast<uint64_t>(oh),
                              static_cast<uint64_t>(n)};
    const uint64_t cs = static_cast<uint64_t>(dz_c_stride) * 2;
    const uint64_t strides[4] = {2, cs, cs * ow, cs * ow * oh};
    const uint32_t box[4] = {64, 8, 16, 1};
    const uint32_t es[4] = {1, 1, 1, 1};
    int rc = din_encode_tmap(&tdz, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(dz), dims, strides, box, es,
                             CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != DIN_OK) return rc;
  }
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(c_in), static_cast<uint64_t>(w), static_cast<uint64_t>(h),
                              static_cast<uint64_t>(n)};
    const uint64_t cs = static_cast<uint64_t>(x_c_stride) * 2;
    const uint64_t strides[4] = {2, cs, cs * w, cs * w * h};
    const uint32_t box[4] = {64, static_cast<uint32_t>(8 + kw - 1), 16, 1};
    const uint32_t es[4] = {1, 1, 1, 1};
    int rc = din_encode_tmap(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                             CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != DIN_OK) return rc;
  }
  const int grid = units * p.splits;
  int rc;
  if (two_cta) {
    rc = kw == 1 ? launch_wgrad_2cta<1>(tdz, tx, p, 2 * grid, st) : launch_wgrad_2cta<3>(tdz, tx, p, 2 * grid, st);
  } else
  switch (kw) {
    case 1: rc = (nci == 128) ? launch_wgrad<128, 1>(tdz, tx, p, grid, st) : launch_wgrad<64, 1>(tdz, tx, p, grid, st); break;
    case 3:
      if (pair64) rc = (nci == 128) ? launch_wgrad<128, 3, true>(tdz, tx, p, grid, st) : launch_wgrad<64, 3, true>(tdz, tx, p, grid, st);
      else rc = (nci == 128) ? launch_wgrad<128, 3>(tdz, tx, p, grid, st) : launch_wgrad<64, 3>(tdz, tx, p, grid, st);
      break;
    case 5: rc = launch_wgrad<64, 5>(tdz, tx, p, grid, st); break;
    default: rc = launch_wgrad<64, 7>(tdz, tx, p, grid, st); break;
  }
  if (rc != DIN_OK) return rc;
  if (dbias != nullptr) {
    DIN_CHECK_ARG(c_out <= 2048, "din_conv2d_wgrad_nhwc_f16: c_out=%d too large for the bias reduction", c_out);
    const long long rows = static_cast<long long>(n) * oh * ow;
    int g = 2 * sms;
    const long long per = 256 / (c_out / 8);
    if (static_cast<long long>(g) * per > rows) g = static_cast<int>((rows + per - 1) / per);
    colsum_f16_kernel<<<g, 256, 0, st>>>(static_cast<const __half*>(dz), dbias, rows, c_out, dz_c_stride, inv_scale);
    DIN_CHECK_CUDA(cudaGetLastError());
  }
  return DIN_OK;
}

extern "C" int din_colsum_nhwc_f16(const void* dz, float* db, long long rows, int c, int c_stride, const float* inv_scale,
                                   void* stream) {
  const char* who = "din_colsum_nhwc_f16";
  DIN_CHECK_ARG(dz && db, "%s: null pointer", who);
  DIN_CHECK_ARG(rows > 0 && c > 0 && c % 8 == 0 && c <= 2048 && c_stride >= c && c_stride % 8 == 0, "%s: bad shape", who);
  DIN_CHECK_ARG((reinterpret_cast<uintptr_t>(dz) & 15) == 0, "%s: dz must be 16-byte aligned", who);
  const int sms = din_num_sms();
  int g = 2 * (sms > 0 ? sms : 148);
  const long long per = 256 / (c / 8);
  if (static_cast<long long>(g) * per > rows) g = static_cast<int>((rows + per - 1) / per);
  colsum_f16_kernel<<<g, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(dz), db, rows, c, c_stride,
                                                                     inv_scale);
  DIN_CHECK_CUDA(cudaGetLastError());
  return DIN_OK;
}
