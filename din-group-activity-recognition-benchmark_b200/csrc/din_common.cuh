// din_common.cuh — shared device/host helpers for libdin_sm100.so (sm_100a only).
//
// Thin inline-PTX wrappers for the Blackwell primitives the hot path uses:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and
// the UMMA shared-memory / instruction descriptors.  Nothing here allocates.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/din_sm100.h"

// ----------------------------------------------------------------------------------------------
// host-side error plumbing (defined in api.cu)
// ----------------------------------------------------------------------------------------------
int din_set_error(int code, const char* fmt, ...);

#define DIN_CHECK_ARG(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return din_set_error(DIN_ERR_INVALID_ARG, __VA_ARGS__); \
  } while (0)

#define DIN_CHECK_CUDA(expr)                                                             \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return din_set_error(DIN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                 \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                  \
  } while (0)

// Dynamic shared memory above 48 KB is an opt-in per (function, device).  The opt-in is made once -- and raised only
// when a launch needs more than any earlier one -- instead of on every launch: in the 130-260-launch training steps
// the per-launch cudaFuncSetAttribute calls were a visible share of the host issue time.
#define DIN_OPT_IN_SMEM(kernel, bytes)                                                                   \
  do {                                                                                                   \
    static std::atomic<int> _opted[64];                                                                  \
    int _dev = 0;                                                                                        \
    DIN_CHECK_CUDA(cudaGetDevice(&_dev));                                                                \
    const int _want = static_cast<int>(bytes);                                                           \
    if (_opted[_dev & 63].load(std::memory_order_relaxed) < _want) {                                     \
      DIN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, _want));  \
      _opted[_dev & 63].store(_want, std::memory_order_relaxed);                                         \
    }                                                                                                    \
  } while (0)

// Encodes a tiled tensor map through the driver entry point (no -lcuda link dependency).
// dims/strides innermost-first; strides[0] is implied (element size). Returns 0 or a DIN_ERR_*.
int din_encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, void* base,
                    const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                    const uint32_t* elem_strides, CUtensorMapSwizzle swizzle);

int din_num_sms();

// tensor-core stem convolution (stem_tc.cu); arguments already validated by din_stem_conv_nchw_f32
int din_stem_tc_launch(const void* x, int x_is_u8, const float* w, const float* bias, void* y, int n, int h,
                       int w_in, int c_out, int kh, int kw, int stride, int pad, int relu, int prep,
                       cudaStream_t st);

// ResNet-18 stem + 3x3/2 max-pool in one launch (stem_tc.cu); DIN_ERR_UNSUPPORTED for image rows TMA cannot address
int din_stem_pool_tc_launch(const void* x, int x_is_u8, const float* w, const float* bias, void* y, int n, int h, int w_in,
                            int prep, cudaStream_t st);

// tensor-core stem weight gradient (stem_tc.cu); arguments already validated by din_stem_wgrad
int din_stem_wgrad_tc_launch(const void* x, int x_is_u8, const void* dz, float* dw, float* dbias, const float* inv_scale,
                             int n, int h, int w_in, int c_out, int kh, int stride, int pad, int prep, cudaStream_t st);

#ifdef __CUDACC__
namespace din {

// ----------------------------------------------------------------------------------------------
// generic helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// fp32 pair -> packed fp16x2 with round-to-nearest, optionally fused ReLU (cvt.rn.relu.f16x2.f32):
// the low half holds `lo`.  ReLU commutes with the (monotonic) rounding, so this equals relu-then-round.
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi, bool relu) {
  uint32_t r;
  if (relu) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// Dynamic shared memory aligned up WITHOUT leaving the shared address space: offsetting the extern array by
// an integer keeps LDS/STS code generation (a uintptr_t round-trip degrades every access to generic LD/ST).
__device__ __forceinline__ uint8_t* align_smem(uint8_t* base, uint32_t alignment) {
  const uint32_t a = smem_u32(base);
  return base + (((a + alignment - 1) & ~(alignment - 1)) - a);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded waits: a mis-programmed pipeline traps (launch fails with an error the host reports) instead of hanging the
// device.
//   mbar_wait          tight poll: for the one warp whose wake-up latency is on the critical path (the MMA issuer).
//   mbar_wait_relaxed  poll, then back off with nanosleep: for every other role.  An mbarrier poll is a shared-memory
//                      operation plus ~4 instructions of loop; ncu on the fused conv1 kernel counted ~640 polls and
//                      ~2800 poll-loop instructions per tile from its ~15 waiting warps -- a quarter of the LSU
//                      wavefronts of a kernel whose limit IS the shared-memory data pipe (tensor-core operand reads +
//                      LSU traffic = 1 wavefront per cycle), and half of all instructions issued.  (try_wait's
//                      suspend-time hint was tried first: on this part it does not lengthen the hardware wait.)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    asm volatile("nanosleep.u32 256;" ::: "memory");
    if (++spins > (1u << 24)) __trap();
  }
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------------
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_in_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_in_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (fp16/bf16 operands, fp32 accumulate). One thread issues.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on `bar` once every tcgen05.mma issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread l of the warp receives TMEM lane (base_lane + l).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// CTA pair (cluster of 2, cta_group::2): the two SMs of a TPC execute ONE tcgen05.mma with M = 256; each CTA
// holds its 128 rows of A and half of B's N rows in its own shared memory, and its 128 rows of D in its own TMEM.
// Only the leader (cluster rank 0) issues MMAs; both CTAs issue their TMA loads, whose completion bytes land on
// the LEADER's mbarrier (the shared::cluster address with the peer bit cleared).
// ----------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address -> the same offset in the pair's even CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `target_rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t target_rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(target_rank)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(void* dst, const CUtensorMap* m, uint64_t* leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2cta(void* dst, const CUtensorMap* m, uint64_t* leader_bar, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_in_smem) {   // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_in_smem)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t tmem_addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once every MMA issued so far has completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

// UMMA shared-memory descriptor, K-major operand, 128-byte swizzle (the layout a TMA box with a
// 128-byte inner extent and CU_TENSOR_MAP_SWIZZLE_128B lands in): rows of 128 B, 8-row groups
// 1024 B apart (SBO), descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);  // start address      [0,14)
  d |= static_cast<uint64_t>(1u) << 16;                    // LBO (unused, K-major swizzled) [16,30)
  d |= static_cast<uint64_t>(1024u >> 4) << 32;            // SBO = 1024 B       [32,46)
  d |= static_cast<uint64_t>(1u) << 46;                    // version            [46,48)
  d |= static_cast<uint64_t>(2u) << 61;                    // SWIZZLE_128B       [61,64)
  return d;
}

// Instruction descriptor for kind::f16: A,B = fp16 (format 0), D = fp32 (1), both K-major,
// N>>3 at bit 17, M>>4 at bit 24.
__host__ __device__ constexpr uint32_t umma_idesc_f16_f32(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// ----------------------------------------------------------------------------------------------
// 3-channel stem helpers shared by stem_tc.cu and the fused conv1_1 + conv1_2 kernel (conv_tcgen05.cu)
// ----------------------------------------------------------------------------------------------
// canonical no-swizzle K-major layout: 16-byte chunk kc of row r of an [rows x k_pad] operand
__device__ __forceinline__ uint32_t canon_off(int r, int kc, int sbo_bytes) {
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

// prep_images with the contraction spelled out (one FMA, one exact doubling), so that every instantiation
// rounds identically: (x/255 - 0.5)*2 (utils.py:14-17); the product by 1/255 differs from the division by at
// most 1 ulp(fp32), far below the fp16 rounding applied next.
// Branch-free: without prep the constants are (1, 0, 1), which reproduce f exactly.
__device__ __forceinline__ float prep_value(float f, bool prep) {
  const float a = prep ? 1.0f / 255.0f : 1.0f, b = prep ? -0.5f : 0.0f, c = prep ? 2.0f : 1.0f;
  return __fmul_rn(__fmaf_rn(f, a, b), c);
}

}  // namespace din
#endif  // __CUDACC__
