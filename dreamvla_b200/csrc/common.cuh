// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers,
// bf16 packing, warp reductions and a counter-based RNG for dropout.
// Everything here is hand-written inline PTX for sm_100a (no CUTLASS dependency in the build).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace dvla {

typedef __nv_bfloat16 bf16;

#ifndef DVLA_WATCHDOG_CYCLES
#define DVLA_WATCHDOG_CYCLES (4000000000ll)  // ~2 s at 2 GHz: turn a protocol hang into a trap
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FFu) == 0) {                  // watchdog: look at the clock only every 1024 failed polls
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > DVLA_WATCHDOG_CYCLES) {
        printf("dvla: mbarrier watchdog (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
        __trap();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), 2-D and 3-D tiled loads into shared memory, completion on mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {  // whole warp, .sync.aligned
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- cta_group::2 (CTA pair) variants --------------------------------------------------------------------------------
// In a 2-CTA cluster the shared::cluster address of an object in CTA r is (cta-local address | r << 24); masking bit 24
// addresses the leader (rank 0) copy.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load executed by either CTA of the pair into ITS OWN smem, completing tx bytes on the LEADER's mbarrier.
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the LEADER CTA's copy of `bar` (callable from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs of this thread completed) on `bar` in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
               : "memory");
}

// TMEM -> registers: 32 lanes x 32 consecutive fp32 columns; thread t of the warp receives lane (base_lane + t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   bits [0,14)  start address >> 4      bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4 bits [46,48) version = 1     bits [61,64) layout type (2 = SWIZZLE_128B)
// K-major tile  (rows x 64 bf16, one 128 B swizzle row per matrix row): SBO = 1024 (8 rows), LBO unused (=1).
// MN-major tile (64-wide MN atoms, K rows of 128 B): SBO = 1024 (8 K-rows), LBO = bytes between 64-wide MN atoms.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// tcgen05 instruction descriptor for kind::f16: bf16 x bf16 -> fp32.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                               // c_format = F32
         | (1u << 7)                             // a_format = BF16
         | (1u << 10)                            // b_format = BF16
         | ((a_mn_major ? 1u : 0u) << 15)        // a_major
         | ((b_mn_major ? 1u : 0u) << 16)        // b_major
         | (static_cast<uint32_t>(N >> 3) << 17) // n_dim
         | (static_cast<uint32_t>(M >> 4) << 24);// m_dim
}

// ---------------------------------------------------------------------------------------------
// misc math
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Activations (fp32 math).  Ids are part of the C ABI (include/dvla.h).
enum Act : int { ACT_NONE = 0, ACT_GELU_ERF = 1, ACT_GELU_TANH = 2, ACT_QUICK_GELU = 3, ACT_RELU = 4, ACT_SILU = 5 };

// Branch-free approximate reciprocal / exp2 (single MUFU each).  The IEEE variants (`/`, __frcp_rn, expf) compile to a MUFU
// plus a guarded slow path -- one branch per element -- which made the GEMM epilogue instruction-bound (47 warp
// instructions per output element in the first ncu capture, profiles/r1_gemm_epilogue_v1.txt).
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sigmoid_fast(float x) {   // 1 / (1 + e^-x)
  return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
}
// erf with |abs error| <= 1.5e-7 + MUFU approximation error (Abramowitz & Stegun 7.1.26): ~16 branch-free instructions.
__device__ __forceinline__ float fast_erf(float x) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = ex2_approx(ax * ax * -1.4426950408889634f);
  const float r = fmaf(-p * t, e, 1.0f);
  return copysignf(r, x);
}

__device__ __forceinline__ float act_fwd(float x, int act) {
  switch (act) {
    case ACT_GELU_ERF: {
      const float hx = 0.5f * x;
      return fmaf(hx, fast_erf(x * 0.70710678118654752f), hx);
    }
    case ACT_GELU_TANH: {  // 0.5x(1+tanh(u)) == x*sigmoid(2u)
      const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608028654f);   // sqrt(2/pi)(x + 0.044715x^3)
      return x * sigmoid_fast(2.0f * u);
    }
    case ACT_QUICK_GELU: return x * sigmoid_fast(1.702f * x);
    case ACT_RELU: return fmaxf(x, 0.0f);
    case ACT_SILU: return x * sigmoid_fast(x);
    default: return x;
  }
}
// Apply an activation to n values with the switch hoisted out of the element loop.
template <int N>
__device__ __forceinline__ void act_fwd_n(float (&v)[N], int act) {
  switch (act) {
    case ACT_GELU_ERF:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = act_fwd(v[j], ACT_GELU_ERF);
      break;
    case ACT_GELU_TANH:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = act_fwd(v[j], ACT_GELU_TANH);
      break;
    case ACT_QUICK_GELU:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = act_fwd(v[j], ACT_QUICK_GELU);
      break;
    case ACT_RELU:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = fmaxf(v[j], 0.f);
      break;
    case ACT_SILU:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = act_fwd(v[j], ACT_SILU);
      break;
    default: break;
  }
}

// d act(x) / dx evaluated at the pre-activation x.
__device__ __forceinline__ float act_bwd(float x, int act) {
  switch (act) {
    case ACT_GELU_ERF: {
      const float cdf = fmaf(0.5f, fast_erf(x * 0.70710678118654752f), 0.5f);
      const float pdf = 0.3989422804014327f * ex2_approx(x * x * -0.7213475204444817f);   // exp(-x^2/2)/sqrt(2pi)
      return fmaf(x, pdf, cdf);
    }
    case ACT_GELU_TANH: {
      const float x2 = x * x;
      const float u = x * fmaf(0.0356774081f, x2, 0.7978845608028654f);
      const float s = sigmoid_fast(2.0f * u);
      const float du = fmaf(0.1070322243f, x2, 0.7978845608028654f);            // sqrt(2/pi)(1 + 3*0.044715 x^2)
      return fmaf(x * s * (1.0f - s), 2.0f * du, s);
    }
    case ACT_QUICK_GELU: {
      const float s = sigmoid_fast(1.702f * x);
      return fmaf(1.702f * x * s, 1.0f - s, s);
    }
    case ACT_RELU: return x > 0.0f ? 1.0f : 0.0f;
    case ACT_SILU: {
      const float s = sigmoid_fast(x);
      return fmaf(x * s, 1.0f - s, s);
    }
    default: return 1.0f;
  }
}

// v[j] *= act'(x[j]) with the switch hoisted out of the element loop.
template <int N>
__device__ __forceinline__ void act_bwd_mul_n(float (&v)[N], const float (&x)[N], int act) {
  switch (act) {
    case ACT_GELU_ERF:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] *= act_bwd(x[j], ACT_GELU_ERF);
      break;
    case ACT_GELU_TANH:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] *= act_bwd(x[j], ACT_GELU_TANH);
      break;
    case ACT_QUICK_GELU:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] *= act_bwd(x[j], ACT_QUICK_GELU);
      break;
    case ACT_RELU:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = x[j] > 0.f ? v[j] : 0.f;
      break;
    case ACT_SILU:
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] *= act_bwd(x[j], ACT_SILU);
      break;
    default: break;
  }
}

// Counter-based RNG (Philox-4x32, Salmon et al. SC'11).  One call yields 4 uniform 32-bit words for (seed, offset).
// 7 rounds: the smallest round count the authors report as passing BigCrush (curand's default of 10 adds margin only);
// the dropout masks need independent Bernoulli draws, not cryptographic strength, and the RNG is the larger half of the
// instructions of the attention kernels' dropout path (profiles/r1_notes.md).
constexpr int DVLA_PHILOX_ROUNDS = 7;
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t offset) {
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  uint32_t c0 = static_cast<uint32_t>(offset), c1 = static_cast<uint32_t>(offset >> 32), c2 = 0x243F6A88u, c3 = 0x85A308D3u;
#pragma unroll
  for (int i = 0; i < DVLA_PHILOX_ROUNDS; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// Dropout keep-mask for the 8 consecutive elements [8*blk, 8*blk+8): 16 random bits per element,
// element j of the block is kept iff its 16-bit draw >= thresh (thresh = round(p * 65536)).
// Returned as an 8-bit mask.  Forward and backward regenerate the same mask from (seed, element index).
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint64_t blk, uint32_t thresh) {
  uint4 r = philox4x32(seed, blk);
  uint32_t m = 0;
  m |= ((r.x & 0xFFFFu) >= thresh) << 0; m |= ((r.x >> 16) >= thresh) << 1;
  m |= ((r.y & 0xFFFFu) >= thresh) << 2; m |= ((r.y >> 16) >= thresh) << 3;
  m |= ((r.z & 0xFFFFu) >= thresh) << 4; m |= ((r.z >> 16) >= thresh) << 5;
  m |= ((r.w & 0xFFFFu) >= thresh) << 6; m |= ((r.w >> 16) >= thresh) << 7;
  return m;
}
// The same mask without materialising it: element j of the block whose Philox words are `r` is kept iff its 16-bit draw
// (low half of word j/2 for even j, high half for odd j) >= thresh.  `th16` = thresh << 16, so the high half compares as
// the whole word and the low half after one shift: 1-2 instructions per element instead of ~6 to build and test mask bits.
__device__ __forceinline__ bool dropout_keep_elem(const uint4& r, uint32_t th16, int j) {
  const uint32_t w = j < 2 ? r.x : j < 4 ? r.y : j < 6 ? r.z : r.w;
  return (j & 1) ? (w >= th16) : ((w << 16) >= th16);
}

}  // namespace dvla
