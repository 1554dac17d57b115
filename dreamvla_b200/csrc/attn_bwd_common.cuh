// Shared pieces of the tcgen05 attention-backward kernels (attention_bwd_tc.cu: single-role CTAs, attention_bwd_ws.cu:
// warp-specialised CTAs): parameter block, 4-D TMA load, half-row smem / global writers, TMEM row loads.
#pragma once
#include "common.cuh"
#include "../../include/dvla.h"

namespace dvla {
void set_error(const char* fmt, ...);
void count_launch();
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn_shared();

constexpr float B_LOG2E = 1.4426950408889634f;

struct AttnBwdTcParams {
  bf16 *dq, *dk, *dv;
  const float* lse;
  const float* delta;
  const uint32_t* mask;      // [Lq, mask_words]   (dq kernel)
  const uint32_t* mask_t;    // [Lk, mask_t_words] (dkv kernel)
  const uint8_t* tile_flags; // 64x64 flags [nqt64, nkt64]
  int B, H, Lq, Lk, nkt64, mask_words, mask_t_words;
  long long dq_sb, dq_ss, dq_sh, dk_sb, dk_ss, dk_sh, dv_sb, dv_ss, dv_sh;
  int q_hi, k_hi, v_hi, do_hi;   // tensor-map dim order flags (head-inner)
  float scale;
  float drop_scale; uint32_t drop_thresh; uint64_t drop_seed; const uint64_t* drop_seed_ptr;
};

__device__ __forceinline__ void tma4(void* dst, const CUtensorMap* m, uint64_t* bar, int head_inner, int row0, int h, int b) {
  const int c1 = head_inner ? h : row0, c2 = head_inner ? row0 : h;
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(0), "r"(c1), "r"(c2), "r"(b)
      : "memory");
}
// row `r` of a 128-row x 64-col bf16 A-operand tile (one 128B-swizzle atom): write 64 values
__device__ __forceinline__ void write_row64(uint8_t* tile, int r, const float (&v)[64]) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4)) =
        make_uint4(pack_bf16x2(v[8 * c], v[8 * c + 1]), pack_bf16x2(v[8 * c + 2], v[8 * c + 3]),
                   pack_bf16x2(v[8 * c + 4], v[8 * c + 5]), pack_bf16x2(v[8 * c + 6], v[8 * c + 7]));
}
__device__ __forceinline__ void store_row64(bf16* dst, const float (&v)[64], float scale) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<uint4*>(dst + c * 8) =
        make_uint4(pack_bf16x2(v[8 * c] * scale, v[8 * c + 1] * scale), pack_bf16x2(v[8 * c + 2] * scale, v[8 * c + 3] * scale),
                   pack_bf16x2(v[8 * c + 4] * scale, v[8 * c + 5] * scale), pack_bf16x2(v[8 * c + 6] * scale, v[8 * c + 7] * scale));
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
  uint32_t r0[32], r1[32];
  tmem_ld_32x32(taddr, r0);
  tmem_ld_32x32(taddr + 32, r1);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
}

// 32-column half-row helpers for the 256-thread kernels (threads t and t+128 share row t, columns [32*half, 32*half+32))
__device__ __forceinline__ void write_row32(uint8_t* tile, int r, int half, const float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(tile + r * 128 + (((half * 4 + c) ^ (r & 7)) << 4)) =
        make_uint4(pack_bf16x2(v[8 * c], v[8 * c + 1]), pack_bf16x2(v[8 * c + 2], v[8 * c + 3]),
                   pack_bf16x2(v[8 * c + 4], v[8 * c + 5]), pack_bf16x2(v[8 * c + 6], v[8 * c + 7]));
}
__device__ __forceinline__ void store_row32(bf16* dst, const float (&v)[32], float scale) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(dst + c * 8) =
        make_uint4(pack_bf16x2(v[8 * c] * scale, v[8 * c + 1] * scale), pack_bf16x2(v[8 * c + 2] * scale, v[8 * c + 3] * scale),
                   pack_bf16x2(v[8 * c + 4] * scale, v[8 * c + 5] * scale), pack_bf16x2(v[8 * c + 6] * scale, v[8 * c + 7] * scale));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  tmem_ld_32x32(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

}  // namespace dvla
