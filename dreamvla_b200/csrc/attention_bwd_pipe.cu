// Warp-specialised, software-pipelined tcgen05 attention BACKWARD (head_dim 64): the MMAs of tile j+1 run while the 16 ALU
// warps still work on tile j.
//
// One CTA per SM-resident work item, 17 warps:
//   warps 0..15  ALU: warp w owns TMEM lane quadrant (w & 3) = 32 rows of the 128-row tile and 16 of the 64 streamed columns
//                (w >> 2); tcgen05.ld S / dP -> exp2 / mask / dropout -> P, dS (bf16) into the smem A-operand tiles
//   warp 16      control: lane 0 issues TMA loads and every tcgen05.mma; the warp also stages the per-tile row scalars
// Double-buffered S/dP accumulators in TMEM (2 x 128 columns) and P/dS operand tiles in smem (2 x 32 KB), 3-stage TMA ring
// for the streamed operand; all hand-offs are mbarriers (tcgen05.commit for MMA completion), no __syncthreads in the loop.
//   dKV kernel: CTA = 128 keys (rows), streams 64-query tiles:   S^T = K Q^T, dP^T = V dO^T | dV += P^T dO, dK += dS^T Q
//   dQ  kernel: CTA = 128 queries (rows), streams 64-key tiles:  S = Q K^T,  dP = dO V^T   | dQ += dS K
// Same math, mask, dropout and output conventions as attention_bwd_tc.cu (the non-pipelined tcgen05 kernels).
#include "common.cuh"
#include "../../include/dvla.h"

namespace dvla {
void set_error(const char* fmt, ...);
void count_launch();
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn_shared();

constexpr float P_LOG2E = 1.4426950408889634f;
constexpr int PIPE_THREADS = 17 * 32;

struct AttnPipeParams {
  bf16 *dq, *dk, *dv;
  const float* lse;
  const float* delta;
  const uint32_t* mask;      // [Lq, mask_words]
  const uint32_t* mask_t;    // [Lk, mask_t_words]
  const uint8_t* tile_flags; // 64x64 flags [nqt64, nkt64]
  int B, H, Lq, Lk, nkt64, mask_words, mask_t_words;
  long long dq_sb, dq_ss, dq_sh, dk_sb, dk_ss, dk_sh, dv_sb, dv_ss, dv_sh;
  int q_hi, k_hi, v_hi, do_hi;
  float scale;
  float drop_scale; uint32_t drop_thresh; uint64_t drop_seed; const uint64_t* drop_seed_ptr;
};

__device__ __forceinline__ void p_tma4(void* dst, const CUtensorMap* m, uint64_t* bar, int head_inner, int row0, int h, int b) {
  const int c1 = head_inner ? h : row0, c2 = head_inner ? row0 : h;
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(0), "r"(c1), "r"(c2), "r"(b)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
// 16 bf16 (columns [16*cg, 16*cg+16)) of row r of a 128-row x 64-col A-operand tile (one 128B-swizzle atom)
__device__ __forceinline__ void write_row16(uint8_t* tile, int r, int cg, const float (&v)[16]) {
#pragma unroll
  for (int c = 0; c < 2; ++c)
    *reinterpret_cast<uint4*>(tile + r * 128 + (((cg * 2 + c) ^ (r & 7)) << 4)) =
        make_uint4(pack_bf16x2(v[8 * c], v[8 * c + 1]), pack_bf16x2(v[8 * c + 2], v[8 * c + 3]),
                   pack_bf16x2(v[8 * c + 4], v[8 * c + 5]), pack_bf16x2(v[8 * c + 6], v[8 * c + 7]));
}
__device__ __forceinline__ void store_row16(bf16* dst, const float (&v)[16], float scale) {
#pragma unroll
  for (int c = 0; c < 2; ++c)
    *reinterpret_cast<uint4*>(dst + c * 8) =
        make_uint4(pack_bf16x2(v[8 * c] * scale, v[8 * c + 1] * scale), pack_bf16x2(v[8 * c + 2] * scale, v[8 * c + 3] * scale),
                   pack_bf16x2(v[8 * c + 4] * scale, v[8 * c + 5] * scale), pack_bf16x2(v[8 * c + 6] * scale, v[8 * c + 7] * scale));
}

// shared-memory map (both kernels): resident pair 2 x 16 KB | streamed pair 3 stages x 2 x 8 KB | P[2], dS[2] 4 x 16 KB | aux | bars
constexpr int PS_RES0 = 0, PS_RES1 = 16384, PS_STR0 = 32768, PS_STR1 = PS_STR0 + 3 * 8192, PS_P = PS_STR1 + 3 * 8192,
              PS_DS = PS_P + 2 * 16384, PS_AUX = PS_DS + 2 * 16384, PS_FLAGS = PS_AUX + 3 * 2 * 64 * 4, PS_BAR = PS_FLAGS + 512;
constexpr int ATTN_PIPE_SMEM = PS_BAR + 256 + 1024;
// barrier slots
enum { B_RES = 0, B_QFULL = 1, B_QFREE = 4, B_AUX = 7, B_SFULL = 10, B_SFREE = 12, B_PFULL = 14, B_PFREE = 16, B_DONE = 18, B_COUNT = 19 };

// ============================================================ dK / dV ==============================================================
__global__ void __launch_bounds__(PIPE_THREADS, 1)
attn_bwd_dkv_pipe_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                         const AttnPipeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_lse = reinterpret_cast<float*>(smem + PS_AUX);      // [3][64]
  float* s_delta = s_lse + 3 * 64;                              // [3][64]
  uint8_t* s_flags = smem + PS_FLAGS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PS_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int k0 = kt * 128;
  const int nqt = (p.Lq + 63) / 64;
  const long long bh = static_cast<long long>(b) * p.H + h;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(&bars[B_RES], 1);
    for (int i = 0; i < 3; ++i) { mbar_init(&bars[B_QFULL + i], 1); mbar_init(&bars[B_QFREE + i], 1); mbar_init(&bars[B_AUX + i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[B_SFULL + i], 1); mbar_init(&bars[B_SFREE + i], 16);
      mbar_init(&bars[B_PFULL + i], 16); mbar_init(&bars[B_PFREE + i], 1);
    }
    mbar_init(&bars[B_DONE], 1);
    fence_barrier_init();
  }
  if (warp == 16) tmem_alloc(tmem_slot, 512);
  for (int qi = tid; qi < nqt && qi < 512; qi += PIPE_THREADS) {   // per 64-query tile flag of this 128-key tile
    int f = 2;
    if (p.tile_flags) {
      int any = 0, all = 1;
#pragma unroll
      for (int dk = 0; dk < 2; ++dk) {
        const int k64 = kt * 2 + dk;
        if (k64 * 64 >= p.Lk) continue;
        const int ff = p.tile_flags[static_cast<long long>(qi) * p.nkt64 + k64];
        any |= (ff != 0);
        all &= (ff == 2);
      }
      f = any ? (all ? 2 : 1) : 0;
    }
    s_flags[qi] = static_cast<uint8_t>(f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto next_tile = [&](int qt) {
    while (qt < nqt && s_flags[qt] == 0) ++qt;
    return qt;
  };
  const uint32_t tDV = tmem_base + 256, tDK = tmem_base + 320;
  constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
  constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);

  if (warp == 16) {
    // ------------------------------------------------------- control warp -------------------------------------------------------
    auto load_tile = [&](int qt, int s) {             // whole warp: TMA (lane 0) + row scalars of the 64 queries
      if (lane == 0) {
        mbar_expect_tx(&bars[B_QFULL + s], 16384);
        p_tma4(smem + PS_STR0 + s * 8192, &tmQ, &bars[B_QFULL + s], p.q_hi, qt * 64, h, b);
        p_tma4(smem + PS_STR1 + s * 8192, &tmDO, &bars[B_QFULL + s], p.do_hi, qt * 64, h, b);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int qi = qt * 64 + lane + i * 32;
        s_lse[s * 64 + lane + i * 32] = (qi < p.Lq) ? p.lse[bh * p.Lq + qi] * P_LOG2E : INFINITY;
        s_delta[s * 64 + lane + i * 32] = (qi < p.Lq) ? p.delta[bh * p.Lq + qi] : 0.f;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_AUX + s]);
    };
    if (lane == 0) {
      mbar_expect_tx(&bars[B_RES], 32768);
      p_tma4(smem + PS_RES0, &tmK, &bars[B_RES], p.k_hi, k0, h, b);
      p_tma4(smem + PS_RES1, &tmV, &bars[B_RES], p.v_hi, k0, h, b);
    }
    int qt_load = next_tile(0), n_loaded = 0;
    for (; n_loaded < 3 && qt_load < nqt; ++n_loaded) { load_tile(qt_load, n_loaded); qt_load = next_tile(qt_load + 1); }
    int j = 0;
    for (int qt = next_tile(0); qt < nqt; qt = next_tile(qt + 1), ++j) {
      const int s = j % 3, bb = j & 1;
      if (j == 0) mbar_wait(&bars[B_RES], 0);
      mbar_wait(&bars[B_QFULL + s], (j / 3) & 1);
      mbar_wait(&bars[B_SFREE + bb], ((j >> 1) & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) {      // MMA1_j: S^T = K Q^T, dP^T = V dO^T into accumulator buffer bb
        const uint32_t ka = smem_u32(smem + PS_RES0), va = smem_u32(smem + PS_RES1);
        const uint32_t qa = smem_u32(smem + PS_STR0 + s * 8192), da = smem_u32(smem + PS_STR1 + s * 8192);
        const uint32_t tS = tmem_base + bb * 128, tDP = tS + 64;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tS, make_smem_desc_sw128(ka + k * 32, 16, 1024), make_smem_desc_sw128(qa + k * 32, 16, 1024), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tDP, make_smem_desc_sw128(va + k * 32, 16, 1024), make_smem_desc_sw128(da + k * 32, 16, 1024), idesc_s, k != 0);
        umma_commit(&bars[B_SFULL + bb]);
      }
      __syncwarp();
      if (j >= 1) {         // MMA2_{j-1}: dV += P^T dO, dK += dS^T Q  (P/dS buffer pb, streamed stage ps)
        const int jj = j - 1, pb = jj & 1, ps = jj % 3;
        mbar_wait(&bars[B_PFULL + pb], (jj >> 1) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t pa = smem_u32(smem + PS_P + pb * 16384), dsa = smem_u32(smem + PS_DS + pb * 16384);
          const uint32_t qa = smem_u32(smem + PS_STR0 + ps * 8192), da = smem_u32(smem + PS_STR1 + ps * 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tDV, make_smem_desc_sw128(pa + k * 32, 16, 1024), make_smem_desc_sw128(da + k * 2048, 8192, 1024), idesc_o,
                         (jj > 0 || k != 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tDK, make_smem_desc_sw128(dsa + k * 32, 16, 1024), make_smem_desc_sw128(qa + k * 2048, 8192, 1024), idesc_o,
                         (jj > 0 || k != 0) ? 1u : 0u);
          umma_commit(&bars[B_PFREE + pb]);
          umma_commit(&bars[B_QFREE + ps]);
        }
        __syncwarp();
        if (qt_load < nqt) {          // refill stage ps (tile j+2) once MMA2_{j-1} has drained it
          mbar_wait(&bars[B_QFREE + ps], (jj / 3) & 1);
          load_tile(qt_load, ps);
          qt_load = next_tile(qt_load + 1);
        }
      }
    }
    if (j >= 1) {           // MMA2 of the last tile
      const int jj = j - 1, pb = jj & 1, ps = jj % 3;
      mbar_wait(&bars[B_PFULL + pb], (jj >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t pa = smem_u32(smem + PS_P + pb * 16384), dsa = smem_u32(smem + PS_DS + pb * 16384);
        const uint32_t qa = smem_u32(smem + PS_STR0 + ps * 8192), da = smem_u32(smem + PS_STR1 + ps * 8192);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tDV, make_smem_desc_sw128(pa + k * 32, 16, 1024), make_smem_desc_sw128(da + k * 2048, 8192, 1024), idesc_o,
                       (jj > 0 || k != 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tDK, make_smem_desc_sw128(dsa + k * 32, 16, 1024), make_smem_desc_sw128(qa + k * 2048, 8192, 1024), idesc_o,
                       (jj > 0 || k != 0) ? 1u : 0u);
        umma_commit(&bars[B_DONE]);
      }
      __syncwarp();
    }
  } else {
    // --------------------------------------------------------- ALU warps ---------------------------------------------------------
    const int q = warp & 3, cg = warp >> 2;
    const int rowi = q * 32 + lane;
    const int key = k0 + rowi;
    const uint32_t lane_off = (static_cast<uint32_t>(q * 32) << 16) + cg * 16;
    const float sc = p.scale * P_LOG2E;
    const uint64_t seed = p.drop_seed + ((p.drop_scale != 0.f && p.drop_seed_ptr) ? __ldg(p.drop_seed_ptr) : 0ull);
    const int nblk = (p.Lk + 7) >> 3;
    auto mask_word = [&](int qtile) -> uint32_t {      // this key row's bits for the 32 queries containing my 16
      uint32_t w = 0xffffffffu;
      if (s_flags[qtile] == 1) {
        const int wi = qtile * 2 + (cg >> 1);
        w = (key < p.Lk && wi < p.mask_t_words) ? p.mask_t[static_cast<long long>(key) * p.mask_t_words + wi] : 0u;
      }
      if (key >= p.Lk) w = 0u;
      return w;
    };
    int qt = next_tile(0);
    uint32_t pf_w = (qt < nqt) ? mask_word(qt) : 0u;
    int j = 0;
    for (; qt < nqt; ++j) {
      const int qt_next = next_tile(qt + 1);
      const int s = j % 3, bb = j & 1;
      const uint32_t wv = pf_w >> ((cg & 1) * 16);
      if (qt_next < nqt) pf_w = mask_word(qt_next);
      const int q0 = qt * 64;
      mbar_wait(&bars[B_AUX + s], (j / 3) & 1);
      mbar_wait(&bars[B_SFULL + bb], (j >> 1) & 1);
      tc_fence_after();
      float sv[16], dpv[16];
      tmem_ld_32x16(tmem_base + bb * 128 + lane_off, sv);
      tmem_ld_32x16(tmem_base + bb * 128 + 64 + lane_off, dpv);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_SFREE + bb]);
      const float* lse_t = s_lse + s * 64 + cg * 16;
      const float* dl_t = s_delta + s * 64 + cg * 16;
      uint32_t keepm[2];
      if (p.drop_scale != 0.f) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const long long qc = q0 + cg * 16 + (lane & 7) + i * 8;
          keepm[i] = dropout_keep8(seed, (bh * p.Lq + qc) * nblk + (key >> 3), p.drop_thresh);
        }
      }
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const bool vis = (wv >> c) & 1u;
        const float pv = vis ? ex2_approx(fmaf(sv[c], sc, -lse_t[c])) : 0.f;
        float dp = dpv[c];
        if (p.drop_scale != 0.f) {
          const uint32_t m = __shfl_sync(0xffffffffu, keepm[c >> 3], (lane & ~7) | (c & 7));
          const bool kp = (m >> (key & 7)) & 1u;
          dp = kp ? dp * p.drop_scale : 0.f;
          dpv[c] = pv * (dp - dl_t[c]);
          sv[c] = kp ? pv * p.drop_scale : 0.f;
        } else {
          dpv[c] = pv * (dp - dl_t[c]);
          sv[c] = pv;
        }
      }
      mbar_wait(&bars[B_PFREE + bb], ((j >> 1) & 1) ^ 1);
      write_row16(smem + PS_P + bb * 16384, rowi, cg, sv);
      write_row16(smem + PS_DS + bb * 16384, rowi, cg, dpv);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_PFULL + bb]);
      qt = qt_next;
    }
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
    if (j > 0) {
      mbar_wait(&bars[B_DONE], 0);
      tc_fence_after();
      tmem_ld_32x16(tDV + lane_off, v);
    }
    if (key < p.Lk) store_row16(p.dv + b * p.dv_sb + static_cast<long long>(key) * p.dv_ss + h * p.dv_sh + cg * 16, v, 1.0f);
    if (j > 0) tmem_ld_32x16(tDK + lane_off, v);
    if (key < p.Lk) store_row16(p.dk + b * p.dk_sb + static_cast<long long>(key) * p.dk_ss + h * p.dk_sh + cg * 16, v, p.scale);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 16) tmem_dealloc(tmem_base, 512);
}

// ============================================================== dQ =================================================================
__global__ void __launch_bounds__(PIPE_THREADS, 1)
attn_bwd_dq_pipe_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                        const AttnPipeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_flags = smem + PS_FLAGS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PS_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * 128;
  const int nkt = (p.Lk + 63) / 64;
  const long long bh = static_cast<long long>(b) * p.H + h;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(&bars[B_RES], 1);
    for (int i = 0; i < 3; ++i) { mbar_init(&bars[B_QFULL + i], 1); mbar_init(&bars[B_QFREE + i], 1); mbar_init(&bars[B_AUX + i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[B_SFULL + i], 1); mbar_init(&bars[B_SFREE + i], 16);
      mbar_init(&bars[B_PFULL + i], 16); mbar_init(&bars[B_PFREE + i], 1);
    }
    mbar_init(&bars[B_DONE], 1);
    fence_barrier_init();
  }
  if (warp == 16) tmem_alloc(tmem_slot, 512);
  for (int ki = tid; ki < nkt && ki < 512; ki += PIPE_THREADS) {
    int f = 2;
    if (p.tile_flags) {
      int any = 0, all = 1;
#pragma unroll
      for (int dq = 0; dq < 2; ++dq) {
        const int q64 = qt * 2 + dq;
        if (q64 * 64 >= p.Lq) continue;
        const int ff = p.tile_flags[static_cast<long long>(q64) * p.nkt64 + ki];
        any |= (ff != 0);
        all &= (ff == 2);
      }
      f = any ? (all ? 2 : 1) : 0;
    }
    s_flags[ki] = static_cast<uint8_t>(f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto next_tile = [&](int kt) {
    while (kt < nkt && s_flags[kt] == 0) ++kt;
    return kt;
  };
  const uint32_t tDQ = tmem_base + 256;
  constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
  constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);

  if (warp == 16) {
    auto load_tile = [&](int kt, int s) {
      if (lane == 0) {
        mbar_expect_tx(&bars[B_QFULL + s], 16384);
        p_tma4(smem + PS_STR0 + s * 8192, &tmK, &bars[B_QFULL + s], p.k_hi, kt * 64, h, b);
        p_tma4(smem + PS_STR1 + s * 8192, &tmV, &bars[B_QFULL + s], p.v_hi, kt * 64, h, b);
      }
      __syncwarp();
    };
    if (lane == 0) {
      mbar_expect_tx(&bars[B_RES], 32768);
      p_tma4(smem + PS_RES0, &tmQ, &bars[B_RES], p.q_hi, q0, h, b);
      p_tma4(smem + PS_RES1, &tmDO, &bars[B_RES], p.do_hi, q0, h, b);
    }
    int kt_load = next_tile(0), n_loaded = 0;
    for (; n_loaded < 3 && kt_load < nkt; ++n_loaded) { load_tile(kt_load, n_loaded); kt_load = next_tile(kt_load + 1); }
    int j = 0;
    for (int kt = next_tile(0); kt < nkt; kt = next_tile(kt + 1), ++j) {
      const int s = j % 3, bb = j & 1;
      if (j == 0) mbar_wait(&bars[B_RES], 0);
      mbar_wait(&bars[B_QFULL + s], (j / 3) & 1);
      mbar_wait(&bars[B_SFREE + bb], ((j >> 1) & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) {      // MMA1_j: S = Q K^T, dP = dO V^T
        const uint32_t qa = smem_u32(smem + PS_RES0), da = smem_u32(smem + PS_RES1);
        const uint32_t ka = smem_u32(smem + PS_STR0 + s * 8192), va = smem_u32(smem + PS_STR1 + s * 8192);
        const uint32_t tS = tmem_base + bb * 128, tDP = tS + 64;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tS, make_smem_desc_sw128(qa + k * 32, 16, 1024), make_smem_desc_sw128(ka + k * 32, 16, 1024), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tDP, make_smem_desc_sw128(da + k * 32, 16, 1024), make_smem_desc_sw128(va + k * 32, 16, 1024), idesc_s, k != 0);
        umma_commit(&bars[B_SFULL + bb]);
      }
      __syncwarp();
      if (j >= 1) {         // MMA2_{j-1}: dQ += dS K
        const int jj = j - 1, pb = jj & 1, ps = jj % 3;
        mbar_wait(&bars[B_PFULL + pb], (jj >> 1) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t dsa = smem_u32(smem + PS_DS + pb * 16384), ka = smem_u32(smem + PS_STR0 + ps * 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tDQ, make_smem_desc_sw128(dsa + k * 32, 16, 1024), make_smem_desc_sw128(ka + k * 2048, 8192, 1024), idesc_o,
                         (jj > 0 || k != 0) ? 1u : 0u);
          umma_commit(&bars[B_PFREE + pb]);
          umma_commit(&bars[B_QFREE + ps]);
        }
        __syncwarp();
        if (kt_load < nkt) {
          mbar_wait(&bars[B_QFREE + ps], (jj / 3) & 1);
          load_tile(kt_load, ps);
          kt_load = next_tile(kt_load + 1);
        }
      }
    }
    if (j >= 1) {
      const int jj = j - 1, pb = jj & 1, ps = jj % 3;
      mbar_wait(&bars[B_PFULL + pb], (jj >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t dsa = smem_u32(smem + PS_DS + pb * 16384), ka = smem_u32(smem + PS_STR0 + ps * 8192);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tDQ, make_smem_desc_sw128(dsa + k * 32, 16, 1024), make_smem_desc_sw128(ka + k * 2048, 8192, 1024), idesc_o,
                       (jj > 0 || k != 0) ? 1u : 0u);
        umma_commit(&bars[B_DONE]);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3, cg = warp >> 2;
    const int rowi = q * 32 + lane;
    const int row = q0 + rowi;
    const uint32_t lane_off = (static_cast<uint32_t>(q * 32) << 16) + cg * 16;
    const float sc = p.scale * P_LOG2E;
    const float lse2 = (row < p.Lq) ? p.lse[bh * p.Lq + row] * P_LOG2E : INFINITY;
    const float dlt = (row < p.Lq) ? p.delta[bh * p.Lq + row] : 0.f;
    const uint64_t seed = p.drop_seed + ((p.drop_scale != 0.f && p.drop_seed_ptr) ? __ldg(p.drop_seed_ptr) : 0ull);
    const int nblk = (p.Lk + 7) >> 3;
    auto mask_word = [&](int ktile) -> uint32_t {
      uint32_t w = 0xffffffffu;
      if (s_flags[ktile] == 1) {
        const int wi = ktile * 2 + (cg >> 1);
        w = (row < p.Lq && wi < p.mask_words) ? p.mask[static_cast<long long>(row) * p.mask_words + wi] : 0u;
      }
      return w;
    };
    int kt = next_tile(0);
    uint32_t pf_w = (kt < nkt) ? mask_word(kt) : 0u;
    int j = 0;
    for (; kt < nkt; ++j) {
      const int kt_next = next_tile(kt + 1);
      const int bb = j & 1;
      uint32_t wv = pf_w >> ((cg & 1) * 16);
      if (kt_next < nkt) pf_w = mask_word(kt_next);
      const int kc0 = kt * 64 + cg * 16;              // first key of my 16 columns
      {
        const int nv = p.Lk - kc0;
        if (nv < 16) wv &= (nv <= 0) ? 0u : ((1u << nv) - 1u);
      }
      mbar_wait(&bars[B_SFULL + bb], (j >> 1) & 1);
      tc_fence_after();
      float sv[16], dpv[16];
      tmem_ld_32x16(tmem_base + bb * 128 + lane_off, sv);
      tmem_ld_32x16(tmem_base + bb * 128 + 64 + lane_off, dpv);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_SFREE + bb]);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t keep = 0xffu;
        if (p.drop_scale != 0.f) keep = dropout_keep8(seed, (bh * p.Lq + row) * nblk + (kc0 >> 3) + g, p.drop_thresh);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int c = g * 8 + jj;
          const bool vis = (wv >> c) & 1u;
          const float pv = vis ? ex2_approx(fmaf(sv[c], sc, -lse2)) : 0.f;
          float dp = dpv[c];
          if (p.drop_scale != 0.f) dp = ((keep >> jj) & 1u) ? dp * p.drop_scale : 0.f;
          sv[c] = pv * (dp - dlt);
        }
      }
      mbar_wait(&bars[B_PFREE + bb], ((j >> 1) & 1) ^ 1);
      write_row16(smem + PS_DS + bb * 16384, rowi, cg, sv);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_PFULL + bb]);
      kt = kt_next;
    }
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
    if (j > 0) {
      mbar_wait(&bars[B_DONE], 0);
      tc_fence_after();
      tmem_ld_32x16(tDQ + lane_off, v);
    }
    if (row < p.Lq) store_row16(p.dq + b * p.dq_sb + static_cast<long long>(row) * p.dq_ss + h * p.dq_sh + cg * 16, v, p.scale);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 16) tmem_dealloc(tmem_base, 512);
}

// ============================================================== host ==============================================================
static bool make_tmap_rows_p(CUtensorMap* out, const void* base, long long L, long long H, long long B, long long ss, long long sh,
                             long long sb, int box_rows, int* head_inner) {
  EncodeTiledFn fn = get_encode_fn_shared();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not found"); return false; }
  *head_inner = (sh <= ss) ? 1 : 0;
  cuuint64_t dims[4]; cuuint64_t strides[3]; cuuint32_t box[4]; cuuint32_t estr[4] = {1, 1, 1, 1};
  dims[0] = 64; box[0] = 64;
  if (*head_inner) { dims[1] = H; dims[2] = L; strides[0] = sh * 2; strides[1] = ss * 2; box[1] = 1; box[2] = box_rows; }
  else             { dims[1] = L; dims[2] = H; strides[0] = ss * 2; strides[1] = sh * 2; box[1] = box_rows; box[2] = 1; }
  dims[3] = B; strides[2] = (B > 1 ? sb : (long long)L * H * 64) * 2; box[3] = 1;
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("attn pipe tensor map encode failed (%d)", (int)r); return false; }
  return true;
}

int attn_bwd_pipe_dispatch(const dvla_attn_bwd_args* a, const uint32_t* mask_t, int mask_t_words, cudaStream_t s) {
  auto okst = [&](long long ss, long long sh, long long sb) {
    return ss > 0 && sh > 0 && (a->B == 1 || sb > 0) && ss % 8 == 0 && sh % 8 == 0 && sb % 8 == 0;
  };
  if (!okst(a->q_ss, a->q_sh, a->q_sb) || !okst(a->k_ss, a->k_sh, a->k_sb) || !okst(a->v_ss, a->v_sh, a->v_sb) ||
      !okst(a->do_ss, a->do_sh, a->do_sb))
    return DVLA_ERR_UNSUPPORTED;
  if (a->mask && !mask_t) return DVLA_ERR_UNSUPPORTED;
  AttnPipeParams p;
  memset(&p, 0, sizeof(p));
  p.dq = (bf16*)a->dq; p.dk = (bf16*)a->dk; p.dv = (bf16*)a->dv; p.lse = a->lse; p.delta = a->delta;
  p.mask = a->mask; p.mask_t = mask_t; p.tile_flags = a->mask ? a->tile_flags : nullptr;
  p.B = (int)a->B; p.H = (int)a->H; p.Lq = (int)a->Lq; p.Lk = (int)a->Lk; p.nkt64 = (p.Lk + 63) / 64;
  p.mask_words = a->mask_words; p.mask_t_words = mask_t_words;
  p.dq_sb = a->dq_sb; p.dq_ss = a->dq_ss; p.dq_sh = a->dq_sh; p.dk_sb = a->dk_sb; p.dk_ss = a->dk_ss; p.dk_sh = a->dk_sh;
  p.dv_sb = a->dv_sb; p.dv_ss = a->dv_ss; p.dv_sh = a->dv_sh;
  p.scale = a->scale;
  if (a->dropout_p > 0.f) {
    p.drop_thresh = (uint32_t)(a->dropout_p * 65536.0f + 0.5f);
    p.drop_scale = 1.0f / (1.0f - (float)p.drop_thresh / 65536.0f);
    p.drop_seed = a->dropout_seed;
    p.drop_seed_ptr = a->dropout_seed_ptr;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(attn_bwd_dkv_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_PIPE_SMEM);
    cudaError_t e2 = cudaFuncSetAttribute(attn_bwd_dq_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_PIPE_SMEM);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("attn_bwd_pipe smem attr failed"); return DVLA_ERR_CUDA; }
    attr_set = true;
  }
  CUtensorMap q64, do64, k128, v128, q128, do128, k64, v64;
  int hi;
  if (!make_tmap_rows_p(&q64, a->q, a->Lq, a->H, a->B, a->q_ss, a->q_sh, a->q_sb, 64, &p.q_hi)) return DVLA_ERR_CUDA;
  if (!make_tmap_rows_p(&do64, a->d_o, a->Lq, a->H, a->B, a->do_ss, a->do_sh, a->do_sb, 64, &p.do_hi)) return DVLA_ERR_CUDA;
  if (!make_tmap_rows_p(&k128, a->k, a->Lk, a->H, a->B, a->k_ss, a->k_sh, a->k_sb, 128, &p.k_hi)) return DVLA_ERR_CUDA;
  if (!make_tmap_rows_p(&v128, a->v, a->Lk, a->H, a->B, a->v_ss, a->v_sh, a->v_sb, 128, &p.v_hi)) return DVLA_ERR_CUDA;
  if (!make_tmap_rows_p(&q128, a->q, a->Lq, a->H, a->B, a->q_ss, a->q_sh, a->q_sb, 128, &hi)) return DVLA_ERR_CUDA;
  if (!make_tmap_rows_p(&do128, a->d_o, a->Lq, a->H, a->B, a->do_ss, a->do_sh, a->do_sb, 128, &hi)) return DVLA_ERR_CUDA;
  if (!make_tmap_rows_p(&k64, a->k, a->Lk, a->H, a->B, a->k_ss, a->k_sh, a->k_sb, 64, &hi)) return DVLA_ERR_CUDA;
  if (!make_tmap_rows_p(&v64, a->v, a->Lk, a->H, a->B, a->v_ss, a->v_sh, a->v_sb, 64, &hi)) return DVLA_ERR_CUDA;
  dim3 gkv((unsigned)((a->Lk + 127) / 128), (unsigned)a->H, (unsigned)a->B);
  attn_bwd_dkv_pipe_kernel<<<gkv, PIPE_THREADS, ATTN_PIPE_SMEM, s>>>(q64, k128, v128, do64, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("attn_bwd_dkv_pipe launch: %s", cudaGetErrorString(e)); return DVLA_ERR_CUDA; }
  count_launch();
  dim3 gq((unsigned)((a->Lq + 127) / 128), (unsigned)a->H, (unsigned)a->B);
  attn_bwd_dq_pipe_kernel<<<gq, PIPE_THREADS, ATTN_PIPE_SMEM, s>>>(q128, k64, v64, do128, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("attn_bwd_dq_pipe launch: %s", cudaGetErrorString(e)); return DVLA_ERR_CUDA; }
  count_launch();
  return DVLA_OK;
}

}  // namespace dvla
