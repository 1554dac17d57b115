// Warp-specialised tcgen05 / TMEM flash attention FORWARD for head_dim = 64 (sm_100a).
//
// One CTA = 256 query rows (two 128-row q-tiles) of one (batch, head); 10 warps:
//   warp 0      TMA producer: Q tiles once, then K/V tiles of 64 keys into a 3-stage ring (4-D tensor maps over the strided
//               [B, L, H, 64] views, 128B swizzle, OOB rows zero-filled)
//   warp 1      MMA issuer (one thread): S_t = Q_t K_j^T (128x64x16 x4) into TMEM, O_t += P_t V_j (128x64x16 x4);
//               S of tile j+1 is issued BEFORE the P.V of tile j so the tensor pipe works while the softmax warps run
//   warps 2-5   softmax of q-tile 0, warps 6-9 softmax of q-tile 1: thread = one S row (TMEM lane), S row (64 fp32) read
//               ONCE into registers, exact tile max, lazy reference max (O / l are rescaled only when the row max grows by
//               more than 2^8, FA4-style, so the common path never touches O), exp2 on the MUFU pipe, P (bf16) written
//               to smem in the K-major 128B-swizzled layout the P.V MMA's A descriptor expects
// Two CTAs per SM (112 KB smem, 256 TMEM columns, <= 102 registers): one CTA's fill / drain overlaps the other's steady
// state; the steady state is bound by the MUFU pipe (one exp2 per visible score).
// Mask / dropout / LSE conventions are those of attention.cu (bit matrix [Lq, ceil(Lk/32)], 64x64 tile flags,
// Philox block = ((b*H+h)*Lq+i)*ceil(Lk/8)+j/8).  Reference semantics: softmax(Q K^T * scale + mask) V with dropout on
// the probabilities (HF GPT2Attention._attn / timm Attention.forward).
#include "common.cuh"
#include "../../include/dvla.h"
#include <stdlib.h>
#include <string.h>

namespace dvla {
void set_error(const char* fmt, ...);
void count_launch();

namespace {

constexpr float WS_LOG2E = 1.4426950408889634f;
constexpr float WS_LN2 = 0.6931471805599453f;
constexpr int WS_THREADS = 320;
constexpr int WS_KV = 64;
constexpr int WS_STAGES = 3;
constexpr int WS_MAX_KT = 128;
constexpr int SM_Q = 0;                                   // 2 x 16 KB
constexpr int SM_KV = 32768;                              // 3 x (8 KB K + 8 KB V)
constexpr int SM_P = SM_KV + WS_STAGES * 16384;           // 2 x 16 KB
constexpr int SM_FLAG = SM_P + 32768;                     // 2 x 128 B
constexpr int SM_BAR = SM_FLAG + 2 * WS_MAX_KT;
constexpr int WS_SMEM = SM_BAR + 128;

struct AttnWsParams {
  bf16* out;
  float* lse;
  const uint32_t* mask;
  const uint8_t* tile_flags;   // 64x64-tile flags [nqt64, nkt64] (0 skip / 1 partial / 2 full) or null
  int B, H, Lq, Lk, nkt, mask_words;
  long long o_sb, o_ss, o_sh;
  int q_head_inner, k_head_inner, v_head_inner;
  float scale;
  float drop_scale; uint32_t drop_thresh; uint64_t drop_seed; const uint64_t* drop_seed_ptr;
  int trace;                   // DVLA_ATTN_TRACE=1: event timestamps (cycles since CTA start) overwrite the LSE rows of q-tile 0
};

__device__ __forceinline__ void tma_load_4d_ws(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void load_rows_ws(void* dst, const CUtensorMap* m, uint64_t* bar, int head_inner, int row0, int h, int b) {
  if (head_inner) tma_load_4d_ws(dst, m, bar, 0, h, row0, b);
  else            tma_load_4d_ws(dst, m, bar, 0, row0, h, b);
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(WS_THREADS, 2)
attn_fwd_ws_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnWsParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sflag = smem + SM_FLAG;                          // [2][WS_MAX_KT]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint64_t* q_full = bars;             // 1
  uint64_t* kv_full = bars + 1;        // [3]
  uint64_t* kv_free = bars + 4;        // [3]
  uint64_t* s_full = bars + 7;         // [2]
  uint64_t* s_free = bars + 9;         // [2]
  uint64_t* p_ready = bars + 11;       // [2]
  uint64_t* pv_done = bars + 13;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // grid = (H, B, q-blocks): CTAs are dispatched x-fastest, so EVERY (batch, head)'s heaviest q-block (the late rows of a
  // block-causal mask visit 4x the key tiles of the early ones) starts before any light one -- longest-job-first over the
  // whole grid instead of per (batch, head); a list-scheduling model of the C2 mask gives -15 % makespan
  const int qt = gridDim.z - 1 - blockIdx.z;
  const int h = blockIdx.x, b = blockIdx.y;
  const int q0 = qt * 256;
  const int nkt = p.nkt;
  const bool tile1 = q0 + 128 < p.Lq;

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) { printf("attn_fwd_ws: dynamic smem base not 1024-aligned\n"); __trap(); }
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < WS_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_free[s], 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1); mbar_init(&s_free[t], 4); mbar_init(&p_ready[t], 4); mbar_init(&pv_done[t], 1);
    }
    fence_barrier_init();
  }
  // per (q-tile, kv-tile) flag: 0 skip, 1 partial (mask words / key tail), 2 full
  for (int i = tid; i < 2 * nkt; i += WS_THREADS) {
    const int t = i / nkt, j = i - t * nkt;
    int f = 0;
    if (t == 0 || tile1) {
      if (!p.tile_flags) {
        f = 2;
      } else {
        int any = 0, all = 1;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const int q64 = qt * 4 + t * 2 + d;
          if (q64 * 64 >= p.Lq) continue;
          const int ff = p.tile_flags[static_cast<long long>(q64) * p.nkt + j];
          any |= (ff != 0);
          all &= (ff == 2);
        }
        f = any ? (all ? 2 : 1) : 0;
      }
      if (f == 2 && (j + 1) * WS_KV > p.Lk) f = 1;      // key tail
    }
    sflag[t * WS_MAX_KT + j] = static_cast<uint8_t>(f);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
#ifdef DVLA_ATTN_TRACE_BUILD      // build with -DDVLA_ATTN_TRACE_BUILD and run with DVLA_ATTN_TRACE=1 (tools/attn_trace.py)
  const long long t_start = clock64();
  float* trace_base = (p.trace && p.lse && qt == 0) ? p.lse + (static_cast<long long>(b) * p.H + h) * p.Lq : nullptr;
  auto trace = [&](int slot) {
    if (trace_base && slot < 256 && slot < p.Lq) trace_base[slot] = static_cast<float>(clock64() - t_start);
  };
#else
  constexpr float* trace_base = nullptr;
  auto trace = [](int) {};
#endif

  if (warp == 0) {
    // ------------------------------------------------ TMA producer ------------------------------------------------
    if (lane == 0) {
      mbar_expect_tx(q_full, tile1 ? 32768u : 16384u);
      load_rows_ws(smem + SM_Q, &tmQ, q_full, p.q_head_inner, q0, h, b);
      if (tile1) load_rows_ws(smem + SM_Q + 16384, &tmQ, q_full, p.q_head_inner, q0 + 128, h, b);
      int idx = 0;
      for (int j = 0; j < nkt; ++j) {
        if ((sflag[j] | sflag[WS_MAX_KT + j]) == 0) continue;
        const int st = idx % WS_STAGES;
        if (idx >= WS_STAGES) mbar_wait(&kv_free[st], ((idx / WS_STAGES) - 1) & 1);
        mbar_expect_tx(&kv_full[st], 16384u);
        load_rows_ws(smem + SM_KV + st * 16384, &tmK, &kv_full[st], p.k_head_inner, j * WS_KV, h, b);
        load_rows_ws(smem + SM_KV + st * 16384 + 8192, &tmV, &kv_full[st], p.v_head_inner, j * WS_KV, h, b);
        ++idx;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer --------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);     // A = P K-major, B = V MN-major
      int cs[2] = {0, 0}, cp[2] = {0, 0};
      int prev_j = -1, prev_st = 0, idx = 0;
      auto issue_pv = [&](int j, int st) {
        const uint32_t va = smem_u32(smem + SM_KV + st * 16384 + 8192);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (sflag[t * WS_MAX_KT + j] == 0) continue;
          mbar_wait(&p_ready[t], cp[t] & 1);
          tc_fence_after();
          trace(4 + cp[t] * 6 + t);
          const uint32_t pa = smem_u32(smem + SM_P + t * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tmem_base + 128 + t * 64, make_smem_desc_sw128(pa + k * 32, 16, 1024),
                         make_smem_desc_sw128(va + k * (16 * 128), 8192, 1024), idesc_o, (cp[t] | k) != 0 ? 1u : 0u);
          umma_commit(&pv_done[t]);
          ++cp[t];
        }
        umma_commit(&kv_free[st]);
      };
      mbar_wait(q_full, 0);
      trace(0);
      for (int j = 0; j < nkt; ++j) {
        if ((sflag[j] | sflag[WS_MAX_KT + j]) == 0) continue;
        const int st = idx % WS_STAGES;
        mbar_wait(&kv_full[st], (idx / WS_STAGES) & 1);
        tc_fence_after();
        trace(1 + idx * 6);
        const uint32_t ka = smem_u32(smem + SM_KV + st * 16384);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (sflag[t * WS_MAX_KT + j] == 0) continue;
          if (cs[t] > 0) { mbar_wait(&s_free[t], (cs[t] - 1) & 1); tc_fence_after(); }
          const uint32_t qa = smem_u32(smem + SM_Q + t * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tmem_base + t * 64, make_smem_desc_sw128(qa + k * 32, 16, 1024),
                         make_smem_desc_sw128(ka + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&s_full[t]);
          ++cs[t];
          trace(2 + idx * 6 + t);
        }
        if (prev_j >= 0) issue_pv(prev_j, prev_st);
        prev_j = j; prev_st = st; ++idx;
      }
      if (prev_j >= 0) issue_pv(prev_j, prev_st);
    }
  } else {
    // ------------------------------------------------ softmax -----------------------------------------------------
    const int t = (warp - 2) >> 2;
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int r_in = quarter * 32 + lane;               // row inside the q-tile
    const int row = q0 + t * 128 + r_in;
    const bool tile_active = (t == 0) || tile1;
    const bool warp_active = (q0 + t * 128 + quarter * 32) < p.Lq;
    if (tile_active) {
      const uint32_t t_s = tmem_base + t * 64 + (static_cast<uint32_t>(quarter * 32) << 16);
      const uint32_t t_o = tmem_base + 128 + t * 64 + (static_cast<uint32_t>(quarter * 32) << 16);
      uint8_t* prow = smem + SM_P + t * 16384 + r_in * 128;
      const float sc = p.scale * WS_LOG2E;
      const long long bh = static_cast<long long>(b) * p.H + h;
      const bool has_drop = p.drop_scale != 0.f;
      const uint32_t th16 = p.drop_thresh << 16;
      const uint64_t seed = p.drop_seed + ((has_drop && p.drop_seed_ptr) ? __ldg(p.drop_seed_ptr) : 0ull);
      const int nblk = (p.Lk + 7) >> 3;
      const uint32_t* mrow = p.mask ? p.mask + static_cast<long long>(row < p.Lq ? row : 0) * p.mask_words : nullptr;
      float m_ref = -INFINITY, l_run = 0.f;
      int c = 0;
      for (int j = 0; j < nkt; ++j) {
        const int f = sflag[t * WS_MAX_KT + j];
        if (f == 0) continue;
        const int k0 = j * WS_KV;
        // mask words of this row / tile (prefetched before the S wait)
        uint32_t w0 = 0xffffffffu, w1 = 0xffffffffu;
        if (f == 1) {
          if (mrow) {
            w0 = (2 * j < p.mask_words) ? __ldg(mrow + 2 * j) : 0u;
            w1 = (2 * j + 1 < p.mask_words) ? __ldg(mrow + 2 * j + 1) : 0u;
          }
          const int rem = p.Lk - k0;                   // > 0
          if (rem < 64) {
            w0 &= rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
            w1 &= rem <= 32 ? 0u : ((1u << (rem - 32)) - 1u);
          }
        }
        const bool tr = (quarter == 2) && lane == 0;
        const int tb = 64 + t * 64 + c * 10;
        if (tr) trace(tb);
        mbar_wait(&s_full[t], c & 1);
        if (tr) trace(tb + 1);
        tc_fence_after();
        if (tr) trace(tb + 2);
        float factor = 1.f;
        bool need = false;
        uint32_t pk[32];
        if (!warp_active) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[t]);
        } else {
          float s[64];
          {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(t_s, r0);
            tmem_ld_32x32(t_s + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) { s[i] = __uint_as_float(r0[i]); s[32 + i] = __uint_as_float(r1[i]); }
          }
          if (tr) trace(tb + 3);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[t]);
          if (tr) trace(tb + 4);
          float tmax = -INFINITY;
          if (f == 2) {
#pragma unroll
            for (int i = 0; i < 64; ++i) tmax = fmaxf(tmax, s[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              s[i] = ((w0 >> i) & 1u) ? s[i] : -INFINITY;
              s[32 + i] = ((w1 >> i) & 1u) ? s[32 + i] : -INFINITY;
              tmax = fmaxf(tmax, fmaxf(s[i], s[32 + i]));
            }
          }
          tmax *= sc;
          need = tmax > m_ref + 8.f;
          if (need) {
            factor = ex2_approx(m_ref - tmax);          // m_ref = -inf -> 0
            m_ref = tmax;
            l_run *= factor;
          }
          const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
          float lsum = 0.f;
          const uint64_t blk0 = (static_cast<uint64_t>(bh) * p.Lq + row) * nblk + (k0 >> 3);
#pragma unroll
          for (int g = 0; g < 8; ++g) {                  // 8 keys at a time: exp2, row sum, dropout, pack to bf16
            if (f == 1 && (((g < 4 ? w0 >> (8 * g) : w1 >> (8 * (g - 4))) & 0xffu) == 0u)) {
              // all 8 keys hidden from this row (block masks hide 60 % of the pairs inside partial tiles): no exp2, no Philox
#pragma unroll
              for (int i = 0; i < 4; ++i) pk[g * 4 + i] = 0u;
              continue;
            }
            float e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              e[i] = ex2_approx(fmaf(s[g * 8 + i], sc, -m_use));   // masked (-inf) -> 0
              lsum += e[i];
            }
            if (has_drop) {                 // dropped probabilities -> 0; the 1/(1-p) factor is applied once, to O (finalize)
              const uint4 rnd = philox4x32(seed, blk0 + g);
#pragma unroll
              for (int i = 0; i < 8; ++i) e[i] = dropout_keep_elem(rnd, th16, i) ? e[i] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) pk[g * 4 + i] = pack_bf16x2(e[2 * i], e[2 * i + 1]);
          }
          l_run += lsum;
        }
        if (tr) trace(tb + 5);
        if (c > 0) {
          mbar_wait(&pv_done[t], (c - 1) & 1);        // P buffer free, O_t quiescent
          if (tr) trace(tb + 6);
          tc_fence_after();
          if (tr) trace(tb + 7);
          if (__any_sync(0xffffffffu, need)) {          // rare: the row max grew by more than 2^8 -> rescale O
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t o[32];
              tmem_ld_32x32(t_o + hh * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
              tmem_st_32x32(t_o + hh * 32, o);
            }
            tmem_st_wait();
          }
        }
        if (warp_active) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<uint4*>(prow + ((g ^ (r_in & 7)) << 4)) =
                make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
        }
        if (tr) trace(tb + 8);
        fence_proxy_async_smem();
        if (tr) trace(tb + 9);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[t]);
        ++c;
      }
      // ---- finalize: O / l -> bf16, LSE ----
      if (c > 0) {
        mbar_wait(&pv_done[t], (c - 1) & 1);
        tc_fence_after();
      }
      if (quarter == 2 && lane == 0) trace(64 + t * 64 + 60);
      if (warp_active) {
        const float inv = l_run > 0.f ? (has_drop ? p.drop_scale : 1.0f) / l_run : 0.f;
        bf16* dst = p.out + b * p.o_sb + static_cast<long long>(row) * p.o_ss + h * p.o_sh;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t o[32];
          if (c > 0) {
            tmem_ld_32x32(t_o + hh * 32, o);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = 0u;
          }
          if (row < p.Lq) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              *reinterpret_cast<uint4*>(dst + hh * 32 + g * 8) =
                  make_uint4(pack_bf16x2(__uint_as_float(o[g * 8]) * inv, __uint_as_float(o[g * 8 + 1]) * inv),
                             pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv),
                             pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv),
                             pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv));
          }
        }
        if (quarter == 2 && lane == 0) trace(64 + t * 64 + 61);
        if (p.lse && row < p.Lq && !trace_base) p.lse[bh * p.Lq + row] = (l_run > 0.f) ? (m_ref + log2f(l_run)) * WS_LN2 : -INFINITY;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace

// ------------------------------------------------------- host -------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn_shared();

// [B, L, H, 64] strided view -> 4-D map; dims ordered so that strides ascend (head vs seq), box_rows x 64-col box.
bool make_attn_tmap_rows(CUtensorMap* out, const void* base, long long L, long long H, long long B, long long ss,
                         long long sh, long long sb, int box_rows, int* head_inner) {
  EncodeTiledFn fn = get_encode_fn_shared();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not found"); return false; }
  *head_inner = (sh <= ss) ? 1 : 0;
  cuuint64_t dims[4]; cuuint64_t strides[3]; cuuint32_t box[4]; cuuint32_t estr[4] = {1, 1, 1, 1};
  dims[0] = 64; box[0] = 64;
  if (*head_inner) { dims[1] = H; dims[2] = L; strides[0] = sh * 2; strides[1] = ss * 2; box[1] = 1; box[2] = box_rows; }
  else             { dims[1] = L; dims[2] = H; strides[0] = ss * 2; strides[1] = sh * 2; box[1] = box_rows; box[2] = 1; }
  dims[3] = B; strides[2] = (B > 1 ? sb : (long long)dims[1] * dims[2] * 64) * 2; box[3] = 1;
  if (strides[2] == 0) strides[2] = 16;
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("attn tensor map encode failed (%d): L=%lld H=%lld B=%lld ss=%lld sh=%lld sb=%lld", (int)r, L, H, B, ss, sh, sb);
    return false;
  }
  return true;
}

// returns DVLA_OK, or DVLA_ERR_UNSUPPORTED when the strides cannot be expressed as a tensor map or Lk is too long for the
// per-CTA flag table (the caller then takes the mma.sync forward kernel, which has no such restriction)
int attn_fwd_ws_dispatch(const dvla_attn_fwd_args* a, cudaStream_t s, long long q_rows) {
  auto ok_strides = [](long long ss, long long sh, long long sb, long long B) {
    return ss > 0 && sh > 0 && (B == 1 || sb > 0) && ss % 8 == 0 && sh % 8 == 0 && sb % 8 == 0;
  };
  if (!ok_strides(a->q_ss, a->q_sh, a->q_sb, a->B) || !ok_strides(a->k_ss, a->k_sh, a->k_sb, a->B) ||
      !ok_strides(a->v_ss, a->v_sh, a->v_sb, a->B))
    return DVLA_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
       reinterpret_cast<uintptr_t>(a->o)) & 15)
    return DVLA_ERR_UNSUPPORTED;
  if (a->o_ss % 8 || a->o_sh % 8 || a->o_sb % 8 || a->scale <= 0.f) return DVLA_ERR_UNSUPPORTED;
  const int nkt = (int)((a->Lk + WS_KV - 1) / WS_KV);
  if (nkt > WS_MAX_KT || a->H > 65535 || a->B > 65535) return DVLA_ERR_UNSUPPORTED;
  AttnWsParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap tmQ, tmK, tmV;
  if (!make_attn_tmap_rows(&tmQ, a->q, a->Lq, a->H, a->B, a->q_ss, a->q_sh, a->q_sb, 128, &p.q_head_inner)) return DVLA_ERR_CUDA;
  if (!make_attn_tmap_rows(&tmK, a->k, a->Lk, a->H, a->B, a->k_ss, a->k_sh, a->k_sb, WS_KV, &p.k_head_inner)) return DVLA_ERR_CUDA;
  if (!make_attn_tmap_rows(&tmV, a->v, a->Lk, a->H, a->B, a->v_ss, a->v_sh, a->v_sb, WS_KV, &p.v_head_inner)) return DVLA_ERR_CUDA;
  p.out = (bf16*)a->o; p.lse = a->lse; p.mask = a->mask; p.tile_flags = a->mask ? a->tile_flags : nullptr;
  p.B = (int)a->B; p.H = (int)a->H; p.Lq = (int)a->Lq; p.Lk = (int)a->Lk;
  p.nkt = nkt; p.mask_words = a->mask_words;
  p.o_sb = a->o_sb; p.o_ss = a->o_ss; p.o_sh = a->o_sh;
  p.scale = a->scale;
  { static const int tr = [] { const char* e = getenv("DVLA_ATTN_TRACE"); return (e && e[0] == '1') ? 1 : 0; }(); p.trace = tr; }
  if (a->dropout_p > 0.f) {
    p.drop_thresh = (uint32_t)(a->dropout_p * 65536.0f + 0.5f);
    p.drop_scale = 1.0f / (1.0f - (float)p.drop_thresh / 65536.0f);
    p.drop_seed = a->dropout_seed;
    p.drop_seed_ptr = a->dropout_seed_ptr;
  }
  static const cudaError_t attr_err = [] {      // once, race-free (C++11 static initialisation)
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM);
    if (e == cudaSuccess) cudaFuncSetAttribute(attn_fwd_ws_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);   // room for 2 CTAs / SM
    return e;
  }();
  if (attr_err != cudaSuccess) { set_error("attn_fwd_ws smem attr: %s", cudaGetErrorString(attr_err)); return DVLA_ERR_CUDA; }
  dim3 grid((unsigned)a->H, (unsigned)a->B, (unsigned)((q_rows + 255) / 256));   // q_rows: Lq, or a multiple of 256 below it
  attn_fwd_ws_kernel<<<grid, WS_THREADS, WS_SMEM, s>>>(tmQ, tmK, tmV, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("attn_fwd_ws launch: %s", cudaGetErrorString(e)); return DVLA_ERR_CUDA; }
  count_launch();
  return DVLA_OK;
}

}  // namespace dvla
