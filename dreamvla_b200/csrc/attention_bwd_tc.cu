// tcgen05 / TMEM flash attention BACKWARD for head_dim = 64 (sm_100a).  Two kernels, no atomics, no [Lq,Lk] tensor in HBM:
//
//   attn_bwd_dkv_tc_kernel  CTA = 128 keys of one (b,h); thread t owns key row t.  Loop over 64-query tiles:
//        S^T  = K Q^T, dP^T = V dO^T          tcgen05.mma 128x64x16 (x4 each)            -> TMEM [0,64), [64,128)
//        P^T  = exp2(S^T*sc - lse[q]),  dS^T = P^T (dP^T - delta[q])   (registers; dropout / mask bits applied)
//        P^T, dS^T (bf16) -> smem in the K-major 128B-swizzled A-operand layout
//        dV  += P^T dO,  dK += dS^T Q         tcgen05.mma 128x64x16 (x4 each), B = dO / Q tile read MN-major -> TMEM [128,256)
//   attn_bwd_dq_tc_kernel   CTA = 128 queries; thread t owns query row t.  Loop over 64-key tiles:
//        S = Q K^T, dP = dO V^T -> TMEM; dS -> smem; dQ += dS K (B = K tile MN-major) accumulated in TMEM
// K/V (resp. Q/dO) tiles stay resident in smem, the streamed operand is double-buffered by TMA; 96 / 80 KB smem and 256
// TMEM columns per CTA => two CTAs per SM.  delta = rowsum(dO*O) comes from attn_delta_kernel (attention.cu).
// Mask for the dKV kernel is read from a TRANSPOSED bit matrix [Lk, ceil(Lq/32)] so that a key row reads its 64 query bits
// as two words.  Dropout regenerates the forward's Philox mask: block = ((b*H+h)*Lq + q)*ceil(Lk/8) + key/8, bit key%8.
#include "attn_bwd_common.cuh"

namespace dvla {

// ============================================================ dK / dV ==============================================================
constexpr int KV_SM_K = 0, KV_SM_V = 16384, KV_SM_Q = 32768, KV_SM_DO = KV_SM_Q + 2 * 8192, KV_SM_P = KV_SM_DO + 2 * 8192,
              KV_SM_DS = KV_SM_P + 16384, KV_SM_LSE = KV_SM_DS + 16384, KV_SM_FLAGS = KV_SM_LSE + 2 * 2 * 64 * 4,
              KV_SM_BAR = KV_SM_FLAGS + 512;
constexpr int ATTN_DKV_SMEM = KV_SM_BAR + 128 + 1024;

__global__ void __launch_bounds__(256, 2)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                       const AttnBwdTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_lse = reinterpret_cast<float*>(smem + KV_SM_LSE);        // [2][64] lse*log2e, then [2][64] delta
  float* s_delta = s_lse + 2 * 64;
  uint64_t* bar_kv = reinterpret_cast<uint64_t*>(smem + KV_SM_BAR);
  uint64_t* bar_q = bar_kv + 1;   // [2]
  uint64_t* bar_s = bar_kv + 3;
  uint64_t* bar_o = bar_kv + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_kv + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rowi = tid & 127;          // key row of the tile owned by this thread (shared with thread tid ^ 128)
  const int half = tid >> 7;           // which 32 of the 64 query columns / output columns
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int k0 = kt * 128;
  const int key = k0 + rowi;
  const int nqt = (p.Lq + 63) / 64;
  const long long bh = static_cast<long long>(b) * p.H + h;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(bar_kv, 1); mbar_init(&bar_q[0], 1); mbar_init(&bar_q[1], 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_off = (static_cast<uint32_t>((warp & 3) * 32) << 16) + half * 32;   // TMEM lane quadrant + column half
  const uint32_t tS = tmem_base, tDP = tmem_base + 64, tDV = tmem_base + 128, tDK = tmem_base + 192;

  // per-(64-query tile) flag of this 128-key tile, computed once into smem: 0 skip, 1 partial, 2 full
  uint8_t* s_flags = smem + KV_SM_FLAGS;
  for (int qi = tid; qi < nqt && qi < 512; qi += 256) {
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
  __syncthreads();
  auto tile_flag = [&](int qt) -> int { return s_flags[qt]; };
  auto next_tile = [&](int qt) {
    while (qt < nqt && s_flags[qt] == 0) ++qt;
    return qt;
  };
  auto load_q = [&](int qt, int st) {        // thread 0: Q and dO tiles of 64 rows -> stage st
    mbar_expect_tx(&bar_q[st], 16384);
    tma4(smem + KV_SM_Q + st * 8192, &tmQ, &bar_q[st], p.q_hi, qt * 64, h, b);
    tma4(smem + KV_SM_DO + st * 8192, &tmDO, &bar_q[st], p.do_hi, qt * 64, h, b);
  };

  int qt = next_tile(0);
  if (tid == 0) {
    mbar_expect_tx(bar_kv, 32768);
    tma4(smem + KV_SM_K, &tmK, bar_kv, p.k_hi, k0, h, b);
    tma4(smem + KV_SM_V, &tmV, bar_kv, p.v_hi, k0, h, b);
    if (qt < nqt) load_q(qt, 0);
  }
  const float sc = p.scale * B_LOG2E;
  const uint64_t seed = p.drop_seed + ((p.drop_scale != 0.f && p.drop_seed_ptr) ? __ldg(p.drop_seed_ptr) : 0ull);
  const int nblk = (p.Lk + 7) >> 3;
  constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);   // A K-major (K/V rows), B K-major (Q/dO rows)
  constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);    // A = P^T/dS^T K-major, B = dO/Q MN-major

  int st = 0;
  uint32_t ph_q0 = 0, ph_q1 = 0, ph_s = 0, ph_o = 0;
  bool first = true, any_iter = false;
  // software prefetch of the per-tile scalars (lse, delta for tid < 64; transposed mask words for every key row): the
  // global loads of tile j+1 are issued at the top of iteration j and consumed at the top of iteration j+1
  float pf_lse = INFINITY, pf_delta = 0.f;
  uint32_t pf_w = 0xffffffffu;         // transposed mask word covering this thread's 32 queries of the next tile
  auto prefetch = [&](int qtile) {
    if (tid < 64) {
      const int qi = qtile * 64 + tid;
      pf_lse = (qi < p.Lq) ? p.lse[bh * p.Lq + qi] * B_LOG2E : INFINITY;
      pf_delta = (qi < p.Lq) ? p.delta[bh * p.Lq + qi] : 0.f;
    }
    pf_w = 0xffffffffu;
    if (s_flags[qtile] == 1) {
      const int wi = qtile * 2 + half;
      pf_w = (key < p.Lk && wi < p.mask_t_words) ? p.mask_t[static_cast<long long>(key) * p.mask_t_words + wi] : 0u;
    }
    if (key >= p.Lk) pf_w = 0u;
  };
  if (qt < nqt) prefetch(qt);
  while (qt < nqt) {
    const int qt_next = next_tile(qt + 1);
    const int q0 = qt * 64;
    if (tid < 64) {                                  // per-query scalars of this tile -> smem (read by every key row)
      s_lse[st * 64 + tid] = pf_lse;
      s_delta[st * 64 + tid] = pf_delta;
    }
    const uint32_t wv = pf_w;
    if (qt_next < nqt) prefetch(qt_next);
    if (tid == 0) {
      if (qt_next < nqt) load_q(qt_next, st ^ 1);
      if (first) mbar_wait(bar_kv, 0);
      mbar_wait(&bar_q[st], st ? ph_q1 : ph_q0);
      tc_fence_after();
      const uint32_t ka = smem_u32(smem + KV_SM_K), va = smem_u32(smem + KV_SM_V);
      const uint32_t qa = smem_u32(smem + KV_SM_Q + st * 8192), da = smem_u32(smem + KV_SM_DO + st * 8192);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_ss(tS, make_smem_desc_sw128(ka + k * 32, 16, 1024), make_smem_desc_sw128(qa + k * 32, 16, 1024), idesc_s, k != 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_ss(tDP, make_smem_desc_sw128(va + k * 32, 16, 1024), make_smem_desc_sw128(da + k * 32, 16, 1024), idesc_s, k != 0);
      umma_commit(bar_s);
    }
    if (st) ph_q1 ^= 1; else ph_q0 ^= 1;
    first = false;
    __syncthreads();                                 // s_lse / s_delta visible
    mbar_wait(bar_s, ph_s);
    ph_s ^= 1;
    tc_fence_after();

    // visibility of (query c, this key) for this thread's 32 queries: bits of wv (transposed bit matrix)
    float sv[32], dpv[32];
    tmem_ld32(tS + lane_off, sv);
    tmem_ld32(tDP + lane_off, dpv);
    const float* lse_t = s_lse + st * 64 + half * 32;
    const float* dl_t = s_delta + st * 64 + half * 32;
    uint32_t keepm[4];                               // dropout: lane (key%8 == r) owns queries c == r (mod 8)
    if (p.drop_scale != 0.f) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long qc = q0 + half * 32 + (lane & 7) + i * 8;
        keepm[i] = dropout_keep8(seed, (bh * p.Lq + qc) * nblk + (key >> 3), p.drop_thresh);
      }
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const bool vis = (wv >> c) & 1u;
      const float pv = vis ? ex2_approx(fmaf(sv[c], sc, -lse_t[c])) : 0.f;       // lse = +inf for padded queries -> 0
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
    write_row32(smem + KV_SM_P, rowi, half, sv);     // P^T  (dropped) -> A operand for dV
    write_row32(smem + KV_SM_DS, rowi, half, dpv);   // dS^T           -> A operand for dK
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t pa = smem_u32(smem + KV_SM_P), dsa = smem_u32(smem + KV_SM_DS);
      const uint32_t qa = smem_u32(smem + KV_SM_Q + st * 8192), da = smem_u32(smem + KV_SM_DO + st * 8192);
#pragma unroll
      for (int k = 0; k < 4; ++k)      // contraction over the 64 queries, 16 per step; B tile rows 16k.. read MN-major
        umma_bf16_ss(tDV, make_smem_desc_sw128(pa + k * 32, 16, 1024), make_smem_desc_sw128(da + k * 2048, 8192, 1024), idesc_o,
                     (any_iter || k != 0) ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_ss(tDK, make_smem_desc_sw128(dsa + k * 32, 16, 1024), make_smem_desc_sw128(qa + k * 2048, 8192, 1024), idesc_o,
                     (any_iter || k != 0) ? 1u : 0u);
      umma_commit(bar_o);
    }
    any_iter = true;
    mbar_wait(bar_o, ph_o);                          // P/dS smem, S/dP TMEM and this Q/dO stage are free again
    ph_o ^= 1;
    tc_fence_after();
    qt = qt_next;
    st ^= 1;
  }

  {   // TMEM loads are warp-collective: every lane executes them, only in-range key rows store (32 of 64 columns each)
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
    if (any_iter) tmem_ld32(tDV + lane_off, v);
    if (key < p.Lk) store_row32(p.dv + b * p.dv_sb + static_cast<long long>(key) * p.dv_ss + h * p.dv_sh + half * 32, v, 1.0f);
    if (any_iter) tmem_ld32(tDK + lane_off, v);
    if (key < p.Lk) store_row32(p.dk + b * p.dk_sb + static_cast<long long>(key) * p.dk_ss + h * p.dk_sh + half * 32, v, p.scale);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ============================================================== dQ =================================================================
constexpr int DQ_SM_Q = 0, DQ_SM_DO = 16384, DQ_SM_K = 32768, DQ_SM_V = DQ_SM_K + 2 * 8192, DQ_SM_DS = DQ_SM_V + 2 * 8192,
              DQ_SM_FLAGS = DQ_SM_DS + 16384, DQ_SM_BAR = DQ_SM_FLAGS + 512;
constexpr int ATTN_DQ_SMEM = DQ_SM_BAR + 128 + 1024;

__global__ void __launch_bounds__(256, 2)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                      const AttnBwdTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + DQ_SM_BAR);
  uint64_t* bar_kv = bar_q + 1;   // [2]
  uint64_t* bar_s = bar_q + 3;
  uint64_t* bar_o = bar_q + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 5);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int rowi = tid & 127, half = tid >> 7;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * 128;
  const int row = q0 + rowi;
  const int nkt = (p.Lk + 63) / 64;
  const long long bh = static_cast<long long>(b) * p.H + h;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(bar_q, 1); mbar_init(&bar_kv[0], 1); mbar_init(&bar_kv[1], 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_off = (static_cast<uint32_t>((warp & 3) * 32) << 16) + half * 32;
  const uint32_t tS = tmem_base, tDP = tmem_base + 64, tDQ = tmem_base + 128;

  uint8_t* s_flags = smem + DQ_SM_FLAGS;       // per-(64-key tile) flag of this 128-query tile
  for (int ki = tid; ki < nkt && ki < 512; ki += 256) {
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
  __syncthreads();
  auto next_tile = [&](int kt) {
    while (kt < nkt && s_flags[kt] == 0) ++kt;
    return kt;
  };
  auto load_kv = [&](int kt, int st) {
    mbar_expect_tx(&bar_kv[st], 16384);
    tma4(smem + DQ_SM_K + st * 8192, &tmK, &bar_kv[st], p.k_hi, kt * 64, h, b);
    tma4(smem + DQ_SM_V + st * 8192, &tmV, &bar_kv[st], p.v_hi, kt * 64, h, b);
  };

  int kt = next_tile(0);
  if (tid == 0) {
    mbar_expect_tx(bar_q, 32768);
    tma4(smem + DQ_SM_Q, &tmQ, bar_q, p.q_hi, q0, h, b);
    tma4(smem + DQ_SM_DO, &tmDO, bar_q, p.do_hi, q0, h, b);
    if (kt < nkt) load_kv(kt, 0);
  }
  const float sc = p.scale * B_LOG2E;
  const float lse2 = (row < p.Lq) ? p.lse[bh * p.Lq + row] * B_LOG2E : INFINITY;
  const float dlt = (row < p.Lq) ? p.delta[bh * p.Lq + row] : 0.f;
  const uint64_t seed = p.drop_seed + ((p.drop_scale != 0.f && p.drop_seed_ptr) ? __ldg(p.drop_seed_ptr) : 0ull);
  const int nblk = (p.Lk + 7) >> 3;
  constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
  constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);

  int st = 0;
  uint32_t ph_kv0 = 0, ph_kv1 = 0, ph_s = 0, ph_o = 0;
  bool first = true, any_iter = false;
  uint32_t pf_w = 0xffffffffu;     // mask word of this thread's 32 keys of the NEXT tile (software prefetch)
  auto prefetch = [&](int ktile) {
    pf_w = 0xffffffffu;
    if (s_flags[ktile] == 1) {
      const int wi = ktile * 2 + half;
      pf_w = (row < p.Lq && wi < p.mask_words) ? p.mask[static_cast<long long>(row) * p.mask_words + wi] : 0u;
    }
  };
  if (kt < nkt) prefetch(kt);
  while (kt < nkt) {
    const int kt_next = next_tile(kt + 1);
    const int k0 = kt * 64;
    uint32_t wv = pf_w;
    if (kt_next < nkt) prefetch(kt_next);
    if (tid == 0) {
      if (kt_next < nkt) load_kv(kt_next, st ^ 1);
      if (first) mbar_wait(bar_q, 0);
      mbar_wait(&bar_kv[st], st ? ph_kv1 : ph_kv0);
      tc_fence_after();
      const uint32_t qa = smem_u32(smem + DQ_SM_Q), da = smem_u32(smem + DQ_SM_DO);
      const uint32_t ka = smem_u32(smem + DQ_SM_K + st * 8192), va = smem_u32(smem + DQ_SM_V + st * 8192);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_ss(tS, make_smem_desc_sw128(qa + k * 32, 16, 1024), make_smem_desc_sw128(ka + k * 32, 16, 1024), idesc_s, k != 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_ss(tDP, make_smem_desc_sw128(da + k * 32, 16, 1024), make_smem_desc_sw128(va + k * 32, 16, 1024), idesc_s, k != 0);
      umma_commit(bar_s);
    }
    if (st) ph_kv1 ^= 1; else ph_kv0 ^= 1;
    first = false;
    mbar_wait(bar_s, ph_s);
    ph_s ^= 1;
    tc_fence_after();

    {
      const int nv = p.Lk - (k0 + half * 32);          // valid keys among this thread's 32
      if (nv < 32) wv &= (nv <= 0) ? 0u : ((1u << nv) - 1u);
    }
    float sv[32], dpv[32];
    tmem_ld32(tS + lane_off, sv);
    tmem_ld32(tDP + lane_off, dpv);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint32_t keep = 0xffu;
      if (p.drop_scale != 0.f) keep = dropout_keep8(seed, (bh * p.Lq + row) * nblk + ((k0 + half * 32) >> 3) + g, p.drop_thresh);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        const bool vis = (wv >> c) & 1u;
        const float pv = vis ? ex2_approx(fmaf(sv[c], sc, -lse2)) : 0.f;
        float dp = dpv[c];
        if (p.drop_scale != 0.f) dp = ((keep >> j) & 1u) ? dp * p.drop_scale : 0.f;
        sv[c] = pv * (dp - dlt);
      }
    }
    write_row32(smem + DQ_SM_DS, rowi, half, sv);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t dsa = smem_u32(smem + DQ_SM_DS), ka = smem_u32(smem + DQ_SM_K + st * 8192);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_ss(tDQ, make_smem_desc_sw128(dsa + k * 32, 16, 1024), make_smem_desc_sw128(ka + k * 2048, 8192, 1024), idesc_o,
                     (any_iter || k != 0) ? 1u : 0u);
      umma_commit(bar_o);
    }
    any_iter = true;
    mbar_wait(bar_o, ph_o);
    ph_o ^= 1;
    tc_fence_after();
    kt = kt_next;
    st ^= 1;
  }
  {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
    if (any_iter) tmem_ld32(tDQ + lane_off, v);
    if (row < p.Lq) store_row32(p.dq + b * p.dq_sb + static_cast<long long>(row) * p.dq_ss + h * p.dq_sh + half * 32, v, p.scale);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ============================================================== host ==============================================================
static bool make_tmap_rows_b(CUtensorMap* out, const void* base, long long L, long long H, long long B, long long ss, long long sh,
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
  if (r != CUDA_SUCCESS) { set_error("attn bwd tensor map encode failed (%d)", (int)r); return false; }
  return true;
}

// DVLA_ERR_UNSUPPORTED -> caller uses the mma.sync backward
int attn_bwd_tc_dispatch(const dvla_attn_bwd_args* a, const uint32_t* mask_t, int mask_t_words, cudaStream_t s) {
  auto okst = [&](long long ss, long long sh, long long sb) {
    return ss > 0 && sh > 0 && (a->B == 1 || sb > 0) && ss % 8 == 0 && sh % 8 == 0 && sb % 8 == 0;
  };
  if (!okst(a->q_ss, a->q_sh, a->q_sb) || !okst(a->k_ss, a->k_sh, a->k_sb) || !okst(a->v_ss, a->v_sh, a->v_sb) ||
      !okst(a->do_ss, a->do_sh, a->do_sb))
    return DVLA_ERR_UNSUPPORTED;
  if (a->mask && !mask_t) return DVLA_ERR_UNSUPPORTED;
  AttnBwdTcParams p;
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
    cudaError_t e1 = cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_DKV_SMEM);
    cudaError_t e2 = cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_DQ_SMEM);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("attn_bwd_tc smem attr failed"); return DVLA_ERR_CUDA; }
    attr_set = true;
  }
  CUtensorMap q64, do64, k128, v128, q128, do128, k64, v64;
  int hi;
  if (!make_tmap_rows_b(&q64, a->q, a->Lq, a->H, a->B, a->q_ss, a->q_sh, a->q_sb, 64, &p.q_hi)) return DVLA_ERR_CUDA;
  if (!make_tmap_rows_b(&do64, a->d_o, a->Lq, a->H, a->B, a->do_ss, a->do_sh, a->do_sb, 64, &p.do_hi)) return DVLA_ERR_CUDA;
  if (!make_tmap_rows_b(&k128, a->k, a->Lk, a->H, a->B, a->k_ss, a->k_sh, a->k_sb, 128, &p.k_hi)) return DVLA_ERR_CUDA;
  if (!make_tmap_rows_b(&v128, a->v, a->Lk, a->H, a->B, a->v_ss, a->v_sh, a->v_sb, 128, &p.v_hi)) return DVLA_ERR_CUDA;
  if (!make_tmap_rows_b(&q128, a->q, a->Lq, a->H, a->B, a->q_ss, a->q_sh, a->q_sb, 128, &hi)) return DVLA_ERR_CUDA;
  if (!make_tmap_rows_b(&do128, a->d_o, a->Lq, a->H, a->B, a->do_ss, a->do_sh, a->do_sb, 128, &hi)) return DVLA_ERR_CUDA;
  if (!make_tmap_rows_b(&k64, a->k, a->Lk, a->H, a->B, a->k_ss, a->k_sh, a->k_sb, 64, &hi)) return DVLA_ERR_CUDA;
  if (!make_tmap_rows_b(&v64, a->v, a->Lk, a->H, a->B, a->v_ss, a->v_sh, a->v_sb, 64, &hi)) return DVLA_ERR_CUDA;
  dim3 gkv((unsigned)((a->Lk + 127) / 128), (unsigned)a->H, (unsigned)a->B);
  attn_bwd_dkv_tc_kernel<<<gkv, 256, ATTN_DKV_SMEM, s>>>(q64, k128, v128, do64, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("attn_bwd_dkv_tc launch: %s", cudaGetErrorString(e)); return DVLA_ERR_CUDA; }
  count_launch();
  dim3 gq((unsigned)((a->Lq + 127) / 128), (unsigned)a->H, (unsigned)a->B);
  attn_bwd_dq_tc_kernel<<<gq, 256, ATTN_DQ_SMEM, s>>>(q128, k64, v64, do128, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("attn_bwd_dq_tc launch: %s", cudaGetErrorString(e)); return DVLA_ERR_CUDA; }
  count_launch();
  return DVLA_OK;
}

}  // namespace dvla
