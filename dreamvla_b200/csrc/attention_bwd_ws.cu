// Warp-specialised tcgen05 / TMEM flash attention BACKWARD for head_dim = 64 (sm_100a).
//
// Same math, tiles and operand layouts as attention_bwd_tc.cu (two kernels, no atomics, no [Lq,Lk] tensor in HBM), but the
// roles are split so that the tensor pipe, TMA and the MUFU/FMA work of a CTA overlap instead of alternating:
//   warp 0      TMA producer (resident tiles once, streamed tiles through a ring; the dKV kernel's warp 0 also stages the
//               per-query lse / delta scalars of each streamed tile in smem)
//   warp 1      MMA issuer (one thread): the two "score" MMAs of tile j+1 (S, dP) are issued as soon as the compute
//               warps have pulled tile j out of TMEM, BEFORE the "gradient" MMAs of tile j, which wait for dS / P in smem
//   warps 2-9   compute: thread = one TMEM lane (row) x 32 of the 64 columns (warps w and w+4 share a lane quarter):
//               P = exp2(S*sc - lse), dS = P (dP - delta), mask bits / Philox dropout, bf16 -> swizzled smem A operand
// All hand-offs are mbarriers (tcgen05.commit for MMA completion); 2 CTAs per SM (<= 100 KB smem, 256 TMEM columns).
//   attn_bwd_dq_ws_kernel   CTA = 128 queries, loop over 64-key tiles:  dQ += dS K
//   attn_bwd_dkv_ws_kernel  CTA = 128 keys, loop over 64-query tiles:   dV += P^T dO, dK += dS^T Q
// Reference semantics: autograd of softmax(Q K^T * scale + mask) -> dropout -> @V (HF GPT2Attention._attn, timm Attention).
#include "attn_bwd_common.cuh"

namespace dvla {
namespace {

constexpr int BW_THREADS = 320;
constexpr int BW_MAX_T = 128;      // max streamed tiles per CTA (flag table)

// ================================================================ dQ ===============================================================
constexpr int QW_STAGES = 3;
constexpr int QW_SM_Q = 0, QW_SM_DO = 16384, QW_SM_KV = 32768, QW_SM_DS = QW_SM_KV + QW_STAGES * 16384,
              QW_SM_FLAGS = QW_SM_DS + 16384, QW_SM_BAR = QW_SM_FLAGS + BW_MAX_T;
constexpr int ATTN_DQ_WS_SMEM = QW_SM_BAR + 128;

__global__ void __launch_bounds__(BW_THREADS, 2)
attn_bwd_dq_ws_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                      const AttnBwdTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sflag = smem + QW_SM_FLAGS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + QW_SM_BAR);
  uint64_t* q_full = bars;            // 1
  uint64_t* kv_full = bars + 1;       // [3]
  uint64_t* kv_free = bars + 4;       // [3]
  uint64_t* sdp_full = bars + 7;
  uint64_t* sdp_free = bars + 8;      // 8 warps
  uint64_t* ds_ready = bars + 9;      // 8 warps
  uint64_t* dq_done = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = gridDim.z - 1 - blockIdx.z, h = blockIdx.x, b = blockIdx.y;   // heaviest q-tiles of ALL (b, h) first (see attention_fwd_ws.cu)
  const int q0 = qt * 128;
  const int nkt = (p.Lk + 63) / 64;
  const long long bh = static_cast<long long>(b) * p.H + h;

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) { printf("attn_bwd_dq_ws: dynamic smem base not 1024-aligned\n"); __trap(); }
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(q_full, 1);
    for (int s = 0; s < QW_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_free[s], 1); }
    mbar_init(sdp_full, 1); mbar_init(sdp_free, 8); mbar_init(ds_ready, 8); mbar_init(dq_done, 1);
    fence_barrier_init();
  }
  for (int ki = tid; ki < nkt; ki += BW_THREADS) {     // flag of this 128-query tile vs 64-key tile ki
    int f = 2;
    if (p.tile_flags) {
      int any = 0, all = 1;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const int q64 = qt * 2 + d;
        if (q64 * 64 >= p.Lq) continue;
        const int ff = p.tile_flags[static_cast<long long>(q64) * p.nkt64 + ki];
        any |= (ff != 0);
        all &= (ff == 2);
      }
      f = any ? (all ? 2 : 1) : 0;
    }
    if (f == 2 && (ki + 1) * 64 > p.Lk) f = 1;
    sflag[ki] = static_cast<uint8_t>(f);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tDP = tmem_base + 64, tDQ = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 32768);
      tma4(smem + QW_SM_Q, &tmQ, q_full, p.q_hi, q0, h, b);
      tma4(smem + QW_SM_DO, &tmDO, q_full, p.do_hi, q0, h, b);
      int idx = 0;
      for (int j = 0; j < nkt; ++j) {
        if (sflag[j] == 0) continue;
        const int st = idx % QW_STAGES;
        if (idx >= QW_STAGES) mbar_wait(&kv_free[st], ((idx / QW_STAGES) - 1) & 1);
        mbar_expect_tx(&kv_full[st], 16384);
        tma4(smem + QW_SM_KV + st * 16384, &tmK, &kv_full[st], p.k_hi, j * 64, h, b);
        tma4(smem + QW_SM_KV + st * 16384 + 8192, &tmV, &kv_full[st], p.v_hi, j * 64, h, b);
        ++idx;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);    // A = dS K-major, B = K tile MN-major
      const uint32_t qa = smem_u32(smem + QW_SM_Q), da = smem_u32(smem + QW_SM_DO), dsa = smem_u32(smem + QW_SM_DS);
      int cnt = 0, cg = 0, idx = 0, prev_st = -1;
      auto issue_dq = [&](int st) {
        const uint32_t ka = smem_u32(smem + QW_SM_KV + st * 16384);
        mbar_wait(ds_ready, cg & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tDQ, make_smem_desc_sw128(dsa + k * 32, 16, 1024), make_smem_desc_sw128(ka + k * 2048, 8192, 1024), idesc_o,
                       (cg | k) != 0 ? 1u : 0u);
        umma_commit(dq_done);
        umma_commit(&kv_free[st]);
        ++cg;
      };
      mbar_wait(q_full, 0);
      for (int j = 0; j < nkt; ++j) {
        if (sflag[j] == 0) continue;
        const int st = idx % QW_STAGES;
        mbar_wait(&kv_full[st], (idx / QW_STAGES) & 1);
        if (cnt > 0) mbar_wait(sdp_free, (cnt - 1) & 1);
        tc_fence_after();
        const uint32_t ka = smem_u32(smem + QW_SM_KV + st * 16384), va = ka + 8192;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tS, make_smem_desc_sw128(qa + k * 32, 16, 1024), make_smem_desc_sw128(ka + k * 32, 16, 1024), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tDP, make_smem_desc_sw128(da + k * 32, 16, 1024), make_smem_desc_sw128(va + k * 32, 16, 1024), idesc_s, k != 0);
        umma_commit(sdp_full);
        ++cnt;
        if (prev_st >= 0) issue_dq(prev_st);
        prev_st = st; ++idx;
      }
      if (prev_st >= 0) issue_dq(prev_st);
    }
  } else {
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int rowi = quarter * 32 + lane;
    const int row = q0 + rowi;
    const bool warp_active = (q0 + quarter * 32) < p.Lq;
    const uint32_t lane_off = (static_cast<uint32_t>(quarter * 32) << 16) + half * 32;
    const float sc = p.scale * B_LOG2E;
    const bool has_drop = p.drop_scale != 0.f;
    const uint32_t th16 = p.drop_thresh << 16;
    const uint64_t seed = p.drop_seed + ((has_drop && p.drop_seed_ptr) ? __ldg(p.drop_seed_ptr) : 0ull);
    const int nblk = (p.Lk + 7) >> 3;
    const float lse2 = (row < p.Lq) ? p.lse[bh * p.Lq + row] * B_LOG2E : INFINITY;
    const float dlt = (row < p.Lq) ? p.delta[bh * p.Lq + row] : 0.f;
    const uint32_t* mrow = p.mask ? p.mask + static_cast<long long>(row < p.Lq ? row : 0) * p.mask_words : nullptr;
    int c = 0;
    for (int j = 0; j < nkt; ++j) {
      const int f = sflag[j];
      if (f == 0) continue;
      const int k0 = j * 64 + half * 32;
      uint32_t wv = 0xffffffffu;
      if (f == 1) {
        if (mrow) wv = (2 * j + half < p.mask_words) ? __ldg(mrow + 2 * j + half) : 0u;
        const int nv = p.Lk - k0;
        if (nv < 32) wv &= (nv <= 0) ? 0u : ((1u << nv) - 1u);
      }
      mbar_wait(sdp_full, c & 1);
      tc_fence_after();
      uint32_t pk[16];
      if (!warp_active) {
        __syncwarp();
        if (lane == 0) mbar_arrive(sdp_free);
      } else {
        float sv[32], dpv[32];
        {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(tS + lane_off, r0);
          tmem_ld_32x32(tDP + lane_off, r1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) { sv[i] = __uint_as_float(r0[i]); dpv[i] = __uint_as_float(r1[i]); }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sdp_free);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (f == 1 && ((wv >> (8 * g)) & 0xffu) == 0u) {      // 8 hidden keys: dS = 0 without exp2 / Philox
#pragma unroll
            for (int i = 0; i < 4; ++i) pk[g * 4 + i] = 0u;
            continue;
          }
          uint4 rnd = make_uint4(0, 0, 0, 0);
          if (has_drop) rnd = philox4x32(seed, (bh * p.Lq + row) * nblk + (k0 >> 3) + g);
          float ds[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int cc = g * 8 + i;
            float pv = ex2_approx(fmaf(sv[cc], sc, -lse2));
            if (f == 1) pv = ((wv >> cc) & 1u) ? pv : 0.f;
            float dp = dpv[cc];
            if (has_drop) dp = dropout_keep_elem(rnd, th16, i) ? dp * p.drop_scale : 0.f;
            ds[i] = pv * (dp - dlt);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) pk[g * 4 + i] = pack_bf16x2(ds[2 * i], ds[2 * i + 1]);
        }
      }
      if (c > 0) { mbar_wait(dq_done, (c - 1) & 1); tc_fence_after(); }      // dS buffer free again
      if (warp_active) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(smem + QW_SM_DS + rowi * 128 + (((half * 4 + g) ^ (rowi & 7)) << 4)) =
              make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_ready);
      ++c;
    }
    if (c > 0) { mbar_wait(dq_done, (c - 1) & 1); tc_fence_after(); }
    if (warp_active) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;
      if (c > 0) tmem_ld32(tDQ + lane_off, v);
      if (row < p.Lq) store_row32(p.dq + b * p.dq_sb + static_cast<long long>(row) * p.dq_ss + h * p.dq_sh + half * 32, v, p.scale);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ============================================================== dK / dV ============================================================
constexpr int KW_STAGES = 2;
constexpr int KW_SM_K = 0, KW_SM_V = 16384, KW_SM_QD = 32768, KW_SM_P = KW_SM_QD + KW_STAGES * 16384, KW_SM_DS = KW_SM_P + 16384,
              KW_SM_LSE = KW_SM_DS + 16384, KW_SM_FLAGS = KW_SM_LSE + KW_STAGES * 2 * 64 * 4, KW_SM_BAR = KW_SM_FLAGS + BW_MAX_T;
constexpr int ATTN_DKV_WS_SMEM = KW_SM_BAR + 128;

__global__ void __launch_bounds__(BW_THREADS, 2)
attn_bwd_dkv_ws_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                       const AttnBwdTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* s_lse = reinterpret_cast<float*>(smem + KW_SM_LSE);     // [stage][64] lse*log2e, [stage][64] delta
  uint8_t* sflag = smem + KW_SM_FLAGS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + KW_SM_BAR);
  uint64_t* kv_full = bars;           // 1
  uint64_t* q_full = bars + 1;        // [2]
  uint64_t* q_free = bars + 3;        // [2]
  uint64_t* sdp_full = bars + 5;
  uint64_t* sdp_free = bars + 6;      // 8 warps
  uint64_t* pds_ready = bars + 7;     // 8 warps
  uint64_t* dvk_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kt = blockIdx.z, h = blockIdx.x, b = blockIdx.y;   // early key tiles (seen by every later frame) of ALL (b, h) first
  const int k0 = kt * 128;
  const int nqt = (p.Lq + 63) / 64;
  const long long bh = static_cast<long long>(b) * p.H + h;

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) { printf("attn_bwd_dkv_ws: dynamic smem base not 1024-aligned\n"); __trap(); }
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    for (int s = 0; s < KW_STAGES; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_free[s], 1); }
    mbar_init(sdp_full, 1); mbar_init(sdp_free, 8); mbar_init(pds_ready, 8); mbar_init(dvk_done, 1);
    fence_barrier_init();
  }
  for (int qi = tid; qi < nqt; qi += BW_THREADS) {     // flag of 64-query tile qi vs this 128-key tile
    int f = 2;
    if (p.tile_flags) {
      int any = 0, all = 1;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const int k64 = kt * 2 + d;
        if (k64 * 64 >= p.Lk) continue;
        const int ff = p.tile_flags[static_cast<long long>(qi) * p.nkt64 + k64];
        any |= (ff != 0);
        all &= (ff == 2);
      }
      f = any ? (all ? 2 : 1) : 0;
    }
    sflag[qi] = static_cast<uint8_t>(f);
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tDP = tmem_base + 64, tDV = tmem_base + 128, tDK = tmem_base + 192;

  if (warp == 0) {
    // TMA producer; all 32 lanes stage the per-query scalars of the tile before lane 0 arms the barrier for its Q / dO boxes
    if (lane == 0) {
      mbar_expect_tx(kv_full, 32768);
      tma4(smem + KW_SM_K, &tmK, kv_full, p.k_hi, k0, h, b);
      tma4(smem + KW_SM_V, &tmV, kv_full, p.v_hi, k0, h, b);
    }
    int idx = 0;
    for (int i = 0; i < nqt; ++i) {
      if (sflag[i] == 0) continue;
      const int st = idx % KW_STAGES;
      if (idx >= KW_STAGES) mbar_wait(&q_free[st], ((idx / KW_STAGES) - 1) & 1);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int qi = i * 64 + r * 32 + lane;
        s_lse[st * 128 + r * 32 + lane] = (qi < p.Lq) ? p.lse[bh * p.Lq + qi] * B_LOG2E : INFINITY;
        s_lse[st * 128 + 64 + r * 32 + lane] = (qi < p.Lq) ? p.delta[bh * p.Lq + qi] : 0.f;
      }
      __syncwarp();
      if (lane == 0) {
        mbar_expect_tx(&q_full[st], 16384);
        tma4(smem + KW_SM_QD + st * 16384, &tmQ, &q_full[st], p.q_hi, i * 64, h, b);
        tma4(smem + KW_SM_QD + st * 16384 + 8192, &tmDO, &q_full[st], p.do_hi, i * 64, h, b);
      }
      ++idx;
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, false, false);   // A = K / V rows, B = Q / dO rows, both K-major
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);    // A = P^T / dS^T K-major, B = dO / Q MN-major
      const uint32_t ka = smem_u32(smem + KW_SM_K), va = smem_u32(smem + KW_SM_V);
      const uint32_t pa = smem_u32(smem + KW_SM_P), dsa = smem_u32(smem + KW_SM_DS);
      int cnt = 0, cg = 0, idx = 0, prev_st = -1;
      auto issue_grad = [&](int st) {
        const uint32_t qa = smem_u32(smem + KW_SM_QD + st * 16384), da = qa + 8192;
        mbar_wait(pds_ready, cg & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)      // contraction over the 64 queries, 16 per step; B tile rows 16k.. read MN-major
          umma_bf16_ss(tDV, make_smem_desc_sw128(pa + k * 32, 16, 1024), make_smem_desc_sw128(da + k * 2048, 8192, 1024), idesc_o,
                       (cg | k) != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tDK, make_smem_desc_sw128(dsa + k * 32, 16, 1024), make_smem_desc_sw128(qa + k * 2048, 8192, 1024), idesc_o,
                       (cg | k) != 0 ? 1u : 0u);
        umma_commit(dvk_done);
        umma_commit(&q_free[st]);
        ++cg;
      };
      mbar_wait(kv_full, 0);
      for (int i = 0; i < nqt; ++i) {
        if (sflag[i] == 0) continue;
        const int st = idx % KW_STAGES;
        mbar_wait(&q_full[st], (idx / KW_STAGES) & 1);
        if (cnt > 0) mbar_wait(sdp_free, (cnt - 1) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(smem + KW_SM_QD + st * 16384), da = qa + 8192;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tS, make_smem_desc_sw128(ka + k * 32, 16, 1024), make_smem_desc_sw128(qa + k * 32, 16, 1024), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tDP, make_smem_desc_sw128(va + k * 32, 16, 1024), make_smem_desc_sw128(da + k * 32, 16, 1024), idesc_s, k != 0);
        umma_commit(sdp_full);
        ++cnt;
        if (prev_st >= 0) issue_grad(prev_st);
        prev_st = st; ++idx;
      }
      if (prev_st >= 0) issue_grad(prev_st);
    }
  } else {
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int rowi = quarter * 32 + lane;
    const int key = k0 + rowi;
    const bool warp_active = (k0 + quarter * 32) < p.Lk;
    const uint32_t lane_off = (static_cast<uint32_t>(quarter * 32) << 16) + half * 32;
    const float sc = p.scale * B_LOG2E;
    const bool has_drop = p.drop_scale != 0.f;
    const uint64_t seed = p.drop_seed + ((has_drop && p.drop_seed_ptr) ? __ldg(p.drop_seed_ptr) : 0ull);
    const int nblk = (p.Lk + 7) >> 3;
    const uint32_t* mrow = p.mask_t ? p.mask_t + static_cast<long long>(key < p.Lk ? key : 0) * p.mask_t_words : nullptr;
    int c = 0;
    for (int i = 0; i < nqt; ++i) {
      const int f = sflag[i];
      if (f == 0) continue;
      const int st = c % KW_STAGES;
      const int q0 = i * 64 + half * 32;
      uint32_t wv = 0xffffffffu;
      if (f == 1) {
        if (mrow && p.tile_flags) wv = (2 * i + half < p.mask_t_words) ? __ldg(mrow + 2 * i + half) : 0u;
        if (key >= p.Lk) wv = 0u;
      }
      mbar_wait(sdp_full, c & 1);
      tc_fence_after();
      uint32_t pkp[16], pkd[16];
      if (!warp_active) {
        __syncwarp();
        if (lane == 0) mbar_arrive(sdp_free);
      } else {
        float sv[32], dpv[32];
        {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(tS + lane_off, r0);
          tmem_ld_32x32(tDP + lane_off, r1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { sv[j] = __uint_as_float(r0[j]); dpv[j] = __uint_as_float(r1[j]); }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sdp_free);
        const float4* lse_t = reinterpret_cast<const float4*>(s_lse + st * 128 + half * 32);
        const float4* dl_t = reinterpret_cast<const float4*>(s_lse + st * 128 + 64 + half * 32);
        // groups of 8 queries that no key row of this warp can see (block masks: most of them) are skipped warp-wide
        uint32_t live = 0xfu;
        if (f == 1) {
          live = 0u;
#pragma unroll
          for (int g = 0; g < 4; ++g) live |= __any_sync(0xffffffffu, ((wv >> (8 * g)) & 0xffu) != 0u) ? (1u << g) : 0u;
        }
        uint32_t keepm[4] = {0, 0, 0, 0};           // dropout: lane (key%8 == r) owns queries c == r (mod 8)
        if (has_drop) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (!((live >> g) & 1u)) continue;
            const long long qc = q0 + (lane & 7) + g * 8;
            keepm[g] = dropout_keep8(seed, (bh * p.Lq + qc) * nblk + (key >> 3), p.drop_thresh);
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (!((live >> g) & 1u)) {
#pragma unroll
            for (int u = 0; u < 4; ++u) { pkp[g * 4 + u] = 0u; pkd[g * 4 + u] = 0u; }
            continue;
          }
          float pv[8], ds[8];
#pragma unroll
          for (int q4 = 0; q4 < 2; ++q4) {
            const float4 l4 = lse_t[g * 2 + q4], d4 = dl_t[g * 2 + q4];
            const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int cc = g * 8 + q4 * 4 + u;
              float e = ex2_approx(fmaf(sv[cc], sc, -ls[u]));       // lse = +inf for padded queries -> 0
              if (f == 1) e = ((wv >> cc) & 1u) ? e : 0.f;
              float dp = dpv[cc];
              if (has_drop) {
                const uint32_t m = __shfl_sync(0xffffffffu, keepm[g], (lane & ~7) | (cc & 7));
                const bool kp = (m >> (key & 7)) & 1u;
                dp = kp ? dp * p.drop_scale : 0.f;
                ds[q4 * 4 + u] = e * (dp - dl[u]);
                pv[q4 * 4 + u] = kp ? e * p.drop_scale : 0.f;
              } else {
                ds[q4 * 4 + u] = e * (dp - dl[u]);
                pv[q4 * 4 + u] = e;
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            pkp[g * 4 + u] = pack_bf16x2(pv[2 * u], pv[2 * u + 1]);
            pkd[g * 4 + u] = pack_bf16x2(ds[2 * u], ds[2 * u + 1]);
          }
        }
      }
      if (c > 0) { mbar_wait(dvk_done, (c - 1) & 1); tc_fence_after(); }      // P^T / dS^T buffers free again
      if (warp_active) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int off = rowi * 128 + (((half * 4 + g) ^ (rowi & 7)) << 4);
          *reinterpret_cast<uint4*>(smem + KW_SM_P + off) = make_uint4(pkp[g * 4], pkp[g * 4 + 1], pkp[g * 4 + 2], pkp[g * 4 + 3]);
          *reinterpret_cast<uint4*>(smem + KW_SM_DS + off) = make_uint4(pkd[g * 4], pkd[g * 4 + 1], pkd[g * 4 + 2], pkd[g * 4 + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_ready);
      ++c;
    }
    if (c > 0) { mbar_wait(dvk_done, (c - 1) & 1); tc_fence_after(); }
    if (warp_active) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
      if (c > 0) tmem_ld32(tDV + lane_off, v);
      if (key < p.Lk) store_row32(p.dv + b * p.dv_sb + static_cast<long long>(key) * p.dv_ss + h * p.dv_sh + half * 32, v, 1.0f);
      if (c > 0) tmem_ld32(tDK + lane_off, v);
      if (key < p.Lk) store_row32(p.dk + b * p.dk_sb + static_cast<long long>(key) * p.dk_ss + h * p.dk_sh + half * 32, v, p.scale);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace

bool make_attn_tmap_rows(CUtensorMap* out, const void* base, long long L, long long H, long long B, long long ss,
                         long long sh, long long sb, int box_rows, int* head_inner);   // attention_fwd_ws.cu

// DVLA_ERR_UNSUPPORTED -> caller uses another backward
int attn_bwd_ws_dispatch(const dvla_attn_bwd_args* a, const uint32_t* mask_t, int mask_t_words, cudaStream_t s, long long q_rows,
                         long long k_rows) {   // q_rows / k_rows: Lq / Lk, or a multiple of 128 below (the caller covers the rest)
  auto okst = [&](long long ss, long long sh, long long sb) {
    return ss > 0 && sh > 0 && (a->B == 1 || sb > 0) && ss % 8 == 0 && sh % 8 == 0 && sb % 8 == 0;
  };
  if (!okst(a->q_ss, a->q_sh, a->q_sb) || !okst(a->k_ss, a->k_sh, a->k_sb) || !okst(a->v_ss, a->v_sh, a->v_sb) ||
      !okst(a->do_ss, a->do_sh, a->do_sb))
    return DVLA_ERR_UNSUPPORTED;
  if (a->mask && !mask_t) return DVLA_ERR_UNSUPPORTED;
  if ((a->Lk + 63) / 64 > BW_MAX_T || (a->Lq + 63) / 64 > BW_MAX_T || a->H > 65535 || a->B > 65535) return DVLA_ERR_UNSUPPORTED;
  AttnBwdTcParams p;
  memset(&p, 0, sizeof(p));
  p.dq = (bf16*)a->dq; p.dk = (bf16*)a->dk; p.dv = (bf16*)a->dv; p.lse = a->lse; p.delta = a->delta;
  p.mask = a->mask; p.mask_t = a->mask ? mask_t : nullptr; p.tile_flags = a->mask ? a->tile_flags : nullptr;
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
  static const bool attr_ok = [] {               // once, race-free (C++11 static initialisation)
    cudaError_t e1 = cudaFuncSetAttribute(attn_bwd_dkv_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_DKV_WS_SMEM);
    cudaError_t e2 = cudaFuncSetAttribute(attn_bwd_dq_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_DQ_WS_SMEM);
    if (e1 != cudaSuccess || e2 != cudaSuccess) return false;
    cudaFuncSetAttribute(attn_bwd_dkv_ws_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);   // 2 CTAs / SM
    cudaFuncSetAttribute(attn_bwd_dq_ws_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    return true;
  }();
  if (!attr_ok) { set_error("attn_bwd_ws smem attr failed"); return DVLA_ERR_CUDA; }
  CUtensorMap q64, do64, k128, v128, q128, do128, k64, v64;
  int hi;
  if (!make_attn_tmap_rows(&q64, a->q, a->Lq, a->H, a->B, a->q_ss, a->q_sh, a->q_sb, 64, &p.q_hi)) return DVLA_ERR_CUDA;
  if (!make_attn_tmap_rows(&do64, a->d_o, a->Lq, a->H, a->B, a->do_ss, a->do_sh, a->do_sb, 64, &p.do_hi)) return DVLA_ERR_CUDA;
  if (!make_attn_tmap_rows(&k128, a->k, a->Lk, a->H, a->B, a->k_ss, a->k_sh, a->k_sb, 128, &p.k_hi)) return DVLA_ERR_CUDA;
  if (!make_attn_tmap_rows(&v128, a->v, a->Lk, a->H, a->B, a->v_ss, a->v_sh, a->v_sb, 128, &p.v_hi)) return DVLA_ERR_CUDA;
  if (!make_attn_tmap_rows(&q128, a->q, a->Lq, a->H, a->B, a->q_ss, a->q_sh, a->q_sb, 128, &hi)) return DVLA_ERR_CUDA;
  if (!make_attn_tmap_rows(&do128, a->d_o, a->Lq, a->H, a->B, a->do_ss, a->do_sh, a->do_sb, 128, &hi)) return DVLA_ERR_CUDA;
  if (!make_attn_tmap_rows(&k64, a->k, a->Lk, a->H, a->B, a->k_ss, a->k_sh, a->k_sb, 64, &hi)) return DVLA_ERR_CUDA;
  if (!make_attn_tmap_rows(&v64, a->v, a->Lk, a->H, a->B, a->v_ss, a->v_sh, a->v_sb, 64, &hi)) return DVLA_ERR_CUDA;
  dim3 gkv((unsigned)a->H, (unsigned)a->B, (unsigned)((k_rows + 127) / 128));
  attn_bwd_dkv_ws_kernel<<<gkv, BW_THREADS, ATTN_DKV_WS_SMEM, s>>>(q64, k128, v128, do64, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("attn_bwd_dkv_ws launch: %s", cudaGetErrorString(e)); return DVLA_ERR_CUDA; }
  count_launch();
  dim3 gq((unsigned)a->H, (unsigned)a->B, (unsigned)((q_rows + 127) / 128));
  attn_bwd_dq_ws_kernel<<<gq, BW_THREADS, ATTN_DQ_WS_SMEM, s>>>(q128, k64, v64, do128, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("attn_bwd_dq_ws launch: %s", cudaGetErrorString(e)); return DVLA_ERR_CUDA; }
  count_launch();
  return DVLA_OK;
}

}  // namespace dvla
