// tcgen05 / TMEM flash attention FORWARD for head_dim = 64 (sm_100a).
//
// One CTA = 128 query rows of one (batch, head); thread t owns query row t (== TMEM lane t), so the online softmax needs
// no cross-thread reductions.  Per 128-key tile:
//   TMA (4-D tensor map over the strided [B, L, H, 64] view, 128B swizzle) -> smem K/V tiles, 2 stages
//   S = Q K^T        tcgen05.mma 128x128x16 (x4), both operands K-major, accumulator in TMEM cols [0,128)
//   softmax          tcgen05.ld S row -> mask bits / scale / running max & sum in registers -> P (bf16) written to smem in
//                    the K-major 128B-swizzled layout the next MMA's A descriptor expects (fence.proxy.async)
//   O_part = P V     tcgen05.mma 128x64x16 (x8), B = V tile read MN-major (hd contiguous), TMEM cols [128,192)
//   O = O*alpha + O_part in registers (tcgen05.ld)
// smem 112 KB and 256 TMEM columns per CTA => two CTAs per SM overlap one CTA's softmax with the other's MMAs.
// Same mask / dropout / LSE conventions as attention.cu (the mma.sync kernels, which still serve the backward).
#include "common.cuh"
#include "../../include/dvla.h"

namespace dvla {
void set_error(const char* fmt, ...);
void count_launch();

constexpr float TC_LOG2E = 1.4426950408889634f;
constexpr float TC_LN2 = 0.6931471805599453f;

struct AttnTcParams {
  bf16* out;
  float* lse;
  const uint32_t* mask;
  const uint8_t* tile_flags;   // 64x64-tile flags [nqt64, nkt64] (0 skip / 1 partial / 2 full) or null
  int B, H, Lq, Lk, nkt64, mask_words;
  long long o_sb, o_ss, o_sh;
  int q_head_inner, k_head_inner, v_head_inner;   // tensor-map dim order: (hd, H, L, B) if 1 else (hd, L, H, B)
  float scale;
  float drop_scale; uint32_t drop_thresh; uint64_t drop_seed; const uint64_t* drop_seed_ptr;
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void load_rows(void* dst, const CUtensorMap* m, uint64_t* bar, int head_inner, int row0, int h, int b) {
  if (head_inner) tma_load_4d(dst, m, bar, 0, h, row0, b);
  else            tma_load_4d(dst, m, bar, 0, row0, h, b);
}

constexpr int TQ = 128, TKV = 128;
constexpr int SM_Q = 0, SM_K = 16384, SM_V = SM_K + 2 * 16384, SM_P = SM_V + 2 * 16384, SM_BAR = SM_P + 32768;
constexpr int ATTN_TC_SMEM = SM_BAR + 128 + 1024;

__global__ void __launch_bounds__(128, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint64_t* bar_kv = bar_q + 1;     // [2]
  uint64_t* bar_s = bar_q + 3;
  uint64_t* bar_o = bar_q + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 5);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * TQ;
  const int row = q0 + tid;
  const int nkt = (p.Lk + TKV - 1) / TKV;

  if (tid == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(bar_q, 1); mbar_init(&bar_kv[0], 1); mbar_init(&bar_kv[1], 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;

  // 128x128 tile flag from the 64x64 flags: 0 skip, 1 partial, 2 full
  auto tile_flag = [&](int kt) -> int {
    if (!p.tile_flags) return 2;
    int any = 0, all = 1;
#pragma unroll
    for (int dq = 0; dq < 2; ++dq)
#pragma unroll
      for (int dk = 0; dk < 2; ++dk) {
        const int q64 = qt * 2 + dq, k64 = kt * 2 + dk;
        if (q64 * 64 >= p.Lq || k64 * 64 >= p.Lk) continue;
        const int f = p.tile_flags[static_cast<long long>(q64) * p.nkt64 + k64];
        any |= (f != 0);
        all &= (f == 2);
      }
    return any ? (all ? 2 : 1) : 0;
  };
  auto next_tile = [&](int kt) {
    while (kt < nkt && tile_flag(kt) == 0) ++kt;
    return kt;
  };

  int kt = next_tile(0);
  if (tid == 0) {
    mbar_expect_tx(bar_q, 16384);
    load_rows(smem + SM_Q, &tmQ, bar_q, p.q_head_inner, q0, h, b);
    if (kt < nkt) {
      mbar_expect_tx(&bar_kv[0], 32768);
      load_rows(smem + SM_K, &tmK, &bar_kv[0], p.k_head_inner, kt * TKV, h, b);
      load_rows(smem + SM_V, &tmV, &bar_kv[0], p.v_head_inner, kt * TKV, h, b);
    }
  }

  float o_acc[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) o_acc[j] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  const float sc = p.scale * TC_LOG2E;
  const long long bh = static_cast<long long>(b) * p.H + h;
  const uint64_t seed = p.drop_seed + ((p.drop_scale != 0.f && p.drop_seed_ptr) ? __ldg(p.drop_seed_ptr) : 0ull);
  const int nblk = (p.Lk + 7) >> 3;
  constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);
  constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);   // A = P K-major, B = V MN-major

  int stage = 0;
  uint32_t ph_kv[2] = {0, 0};
  uint32_t ph_s = 0, ph_o = 0;
  bool first = true;
  while (kt < nkt) {
    const int kt_next = next_tile(kt + 1);
    const int k0 = kt * TKV;
    if (tid == 0) {
      if (kt_next < nkt) {      // prefetch into the other stage (its previous reader, PV of tile j-1, has completed)
        const int s2 = stage ^ 1;
        mbar_expect_tx(&bar_kv[s2], 32768);
        load_rows(smem + SM_K + s2 * 16384, &tmK, &bar_kv[s2], p.k_head_inner, kt_next * TKV, h, b);
        load_rows(smem + SM_V + s2 * 16384, &tmV, &bar_kv[s2], p.v_head_inner, kt_next * TKV, h, b);
      }
      if (first) mbar_wait(bar_q, 0);
      mbar_wait(&bar_kv[stage], ph_kv[stage]);
      tc_fence_after();
      const uint32_t qa = smem_u32(smem + SM_Q), ka = smem_u32(smem + SM_K + stage * 16384);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_ss(tmem_S, make_smem_desc_sw128(qa + k * 32, 16, 1024), make_smem_desc_sw128(ka + k * 32, 16, 1024),
                     idesc_s, k != 0);
      umma_commit(bar_s);
    }
    ph_kv[stage] ^= 1;
    first = false;
    mbar_wait(bar_s, ph_s);
    ph_s ^= 1;
    tc_fence_after();

    // ---- mask words for this row / tile ----
    const int flag = tile_flag(kt);
    uint32_t w[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    if (flag == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int wi = kt * 4 + i;
        w[i] = (row < p.Lq && wi < p.mask_words) ? p.mask[static_cast<long long>(row) * p.mask_words + wi] : 0u;
      }
    }
    if (k0 + TKV > p.Lk) {   // tail keys
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kb = k0 + i * 32;
        const uint32_t valid = (kb >= p.Lk) ? 0u : ((p.Lk - kb >= 32) ? 0xffffffffu : ((1u << (p.Lk - kb)) - 1u));
        w[i] &= valid;
      }
    }
    // ---- pass 1: row max ----
    float tmax = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_S + lane_off + c * 32, r);
      tmem_ld_wait();
      const uint32_t wc = w[c];
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if ((wc >> j) & 1u) tmax = fmaxf(tmax, __uint_as_float(r[j]));
    }
    const float m_new = fmaxf(m_run, tmax * sc);
    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
    const float alpha = exp2f(m_run - m_use);
    m_run = m_new;
    l_run *= alpha;
    // ---- pass 2: p = exp2(s*sc - m), row sum, P (bf16) -> smem in the UMMA K-major SW128 layout ----
    uint8_t* prow = smem + SM_P + tid * 128;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_S + lane_off + c * 32, r);
      tmem_ld_wait();
      const uint32_t wc = w[c];
      float pv[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float e = ((wc >> j) & 1u) ? exp2f(fmaf(__uint_as_float(r[j]), sc, -m_use)) : 0.f;
        pv[j] = e;
        l_run += e;
      }
      if (p.drop_scale != 0.f) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t keep = dropout_keep8(seed, (bh * p.Lq + row) * nblk + ((k0 + c * 32) >> 3) + g, p.drop_thresh);
#pragma unroll
          for (int j = 0; j < 8; ++j) pv[g * 8 + j] = ((keep >> j) & 1u) ? pv[g * 8 + j] * p.drop_scale : 0.f;
        }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int chunk = c * 4 + g;              // 16-byte chunk index 0..15 along the 128 keys
        uint8_t* dst = prow + (chunk >> 3) * 16384 + (((chunk & 7) ^ (tid & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(pv[g * 8], pv[g * 8 + 1]), pack_bf16x2(pv[g * 8 + 2], pv[g * 8 + 3]),
                                                    pack_bf16x2(pv[g * 8 + 4], pv[g * 8 + 5]), pack_bf16x2(pv[g * 8 + 6], pv[g * 8 + 7]));
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t pa = smem_u32(smem + SM_P), va = smem_u32(smem + SM_V + stage * 16384);
#pragma unroll
      for (int k = 0; k < 8; ++k)   // 8 k-steps of 16 keys: P atom (k/4), 32 B inside the 128 B row; V rows 16k..16k+15
        umma_bf16_ss(tmem_O, make_smem_desc_sw128(pa + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                     make_smem_desc_sw128(va + k * (16 * 128), 8192, 1024), idesc_o, k != 0);
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, ph_o);
    ph_o ^= 1;
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_O + lane_off + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) o_acc[c * 32 + j] = fmaf(o_acc[c * 32 + j], alpha, __uint_as_float(r[j]));
    }
    tc_fence_before();
    kt = kt_next;
    stage ^= 1;
  }

  // ---- finalize ----
  const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
  if (row < p.Lq) {
    bf16* dst = p.out + b * p.o_sb + static_cast<long long>(row) * p.o_ss + h * p.o_sh;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      *reinterpret_cast<uint4*>(dst + c * 8) =
          make_uint4(pack_bf16x2(o_acc[8 * c] * inv, o_acc[8 * c + 1] * inv), pack_bf16x2(o_acc[8 * c + 2] * inv, o_acc[8 * c + 3] * inv),
                     pack_bf16x2(o_acc[8 * c + 4] * inv, o_acc[8 * c + 5] * inv), pack_bf16x2(o_acc[8 * c + 6] * inv, o_acc[8 * c + 7] * inv));
    if (p.lse) p.lse[bh * p.Lq + row] = (l_run > 0.f) ? (m_run + log2f(l_run)) * TC_LN2 : -INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------- host -------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn_shared();

// [B, L, H, 64] strided view -> 4-D map; dims ordered so that strides ascend (head vs seq), 128-row x 64-col box.
static bool make_tmap_rows(CUtensorMap* out, const void* base, long long L, long long H, long long B, long long ss,
                           long long sh, long long sb, int* head_inner) {
  EncodeTiledFn fn = get_encode_fn_shared();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not found"); return false; }
  *head_inner = (sh <= ss) ? 1 : 0;
  cuuint64_t dims[4]; cuuint64_t strides[3]; cuuint32_t box[4]; cuuint32_t estr[4] = {1, 1, 1, 1};
  dims[0] = 64; box[0] = 64;
  if (*head_inner) { dims[1] = H; dims[2] = L; strides[0] = sh * 2; strides[1] = ss * 2; box[1] = 1; box[2] = 128; }
  else             { dims[1] = L; dims[2] = H; strides[0] = ss * 2; strides[1] = sh * 2; box[1] = 128; box[2] = 1; }
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

// returns DVLA_OK, or DVLA_ERR_UNSUPPORTED when the strides cannot be expressed as a tensor map (caller falls back to
// the mma.sync forward kernel, which has no such restriction)
int attn_fwd_tc_dispatch(const dvla_attn_fwd_args* a, cudaStream_t s) {
  auto ok_strides = [](long long ss, long long sh, long long sb, long long B) {
    return ss > 0 && sh > 0 && (B == 1 || sb > 0) && ss % 8 == 0 && sh % 8 == 0 && sb % 8 == 0;
  };
  if (!ok_strides(a->q_ss, a->q_sh, a->q_sb, a->B) || !ok_strides(a->k_ss, a->k_sh, a->k_sb, a->B) ||
      !ok_strides(a->v_ss, a->v_sh, a->v_sb, a->B))
    return DVLA_ERR_UNSUPPORTED;
  AttnTcParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap tmQ, tmK, tmV;
  if (!make_tmap_rows(&tmQ, a->q, a->Lq, a->H, a->B, a->q_ss, a->q_sh, a->q_sb, &p.q_head_inner)) return DVLA_ERR_CUDA;
  if (!make_tmap_rows(&tmK, a->k, a->Lk, a->H, a->B, a->k_ss, a->k_sh, a->k_sb, &p.k_head_inner)) return DVLA_ERR_CUDA;
  if (!make_tmap_rows(&tmV, a->v, a->Lk, a->H, a->B, a->v_ss, a->v_sh, a->v_sb, &p.v_head_inner)) return DVLA_ERR_CUDA;
  p.out = (bf16*)a->o; p.lse = a->lse; p.mask = a->mask; p.tile_flags = a->mask ? a->tile_flags : nullptr;
  p.B = (int)a->B; p.H = (int)a->H; p.Lq = (int)a->Lq; p.Lk = (int)a->Lk;
  p.nkt64 = (p.Lk + 63) / 64; p.mask_words = a->mask_words;
  p.o_sb = a->o_sb; p.o_ss = a->o_ss; p.o_sh = a->o_sh;
  p.scale = a->scale;
  if (a->dropout_p > 0.f) {
    p.drop_thresh = (uint32_t)(a->dropout_p * 65536.0f + 0.5f);
    p.drop_scale = 1.0f / (1.0f - (float)p.drop_thresh / 65536.0f);
    p.drop_seed = a->dropout_seed;
    p.drop_seed_ptr = a->dropout_seed_ptr;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_TC_SMEM);
    if (e != cudaSuccess) { set_error("attn_fwd_tc smem attr: %s", cudaGetErrorString(e)); return DVLA_ERR_CUDA; }
    attr_set = true;
  }
  dim3 grid((unsigned)((a->Lq + TQ - 1) / TQ), (unsigned)a->H, (unsigned)a->B);
  attn_fwd_tc_kernel<<<grid, 128, ATTN_TC_SMEM, s>>>(tmQ, tmK, tmV, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("attn_fwd_tc launch: %s", cudaGetErrorString(e)); return DVLA_ERR_CUDA; }
  count_launch();
  return DVLA_OK;
}

}  // namespace dvla
