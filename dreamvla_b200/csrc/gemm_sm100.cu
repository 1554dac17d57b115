// Persistent, warp-specialised tcgen05 GEMM with fused epilogue for sm_100a.
//
//   out[M,N] = epi(alpha * A[M,K] * B[N,K]^T)          bf16 operands, fp32 accumulation in TMEM
//
// CTA = 320 threads, one CTA per SM (persistent over output tiles, static round-robin):
//   warp 0      TMA producer   (one lane): cp.async.bulk.tensor -> 128B-swizzled smem ring (4-6 stages), mbarrier tx
//   warp 1      MMA issuer     (one lane): tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16 per instruction,
//                                          tcgen05.commit -> frees smem stage / publishes accumulator
//   warps 2..9  epilogue       (8 warps) : tcgen05.ld 32x32b.x32 from TMEM -> bias/act/dact/dropout/residual -> global
// Two accumulator stages in TMEM (2*BN columns) let the epilogue of tile i overlap the mainloop of tile i+1.
// Both operands may be K-major or MN-major (UMMA descriptor + instruction-descriptor major bits), which covers
// forward, dgrad and wgrad for nn.Linear ([out,in]) and HF Conv1D ([in,out]) weights without any transpose pass.
//
// Replaces the cuBLAS calls behind nn.Linear / Conv1D in the reference (see include/dvla.h for call sites).
#include "common.cuh"
#include "../../include/dvla.h"

namespace dvla {

constexpr int BM = 128;
constexpr int BK = 64;       // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 16;   // 4 per TMEM lane quadrant: each SM sub-partition interleaves 4 epilogue warps
constexpr int EPI_STAGE_BYTES = 2048;  // per epilogue warp: 32 rows x 64 B staging panel
constexpr int GEMM_THREADS = 64 + NUM_EPI_WARPS * 32;

struct GemmParams {
  void* out;
  const bf16* bias;
  const void* residual;
  bf16* aux_out;
  const bf16* aux_in;
  int M, N, K;
  long long ldo, ldr, ld_aux;
  int num_m_tiles, num_n_tiles, num_k_blocks;
  int act, out_fp32, vec_ok, staged_ok;
  int k_splits, kb_per_split;   // split-K: work unit = (tile, split), each covering kb_per_split k-blocks
  int atomic_out;               // epilogue = out += v with red.global.add.bf16x2 (split-K accumulate into a gradient)
  // K-split TAIL (see map_unit): tiles [tail_first, total) are each cut into tail_splits k-ranges of tail_kbps k-blocks; the
  // partial accumulators meet in an fp32 workspace and the LAST unit to arrive runs the normal epilogue on the full sum
  int tail_first, tail_splits, tail_kbps;
  float* tail_ws;               // [CTA slot][16 warp regions][split][BN/16][32 lanes][4] fp32 partial accumulators
  unsigned* tail_cnt;           // [CTA slot][16] arrival counters, all zero between launches
  float alpha;
  float drop_scale;        // 1/(1-p_eff), 0 => dropout disabled
  uint32_t drop_thresh;    // round(p*65536)
  uint64_t drop_seed;
  const uint64_t* drop_seed_ptr;
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int BAR_BYTES = 256;
  static constexpr int STAGING_BYTES = NUM_EPI_WARPS * EPI_STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + BAR_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
};

// ---- epilogue for 8 consecutive columns of one row ---------------------------------------------------------------
__device__ __forceinline__ void epilogue8(float (&v)[8], int row, int col, const GemmParams& p) {
  const bool full = p.vec_ok && (col + 8 <= p.N);
  const int nvalid = min(8, p.N - col);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] *= p.alpha;
  if (p.bias) {
    if (full) {
      uint4 b = __ldg(reinterpret_cast<const uint4*>(p.bias + col));
      float2 f;
      f = unpack_bf16x2(b.x); v[0] += f.x; v[1] += f.y;
      f = unpack_bf16x2(b.y); v[2] += f.x; v[3] += f.y;
      f = unpack_bf16x2(b.z); v[4] += f.x; v[5] += f.y;
      f = unpack_bf16x2(b.w); v[6] += f.x; v[7] += f.y;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nvalid) v[j] += __bfloat162float(p.bias[col + j]);
    }
  }
  if (p.aux_out) {
    bf16* dst = p.aux_out + static_cast<long long>(row) * p.ld_aux + col;
    if (full) {
      uint4 o = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                           pack_bf16x2(v[6], v[7]));
      *reinterpret_cast<uint4*>(dst) = o;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nvalid) dst[j] = __float2bfloat16(v[j]);
    }
  }
  if (p.aux_in) {
    const bf16* src = p.aux_in + static_cast<long long>(row) * p.ld_aux + col;
    float x[8];
    if (full) {
      uint4 a = __ldg(reinterpret_cast<const uint4*>(src));
      float2 f;
      f = unpack_bf16x2(a.x); x[0] = f.x; x[1] = f.y;
      f = unpack_bf16x2(a.y); x[2] = f.x; x[3] = f.y;
      f = unpack_bf16x2(a.z); x[4] = f.x; x[5] = f.y;
      f = unpack_bf16x2(a.w); x[6] = f.x; x[7] = f.y;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = (j < nvalid) ? __bfloat162float(src[j]) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= act_bwd(x[j], p.act);
  } else if (p.act != ACT_NONE) {
    act_fwd_n<8>(v, p.act);
  }
  if (p.drop_scale != 0.f) {
    // element index = row*N + col; 8-element RNG blocks need col%8==0 alignment relative to row*N -> use (row*N+col)/8
    // only when N%8==0, otherwise fall back to per-row block indexing (row*ceil(N/8) + col/8).
    const uint64_t blk = static_cast<uint64_t>(row) * static_cast<uint64_t>((p.N + 7) >> 3) + (col >> 3);
    const uint64_t seed = p.drop_seed + (p.drop_seed_ptr ? __ldg(p.drop_seed_ptr) : 0ull);
    const uint32_t keep = dropout_keep8(seed, blk, p.drop_thresh);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ((keep >> j) & 1u) ? v[j] * p.drop_scale : 0.f;
  }
  if (p.out_fp32) {
    float* dst = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col;
    if (p.residual) {
      const float* r = reinterpret_cast<const float*>(p.residual) + static_cast<long long>(row) * p.ldr + col;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nvalid) v[j] += r[j];
    }
    if (full) {
      reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nvalid) dst[j] = v[j];
    }
  } else {
    bf16* dst = reinterpret_cast<bf16*>(p.out) + static_cast<long long>(row) * p.ldo + col;
    if (p.residual) {
      const bf16* r = reinterpret_cast<const bf16*>(p.residual) + static_cast<long long>(row) * p.ldr + col;
      if (full) {
        uint4 a = *reinterpret_cast<const uint4*>(r);
        float2 f;
        f = unpack_bf16x2(a.x); v[0] += f.x; v[1] += f.y;
        f = unpack_bf16x2(a.y); v[2] += f.x; v[3] += f.y;
        f = unpack_bf16x2(a.z); v[4] += f.x; v[5] += f.y;
        f = unpack_bf16x2(a.w); v[6] += f.x; v[7] += f.y;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < nvalid) v[j] += __bfloat162float(r[j]);
      }
    }
    if (full) {
      uint4 o = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                           pack_bf16x2(v[6], v[7]));
      *reinterpret_cast<uint4*>(dst) = o;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nvalid) dst[j] = __float2bfloat16(v[j]);
    }
  }
}

// ---- staged epilogue: per-warp 32 x 64 B smem panel (32 rows x 32 bf16 columns) ---------------------------------------
// The thread<->row TMEM layout gives each lane 32 contiguous bf16 of ONE row; writing those straight to global costs 32 L1
// wavefronts per store instruction.  Staging through this swizzled panel (16-byte chunk c of row r stored at
// c ^ ((r >> 1) & 3): conflict-free for both the row-wise writes and the 8-rows-per-instruction reads) turns every
// global access of the epilogue (output, pre-activation copy, residual) into full 64-byte row segments.
__device__ __forceinline__ uint4* stage_ptr(uint8_t* base, int row, int chunk) {
  return reinterpret_cast<uint4*>(base + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}
__device__ __forceinline__ void stage_write_row(uint8_t* base, int lane, const float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *stage_ptr(base, lane, c) = make_uint4(pack_bf16x2(v[8 * c], v[8 * c + 1]), pack_bf16x2(v[8 * c + 2], v[8 * c + 3]),
                                           pack_bf16x2(v[8 * c + 4], v[8 * c + 5]), pack_bf16x2(v[8 * c + 6], v[8 * c + 7]));
}
// panel (rows row0.., cols col0..col0+31) <-> global [*, ld] bf16; rows >= M and 8-col groups >= N are skipped
__device__ __forceinline__ void stage_flush(uint8_t* base, int lane, bf16* g, long long ld, int row0, int col0, int M, int N) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2), c = lane & 3;
    if (row0 + r < M && col0 + c * 8 < N)
      *reinterpret_cast<uint4*>(g + static_cast<long long>(row0 + r) * ld + col0 + c * 8) = *stage_ptr(base, r, c);
  }
}
// out += panel with one 16-byte vector reduction per 8 columns (REDG.E.ADD.BF16x8)
__device__ __forceinline__ void stage_flush_atomic(uint8_t* base, int lane, bf16* g, long long ld, int row0, int col0, int M, int N) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2), c = lane & 3;
    if (row0 + r < M && col0 + c * 8 < N) {
      const uint4 v = *stage_ptr(base, r, c);
      asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};"
                   :: "l"(g + static_cast<long long>(row0 + r) * ld + col0 + c * 8), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                   : "memory");
    }
  }
}
__device__ __forceinline__ void stage_load(uint8_t* base, int lane, const bf16* g, long long ld, int row0, int col0, int M, int N) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2), c = lane & 3;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (row0 + r < M && col0 + c * 8 < N)
      val = *reinterpret_cast<const uint4*>(g + static_cast<long long>(row0 + r) * ld + col0 + c * 8);
    *stage_ptr(base, r, c) = val;
  }
}

// one 32-column chunk of one warp: v[32] = this lane's row (row0 + lane), columns col0 .. col0+31
__device__ __forceinline__ void epilogue_chunk_staged(float (&v)[32], uint8_t* stage, int lane, int row0, int col0,
                                                      const GemmParams& p) {
  const int row = row0 + lane;
  if (p.alpha != 1.0f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
  }
  if (p.bias) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (col0 + c * 8 < p.N) {
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(p.bias + col0 + c * 8));
        float2 f;
        f = unpack_bf16x2(b.x); v[8 * c] += f.x; v[8 * c + 1] += f.y;
        f = unpack_bf16x2(b.y); v[8 * c + 2] += f.x; v[8 * c + 3] += f.y;
        f = unpack_bf16x2(b.z); v[8 * c + 4] += f.x; v[8 * c + 5] += f.y;
        f = unpack_bf16x2(b.w); v[8 * c + 6] += f.x; v[8 * c + 7] += f.y;
      }
    }
  }
  if (p.aux_out) {
    stage_write_row(stage, lane, v);
    __syncwarp();
    stage_flush(stage, lane, p.aux_out, p.ld_aux, row0, col0, p.M, p.N);
    __syncwarp();
  }
  if (p.aux_in) {
    stage_load(stage, lane, p.aux_in, p.ld_aux, row0, col0, p.M, p.N);
    __syncwarp();
    float x[32];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 a = *stage_ptr(stage, lane, c);
      float2 f;
      f = unpack_bf16x2(a.x); x[8 * c] = f.x; x[8 * c + 1] = f.y;
      f = unpack_bf16x2(a.y); x[8 * c + 2] = f.x; x[8 * c + 3] = f.y;
      f = unpack_bf16x2(a.z); x[8 * c + 4] = f.x; x[8 * c + 5] = f.y;
      f = unpack_bf16x2(a.w); x[8 * c + 6] = f.x; x[8 * c + 7] = f.y;
    }
    act_bwd_mul_n<32>(v, x, p.act);
    __syncwarp();
  } else if (p.act != ACT_NONE) {
    act_fwd_n<32>(v, p.act);
  }
  if (p.drop_scale != 0.f) {
    const uint64_t seed = p.drop_seed + (p.drop_seed_ptr ? __ldg(p.drop_seed_ptr) : 0ull);
    const uint64_t blk0 = static_cast<uint64_t>(row) * static_cast<uint64_t>((p.N + 7) >> 3) + (col0 >> 3);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t keep = dropout_keep8(seed, blk0 + c, p.drop_thresh);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[8 * c + j] = ((keep >> j) & 1u) ? v[8 * c + j] * p.drop_scale : 0.f;
    }
  }
  if (p.residual) {
    stage_load(stage, lane, reinterpret_cast<const bf16*>(p.residual), p.ldr, row0, col0, p.M, p.N);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 a = *stage_ptr(stage, lane, c);
      float2 f;
      f = unpack_bf16x2(a.x); v[8 * c] += f.x; v[8 * c + 1] += f.y;
      f = unpack_bf16x2(a.y); v[8 * c + 2] += f.x; v[8 * c + 3] += f.y;
      f = unpack_bf16x2(a.z); v[8 * c + 4] += f.x; v[8 * c + 5] += f.y;
      f = unpack_bf16x2(a.w); v[8 * c + 6] += f.x; v[8 * c + 7] += f.y;
    }
    __syncwarp();
  }
  stage_write_row(stage, lane, v);
  __syncwarp();
  if (p.atomic_out) stage_flush_atomic(stage, lane, reinterpret_cast<bf16*>(p.out), p.ldo, row0, col0, p.M, p.N);
  else              stage_flush(stage, lane, reinterpret_cast<bf16*>(p.out), p.ldo, row0, col0, p.M, p.N);
  __syncwarp();
}

// Epilogue of one output tile for one epilogue warp (shared by the single-CTA and CTA-pair kernels).
// Warp w (2..17): TMEM lane quadrant w & 3, column quarter (w - 2) >> 2; BN/4 columns in chunks of 32.
template <int BN, typename ArriveFn>
__device__ __forceinline__ void epilogue_tile(uint32_t t_acc /* tmem base + acc*BN */, uint8_t* staging, int warp, int lane,
                                              int m0, int n0, const GemmParams& p, ArriveFn arrive_tmem_free) {
  const int q = warp & 3;
  const int quarter = (warp - 2) >> 2;
  const uint32_t t_lane = t_acc + (static_cast<uint32_t>(q * 32) << 16);
  uint8_t* stage = staging + (warp - 2) * EPI_STAGE_BYTES;
  constexpr int CPW = BN / 4 / 32;
#pragma unroll 1
  for (int ci = 0; ci < CPW; ++ci) {
    const int c = quarter * (BN / 4) + ci * 32;
    uint32_t r[32];
    tmem_ld_32x32(t_lane + c, r);
    tmem_ld_wait();
    if (ci == CPW - 1) {   // all TMEM reads of this accumulator stage are done -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_tmem_free();
    }
    if (n0 + c >= p.N) continue;     // warp-uniform
    if (p.staged_ok) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      epilogue_chunk_staged(v, stage, lane, m0 + q * 32, n0 + c, p);
    } else {
      const int row = m0 + q * 32 + lane;
      if (row < p.M) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = n0 + c + g * 8;
          if (col < p.N) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
            epilogue8(v, row, col, p);
          }
        }
      }
    }
  }
}

// ---- work units ---------------------------------------------------------------------------------------------------
// A persistent CTA (pair) walks units round-robin.  Without a tail: unit = (tile, uniform k-split).  With a tail
// (p.tail_splits > 1): units [0, tail_first) are whole tiles -- a whole number of waves -- and the tiles that would form
// the last, partly filled wave are cut along K so that every CTA (pair) gets ~1/tail_splits of a tile instead of a
// few getting a whole one while the rest idle:  2.2 waves of tiles cost 2.2 tile-times instead of 3.
struct Unit { int tile, kb0, kb1, slot, split; };   // slot < 0: whole tile / atomic split-K; >= 0: tail tile index
__host__ __device__ __forceinline__ int gemm_total_units(int total_tiles, const GemmParams& p) {
  return p.tail_splits > 1 ? p.tail_first + (total_tiles - p.tail_first) * p.tail_splits : total_tiles * p.k_splits;
}
__host__ __device__ __forceinline__ Unit map_unit(int unit, int total_tiles, const GemmParams& p) {
  Unit u;
  if (p.tail_splits > 1) {
    u.split = 0;
    if (unit < p.tail_first) { u.tile = unit; u.kb0 = 0; u.kb1 = p.num_k_blocks; u.slot = -1; return u; }
    const int t = unit - p.tail_first, nt = total_tiles - p.tail_first;
    u.slot = t % nt;
    u.tile = p.tail_first + u.slot;
    u.split = t / nt;
    u.kb0 = u.split * p.tail_kbps;
    u.kb1 = u.kb0 + p.tail_kbps < p.num_k_blocks ? u.kb0 + p.tail_kbps : p.num_k_blocks;
    return u;
  }
  u.tile = unit % total_tiles;
  u.kb0 = (unit / total_tiles) * p.kb_per_split;
  u.kb1 = u.kb0 + p.kb_per_split < p.num_k_blocks ? u.kb0 + p.kb_per_split : p.num_k_blocks;
  u.slot = -1;
  u.split = 0;
  return u;
}

// Epilogue of one K-split tail unit for one epilogue warp.  The warp owns the same 32-row x BN/4-column region of the
// tile in every split, so the reduction is per warp region, with no CTA-wide synchronisation and no atomics on data:
//   1. TMEM -> registers -> this split's fp32 slice of the workspace, 16 bytes per lane per store ([col/4][lane][4]: a warp
//      instruction writes 512 contiguous bytes); the accumulator stage is handed back to the MMA warp once it is read;
//   2. __threadfence, one atomicAdd on the region's arrival counter;
//   3. the warp that arrives LAST sums the slices of all splits in split order (L2 reads, ld.cg) -- a fixed order, so the
//      result does not depend on which split finished last -- resets the counter, and runs the ordinary fused epilogue
//      (bias / act / dropout / residual / aux) on the sum: one rounding to bf16 whatever the number of splits.
// (fp32 red.global.add into one shared slice was measured first: ~20 us per GEMM for 17 MB of partials, the L2 atomic
//  units retire about one 4-byte element per clock per slice; plain vector stores + one read pass cost a third of that.)
template <int BN, typename ArriveFn>
__device__ __forceinline__ void epilogue_tile_tail(uint32_t t_acc, uint8_t* staging, int warp, int lane, int m0, int n0,
                                                   const GemmParams& p, int cta_slot, int split, ArriveFn arrive_tmem_free) {
  const int q = warp & 3;
  const int quarter = (warp - 2) >> 2;
  const uint32_t t_lane = t_acc + (static_cast<uint32_t>(q * 32) << 16);
  uint8_t* stage = staging + (warp - 2) * EPI_STAGE_BYTES;
  constexpr int CPW = BN / 4 / 32;
  constexpr int REGION = (BN / 4) * 32;                     // floats per warp region
  const long long region = static_cast<long long>(cta_slot) * NUM_EPI_WARPS + (warp - 2);
  float* base = p.tail_ws + region * p.tail_splits * REGION + lane * 4;     // + split * REGION + (col / 4) * 128
  float* mine = base + static_cast<long long>(split) * REGION;
#pragma unroll 1
  for (int ci = 0; ci < CPW; ++ci) {
    uint32_t r[32];
    tmem_ld_32x32(t_lane + quarter * (BN / 4) + ci * 32, r);
    tmem_ld_wait();
    if (ci == CPW - 1) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_tmem_free();
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      __stcg(reinterpret_cast<float4*>(mine + (ci * 8 + j) * 128),
             make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                         __uint_as_float(r[4 * j + 3])));
  }
  __threadfence();
  __syncwarp();
  unsigned arrived = 0;
  if (lane == 0) arrived = atomicAdd(p.tail_cnt + region, 1u);
  arrived = __shfl_sync(0xffffffffu, arrived, 0);
  if (arrived != static_cast<unsigned>(p.tail_splits - 1)) return;
  __threadfence();
  if (lane == 0) p.tail_cnt[region] = 0u;
#pragma unroll 1
  for (int ci = 0; ci < CPW; ++ci) {
    const int c = quarter * (BN / 4) + ci * 32;
    if (n0 + c >= p.N) continue;     // warp-uniform
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
#pragma unroll 1
    for (int sp = 0; sp < p.tail_splits; ++sp) {
      const float* src = base + static_cast<long long>(sp) * REGION + ci * 8 * 128;
      float4 t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = __ldcg(reinterpret_cast<const float4*>(src + j * 128));
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[4 * j] += t[j].x; v[4 * j + 1] += t[j].y; v[4 * j + 2] += t[j].z; v[4 * j + 3] += t[j].w; }
    }
    epilogue_chunk_staged(v, stage, lane, m0 + q * 32, n0 + c, p);
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::STAGES * Cfg::A_BYTES;
  uint8_t* staging = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.num_m_tiles * p.num_n_tiles;
  const int total_units = gemm_total_units(total_tiles, p);

  if (warp == 0) {
    // ------------------------------------------------ TMA producer ------------------------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        const Unit u = map_unit(unit, total_tiles, p);
        const int m0 = (u.tile % p.num_m_tiles) * BM;
        const int n0 = (u.tile / p.num_m_tiles) * BN;
        const int kb0 = u.kb0, kb1 = u.kb1;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
          uint8_t* a_dst = sA + s * Cfg::A_BYTES;
          uint8_t* b_dst = sB + s * Cfg::B_BYTES;
          if (!A_MN) {
            tma_load_2d(a_dst, &tmA, &full_bar[s], kb * BK, m0);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_2d(a_dst + i * (BK * 128), &tmA, &full_bar[s], m0 + i * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(b_dst, &tmB, &full_bar[s], kb * BK, n0);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_2d(b_dst + i * (BK * 128), &tmB, &full_bar[s], n0 + i * 64, kb * BK);
          }
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer --------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      int s = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const Unit u = map_unit(unit, total_tiles, p);
        const int kb0 = u.kb0, kb1 = u.kb1;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + s * Cfg::A_BYTES);
          const uint32_t b_base = smem_u32(sB + s * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t a_desc = A_MN ? make_smem_desc_sw128(a_base + k * (UMMA_K * 128), BK * 128, 1024)
                                         : make_smem_desc_sw128(a_base + k * (UMMA_K * 2), 16, 1024);
            const uint64_t b_desc = B_MN ? make_smem_desc_sw128(b_base + k * (UMMA_K * 128), BK * 128, 1024)
                                         : make_smem_desc_sw128(b_base + k * (UMMA_K * 2), 16, 1024);
            umma_bf16_ss(d_tmem, a_desc, b_desc, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_ph ^= 1;
      }
    }
  } else {
    // ------------------------------------------------ epilogue ----------------------------------------------------
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
      const Unit u = map_unit(unit, total_tiles, p);
      const int m0 = (u.tile % p.num_m_tiles) * BM;
      const int n0 = (u.tile / p.num_m_tiles) * BN;
      mbar_wait(&tfull_bar[acc], acc_ph);
      tc_fence_after();
      uint64_t* free_bar = &tempty_bar[acc];
      if (u.slot >= 0)
        epilogue_tile_tail<BN>(tmem_base + acc * BN, staging, warp, lane, m0, n0, p, u.slot, u.split, [free_bar] { mbar_arrive(free_bar); });
      else
        epilogue_tile<BN>(tmem_base + acc * BN, staging, warp, lane, m0, n0, p, [free_bar] { mbar_arrive(free_bar); });
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}


// =====================================================================================================================
// CTA-pair variant: tcgen05.mma.cta_group::2, 256 x 256 output tile per cluster of 2 CTAs.
// Each CTA stages ITS 128 rows of A and ITS 128 rows (N-half) of B per k-block (32 KB/stage instead of 48 KB for the same
// MMA work): L2->SM operand traffic per FLOP drops by a third, which is what bounds the single-CTA kernel on large GEMMs.
// Leader CTA (rank 0) issues all MMAs; its `full` barriers collect the TMA bytes of both CTAs; tcgen05.commit multicasts
// the stage-free / accumulator-ready arrivals to both CTAs; every epilogue warp of both CTAs arrives on the leader's
// accumulator-free barrier.
// =====================================================================================================================
struct Gemm2Cfg {
  static constexpr int BN = 256;
  static constexpr int A_BYTES = BM * BK * 2;        // this CTA's 128 rows of A
  static constexpr int B_BYTES = 128 * BK * 2;       // this CTA's 128 rows (N-half) of B
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 6;
  static constexpr int BAR_BYTES = 256;
  static constexpr int STAGING_BYTES = NUM_EPI_WARPS * EPI_STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + BAR_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
};

template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmParams p) {
  using Cfg = Gemm2Cfg;
  constexpr int BN = Cfg::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::STAGES * Cfg::A_BYTES;
  uint8_t* staging = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 2);            // leader's own arrive.expect_tx + the peer's remote arrive
      mbar_init(&empty_bar[s], 1);           // multicast tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);           // multicast tcgen05.commit
      mbar_init(&tempty_bar[s], 2 * NUM_EPI_WARPS);   // epilogue warps of both CTAs (used on the leader only)
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m2 = (p.M + 2 * BM - 1) / (2 * BM);
  const int total_tiles = num_m2 * p.num_n_tiles;
  const int total_units = gemm_total_units(total_tiles, p);
  const int num_clusters = gridDim.x >> 1;
  const int cluster_id = blockIdx.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int unit = cluster_id; unit < total_units; unit += num_clusters) {
        const Unit u = map_unit(unit, total_tiles, p);
        const int m0 = (u.tile % num_m2) * (2 * BM) + rank * BM;
        const int n0 = (u.tile / num_m2) * BN + rank * 128;
        const int kb0 = u.kb0, kb1 = u.kb1;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (leader) mbar_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);
          else        mbar_arrive_leader(&full_bar[s]);
          uint8_t* a_dst = sA + s * Cfg::A_BYTES;
          uint8_t* b_dst = sB + s * Cfg::B_BYTES;
          if (!A_MN) {
            tma_load_2d_2cta(a_dst, &tmA, &full_bar[s], kb * BK, m0);
          } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) tma_load_2d_2cta(a_dst + i * (BK * 128), &tmA, &full_bar[s], m0 + i * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d_2cta(b_dst, &tmB, &full_bar[s], kb * BK, n0);
          } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) tma_load_2d_2cta(b_dst + i * (BK * 128), &tmB, &full_bar[s], n0 + i * 64, kb * BK);
          }
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, A_MN, B_MN);
      int s = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      for (int unit = cluster_id; unit < total_units; unit += num_clusters) {
        mbar_wait(&tempty_bar[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const Unit u = map_unit(unit, total_tiles, p);
        const int kb0 = u.kb0, kb1 = u.kb1;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + s * Cfg::A_BYTES);
          const uint32_t b_base = smem_u32(sB + s * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t a_desc = A_MN ? make_smem_desc_sw128(a_base + k * (UMMA_K * 128), BK * 128, 1024)
                                         : make_smem_desc_sw128(a_base + k * (UMMA_K * 2), 16, 1024);
            const uint64_t b_desc = B_MN ? make_smem_desc_sw128(b_base + k * (UMMA_K * 128), BK * 128, 1024)
                                         : make_smem_desc_sw128(b_base + k * (UMMA_K * 2), 16, 1024);
            umma_bf16_ss_2cta(d_tmem, a_desc, b_desc, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
          }
          umma_commit_2cta(&empty_bar[s]);
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit_2cta(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_ph ^= 1;
      }
    }
  } else {
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int unit = cluster_id; unit < total_units; unit += num_clusters) {
      const Unit u = map_unit(unit, total_tiles, p);
      const int m0 = (u.tile % num_m2) * (2 * BM) + rank * BM;
      const int n0 = (u.tile / num_m2) * BN;
      mbar_wait(&tfull_bar[acc], acc_ph);
      tc_fence_after();
      uint64_t* free_bar = &tempty_bar[acc];
      if (u.slot >= 0)
        epilogue_tile_tail<BN>(tmem_base + acc * BN, staging, warp, lane, m0, n0, p, 2 * u.slot + static_cast<int>(rank), u.split,
                               [free_bar] { mbar_arrive_leader(free_bar); });
      else
        epilogue_tile<BN>(tmem_base + acc * BN, staging, warp, lane, m0, n0, p, [free_bar] { mbar_arrive_leader(free_bar); });
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 1) tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
}

// ---- SIMT kernels of identical semantics for operands TMA cannot address (unaligned strides, K or N < 8) -------------
// thread-per-(row, 8 columns) for short contractions; warp-per-(row, 8 columns) with a lane-strided k loop when the
// contraction is long and the output small (wgrad of the 7-wide DiT layers: 5 k outputs x K = 960).
__device__ __forceinline__ float simt_a(const bf16* a, long long lda, int a_mn, int row, int k) {
  return __bfloat162float(a_mn ? a[static_cast<long long>(k) * lda + row] : a[static_cast<long long>(row) * lda + k]);
}
__device__ __forceinline__ float simt_b(const bf16* b, long long ldb, int b_mn, int n, int k) {
  return __bfloat162float(b_mn ? b[static_cast<long long>(k) * ldb + n] : b[static_cast<long long>(n) * ldb + k]);
}
__global__ void gemm_simt_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, long long lda, long long ldb,
                                 int a_mn, int b_mn, GemmParams p) {
  const long long groups_n = (p.N + 7) / 8;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long long>(p.M) * groups_n) return;
  const int row = static_cast<int>(gid / groups_n);
  const int col = static_cast<int>(gid % groups_n) * 8;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  const int nvalid = min(8, p.N - col);
  for (int k = 0; k < p.K; ++k) {
    const float av = simt_a(a, lda, a_mn, row, k);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < nvalid) v[j] = fmaf(av, simt_b(b, ldb, b_mn, col + j, k), v[j]);
  }
  epilogue8(v, row, col, p);
}
__global__ void gemm_simt_warp_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, long long lda, long long ldb,
                                      int a_mn, int b_mn, GemmParams p) {
  const long long groups_n = (p.N + 7) / 8;
  const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= static_cast<long long>(p.M) * groups_n) return;      // warp-uniform
  const int row = static_cast<int>(wid / groups_n);
  const int col = static_cast<int>(wid % groups_n) * 8;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  const int nvalid = min(8, p.N - col);
  for (int k = lane; k < p.K; k += 32) {
    const float av = simt_a(a, lda, a_mn, row, k);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < nvalid) v[j] = fmaf(av, simt_b(b, ldb, b_mn, col + j, k), v[j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = warp_sum(v[j]);
  if (lane == 0) epilogue8(v, row, col, p);
}

}  // namespace dvla

// ============================================== host side ========================================================
#include <mutex>
#include <unordered_map>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace dvla {
void set_error(const char* fmt, ...);
void count_launch();
int num_sms();

// dvla_gemm_plan: the dispatcher runs as usual up to the point where it would encode tensor maps and launch; with this
// thread-local set it records its decisions instead (host only -- no CUDA call, usable without a GPU)
static thread_local dvla_gemm_plan_info* g_plan = nullptr;
static void record_plan(const GemmParams& p, int kernel, int tile_m, int tile_n, int m_tiles, int grid_ctas) {
  dvla_gemm_plan_info& o = *g_plan;
  o.kernel = kernel; o.tile_m = tile_m; o.tile_n = tile_n;
  o.m_tiles = m_tiles; o.n_tiles = p.num_n_tiles; o.k_blocks = p.num_k_blocks;
  o.k_splits = p.k_splits; o.kb_per_split = p.kb_per_split; o.atomic_out = p.atomic_out;
  o.tail_first = p.tail_first; o.tail_splits = p.tail_splits; o.tail_kbps = p.tail_kbps;
  o.units = kernel >= 2 ? gemm_total_units(m_tiles * p.num_n_tiles, p) : 0;
  o.grid_ctas = grid_ctas;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

EncodeTiledFn get_encode_fn_shared() { return get_encode_fn(); }

// Descriptor cache: the only mutable module-level state besides the launch counter.  Keyed by everything
// cuTensorMapEncodeTiled sees; guarded by a mutex because forward (main thread) and backward (autograd's device thread)
// call into the library concurrently.  Bounded: cleared when full (activations are recycled by the caching allocator, so
// a training loop touches a few thousand distinct (pointer, shape) pairs).
struct TmapKey {
  const void* base; uint64_t inner, outer, ld; uint32_t box_inner, box_outer;
  bool operator==(const TmapKey& o) const {
    return base == o.base && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner && box_outer == o.box_outer;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.base) * 0x9E3779B97F4A7C15ull;
    h ^= (k.inner + 0x632BE59BD9B4E019ull) * 0xD1342543DE82EF95ull; h = (h << 13) | (h >> 51);
    h ^= (k.outer << 20) ^ (k.ld * 0xA0761D6478BD642Full) ^ ((uint64_t)k.box_inner << 40) ^ ((uint64_t)k.box_outer << 52);
    return static_cast<size_t>(h ^ (h >> 29));
  }
};
static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
constexpr size_t TMAP_CACHE_MAX = 8192;

// 2-D bf16 tensor map: inner dim contiguous, 128B swizzle, zero OOB fill.
bool make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                       uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not found"); return false; }
  const TmapKey key{base, inner, outer, ld_elems, box_inner, box_outer};
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) { *out = it->second; return true; }
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu ld=%llu box=%ux%u base=%p", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld_elems, box_inner, box_outer,
              base);
    return false;
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmap_cache.size() >= TMAP_CACHE_MAX) g_tmap_cache.clear();
    g_tmap_cache.emplace(key, *out);
  }
  return true;
}

// cudaFuncSetAttribute once per kernel instantiation, race-free between the forward and the autograd thread
template <typename K>
static bool set_smem_attr_once(std::once_flag& once, cudaError_t& err, K kern, int bytes) {
  std::call_once(once, [&] { err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); });
  return err == cudaSuccess;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_tc(const dvla_gemm_args* a, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  if (g_plan) {
    const int units = gemm_total_units(p.num_m_tiles * p.num_n_tiles, p);
    record_plan(p, 2, BM, BN, p.num_m_tiles, units < num_sms() ? units : num_sms());
    return DVLA_OK;
  }
  CUtensorMap tmA, tmB;
  if (!A_MN) { if (!make_tmap_2d_bf16(&tmA, a->a, a->K, a->M, a->lda, BK, BM)) return DVLA_ERR_CUDA; }
  else       { if (!make_tmap_2d_bf16(&tmA, a->a, a->M, a->K, a->lda, 64, BK)) return DVLA_ERR_CUDA; }
  if (!B_MN) { if (!make_tmap_2d_bf16(&tmB, a->b, a->K, a->N, a->ldb, BK, BN)) return DVLA_ERR_CUDA; }
  else       { if (!make_tmap_2d_bf16(&tmB, a->b, a->N, a->K, a->ldb, 64, BK)) return DVLA_ERR_CUDA; }
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN>;
  static std::once_flag once;      // per instantiation
  static cudaError_t attr_err = cudaSuccess;
  if (!set_smem_attr_once(once, attr_err, kern, Cfg::SMEM_BYTES)) {
    set_error("cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(attr_err)); return DVLA_ERR_CUDA;
  }
  const int tiles_mn = p.num_m_tiles * p.num_n_tiles;
  const int tiles = p.tail_splits > 1 ? p.tail_first + (tiles_mn - p.tail_first) * p.tail_splits : tiles_mn * p.k_splits;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("gemm_tcgen05 launch: %s", cudaGetErrorString(e)); return DVLA_ERR_CUDA; }
  count_launch();
  return DVLA_OK;
}

template <bool A_MN, bool B_MN>
static int launch_tc2(const dvla_gemm_args* a, const GemmParams& p, cudaStream_t stream) {
  using Cfg = Gemm2Cfg;
  if (g_plan) {
    const int m2 = (p.M + 2 * BM - 1) / (2 * BM);
    const int units = gemm_total_units(m2 * p.num_n_tiles, p);
    record_plan(p, 3, 2 * BM, 256, m2, 2 * (units < num_sms() / 2 ? units : num_sms() / 2));
    return DVLA_OK;
  }
  CUtensorMap tmA, tmB;
  if (!A_MN) { if (!make_tmap_2d_bf16(&tmA, a->a, a->K, a->M, a->lda, BK, BM)) return DVLA_ERR_CUDA; }
  else       { if (!make_tmap_2d_bf16(&tmA, a->a, a->M, a->K, a->lda, 64, BK)) return DVLA_ERR_CUDA; }
  if (!B_MN) { if (!make_tmap_2d_bf16(&tmB, a->b, a->K, a->N, a->ldb, BK, 128)) return DVLA_ERR_CUDA; }
  else       { if (!make_tmap_2d_bf16(&tmB, a->b, a->N, a->K, a->ldb, 64, BK)) return DVLA_ERR_CUDA; }
  auto kern = gemm_tcgen05_2cta_kernel<A_MN, B_MN>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  if (!set_smem_attr_once(once, attr_err, kern, Cfg::SMEM_BYTES)) {
    set_error("cudaFuncSetAttribute(2cta smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(attr_err)); return DVLA_ERR_CUDA;
  }
  const int tiles_mn = ((p.M + 2 * BM - 1) / (2 * BM)) * p.num_n_tiles;
  const int tiles = p.tail_splits > 1 ? p.tail_first + (tiles_mn - p.tail_first) * p.tail_splits : tiles_mn * p.k_splits;
  const int max_clusters = num_sms() / 2;
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  kern<<<2 * clusters, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("gemm_tcgen05_2cta launch: %s", cudaGetErrorString(e)); return DVLA_ERR_CUDA; }
  count_launch();
  return DVLA_OK;
}

// 0 = auto, 1 = force single-CTA kernels, 2 = force the CTA-pair kernel whenever legal
static int gemm_mode() {
  static const int mode = [] {
    const char* e = getenv("DVLA_GEMM");
    return (e && !strcmp(e, "1cta")) ? 1 : (e && !strcmp(e, "2cta")) ? 2 : 0;
  }();
  return mode;
}

// DVLA_GEMM_SPLITK=0 disables the atomic split-K path (bit-reproducible gradient accumulation order)
static bool splitk_enabled() {
  static const bool on = [] { const char* e = getenv("DVLA_GEMM_SPLITK"); return !(e && !strcmp(e, "0")); }();
  return on;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// DVLA_GEMM_TAIL=0 disables the K-split tail (every unit of a plain GEMM is then a whole tile)
static bool tail_enabled() {
  static const bool on = [] { const char* e = getenv("DVLA_GEMM_TAIL"); return !(e && !strcmp(e, "0")); }();
  return on;
}
// split-K weight gradients: partial sums added with red.global.add.bf16x2 (default), or with DVLA_GEMM_SPLITK=fp32 the same
// fp32 slices + ordered fix-up as the tail: one bf16 rounding per accumulation (rel-L2 vs fp32 1.66e-3 instead of
// 2.7e-3 - 4.3e-3) and a reproducible summation order, for +5 % ... +49 % per weight-gradient GEMM (the last unit's read-back
// of the other splits' slices is exposed at the end of the kernel) = +1.4 ms per C2 step (profiles/r2_notes.md)
static bool splitk_fp32() {
  static const bool on = [] { const char* e = getenv("DVLA_GEMM_SPLITK"); return e && !strcmp(e, "fp32"); }();
  return on;
}
// workspace = [16384 arrival counters (64 KB), zero between launches] [fp32 partial tiles: 128 x BN x 4 bytes per CTA and split]
constexpr int64_t TAIL_CNT_BYTES = 65536;
constexpr int64_t TAIL_WS_SLICES = 512;              // 128 x 256 fp32 slices (64 MB): a split tail wave of 74 CTA pairs needs
                                                     // <= 148, a 4096 x 1024 weight gradient split in two 256
int64_t gemm_workspace_bytes() { return TAIL_CNT_BYTES + TAIL_WS_SLICES * 128 * 256 * 4; }

static bool tail_fits(const dvla_gemm_args* a, long long tail_tiles, int ctas_per_tile, int bn, int splits) {
  if (!a->workspace || (reinterpret_cast<uintptr_t>(a->workspace) & 15)) return false;
  const long long slots = tail_tiles * ctas_per_tile;
  return slots * NUM_EPI_WARPS * 4 <= TAIL_CNT_BYTES && TAIL_CNT_BYTES + slots * splits * 128LL * bn * 4 <= a->workspace_bytes;
}
static void set_tail(GemmParams& p, const dvla_gemm_args* a, int first, int splits, int kbps) {
  p.tail_first = first; p.tail_splits = splits; p.tail_kbps = kbps;
  p.tail_cnt = reinterpret_cast<unsigned*>(a->workspace);
  p.tail_ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a->workspace) + TAIL_CNT_BYTES);
}
// The tiles of a partly filled last wave are cut along K (see map_unit).  Model, in k-block times of one tile: a unit costs
// its k-blocks + c0 (fill / drain / epilogue); a tail unit additionally the fp32 round trip through L2 (c_fix).
static void plan_tail(GemmParams& p, const dvla_gemm_args* a, long long total_tiles, int slots, int ctas_per_tile, int bn) {
  if (!tail_enabled() || !p.staged_ok || p.k_splits != 1 || total_tiles <= 0) return;
  const long long full_waves = total_tiles / slots;
  const long long rem = total_tiles - full_waves * slots;
  // Only GEMMs of at least one full wave: a row's result then depends on M only for the large training GEMMs.  Below one
  // wave (every GEMM of action inference) each output row is computed in the same K order whatever the batch it is in --
  // the incremental rollout's "bit-identical to the full window" property (tests/test_rollout_gpu.py) rests on that.
  if (rem == 0 || full_waves == 0) return;
  int tsp = (int)(slots / rem);
  if (tsp > 8) tsp = 8;
  if (tsp > p.num_k_blocks / 4) tsp = p.num_k_blocks / 4;
  if (tsp < 2) return;
  const int kbps = (p.num_k_blocks + tsp - 1) / tsp;
  tsp = (p.num_k_blocks + kbps - 1) / kbps;
  if (tsp < 2) return;
  const float c0 = 8.f, c_fix = 20.f;    // ~8 us of slice writes + read-back vs 0.43 us per k-block of a 256 x 256 tile
  const float now = (float)(full_waves + 1) * ((float)p.num_k_blocks + c0);
  const float with_tail = (float)full_waves * ((float)p.num_k_blocks + c0) + (float)kbps + c0 + c_fix;
  if (with_tail > 0.93f * now) return;
  if (!tail_fits(a, rem, ctas_per_tile, bn, tsp)) return;
  set_tail(p, a, (int)(full_waves * slots), tsp, kbps);
}

int gemm_dispatch(const dvla_gemm_args* a, cudaStream_t stream) {
  if (!a || !a->a || !a->b || !a->out) { set_error("dvla_gemm: null pointer"); return DVLA_ERR_INVALID; }
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) { set_error("dvla_gemm: non-positive dims M=%lld N=%lld K=%lld", (long long)a->M, (long long)a->N, (long long)a->K); return DVLA_ERR_INVALID; }
  if (a->aux_in && a->aux_out) { set_error("dvla_gemm: aux_in and aux_out are exclusive"); return DVLA_ERR_INVALID; }
  if (a->dropout_p < 0.f || a->dropout_p >= 1.f) { set_error("dvla_gemm: dropout_p out of range"); return DVLA_ERR_INVALID; }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.out = a->out; p.bias = (const bf16*)a->bias; p.residual = a->residual;
  p.aux_out = (bf16*)a->aux_out; p.aux_in = (const bf16*)a->aux_in;
  p.M = (int)a->M; p.N = (int)a->N; p.K = (int)a->K;
  p.ldo = a->ldo; p.ldr = a->ldr; p.ld_aux = a->ld_aux;
  p.act = a->act; p.out_fp32 = a->out_fp32; p.alpha = a->alpha;
  if (a->dropout_p > 0.f) {
    p.drop_thresh = (uint32_t)(a->dropout_p * 65536.0f + 0.5f);
    p.drop_scale = 1.0f / (1.0f - (float)p.drop_thresh / 65536.0f);
    p.drop_seed = a->dropout_seed;
    p.drop_seed_ptr = a->dropout_seed_ptr;
  }
  const int oal = a->out_fp32 ? 4 : 8;
  p.vec_ok = aligned16(a->out) && (a->ldo % oal == 0) &&
             (!a->bias || aligned16(a->bias)) &&
             (!a->residual || (aligned16(a->residual) && a->ldr % oal == 0)) &&
             (!a->aux_out || (aligned16(a->aux_out) && a->ld_aux % 8 == 0)) &&
             (!a->aux_in || (aligned16(a->aux_in) && a->ld_aux % 8 == 0));
  p.staged_ok = p.vec_ok && !a->out_fp32 && (a->N % 8 == 0);
  p.num_m_tiles = (p.M + BM - 1) / BM;
  p.num_k_blocks = (p.K + BK - 1) / BK;

  const bool tma_ok = aligned16(a->a) && aligned16(a->b) && (a->lda % 8 == 0) && (a->ldb % 8 == 0);
  if (!tma_ok) {
    const long long groups = (long long)p.M * ((p.N + 7) / 8);
    const int threads = 128;
    if (g_plan) {
      const bool warp_kernel = p.K >= 128 && groups <= 65536;
      record_plan(p, warp_kernel ? 1 : 0, 1, 8, p.M, (int)(((warp_kernel ? groups * 32 : groups) + threads - 1) / threads));
      return DVLA_OK;
    }
    if (p.K >= 128 && groups <= 65536) {      // long contraction, small output: one warp per output group
      const long long blocks = (groups * 32 + threads - 1) / threads;
      gemm_simt_warp_kernel<<<(unsigned)blocks, threads, 0, stream>>>((const bf16*)a->a, (const bf16*)a->b, a->lda, a->ldb,
                                                                    a->a_mn_major, a->b_mn_major, p);
    } else {
      const long long blocks = (groups + threads - 1) / threads;
      gemm_simt_kernel<<<(unsigned)blocks, threads, 0, stream>>>((const bf16*)a->a, (const bf16*)a->b, a->lda, a->ldb,
                                                               a->a_mn_major, a->b_mn_major, p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("gemm_simt launch: %s", cudaGetErrorString(e)); return DVLA_ERR_CUDA; }
    count_launch();
    return DVLA_OK;
  }
  const int sms = num_sms();
  p.k_splits = 1;
  p.kb_per_split = p.num_k_blocks;
  // Split-K for pure accumulations (out += A.B: the weight-gradient GEMMs, few output tiles and a very long contraction):
  // work unit = (tile, k-range), partial sums added with red.global.add.bf16x2.  Pick the (kernel, split) pair with the
  // lowest modelled time = waves x (k-blocks per unit + fixed per-unit cost) x per-k-block tile cost.
  const bool accum_only = p.staged_ok && a->residual == a->out && a->ldr == a->ldo && !a->bias && !a->aux_in &&
                          !a->aux_out && a->act == DVLA_ACT_NONE && a->dropout_p == 0.f;
  if (accum_only && splitk_enabled() && gemm_mode() == 0 && p.num_k_blocks >= 32) {
    const int smax = p.num_k_blocks / 16 < 16 ? p.num_k_blocks / 16 : 16;
    const long long tl[3] = {(long long)((p.M + 2 * BM - 1) / (2 * BM)) * ((p.N + 255) / 256),
                             (long long)p.num_m_tiles * ((p.N + 255) / 256), (long long)p.num_m_tiles * ((p.N + 127) / 128)};
    const int slots[3] = {sms / 2, sms, sms};
    const float w[3] = {1.7f, 1.8f, 1.0f};
    const float c0 = 8.f;
    float best = 1e30f; int best_cfg = -1, best_s = 1;
    for (int cfg = 0; cfg < 3; ++cfg) {
      if (cfg < 2 && p.N <= 128) continue;
      for (int sp = 1; sp <= smax; ++sp) {
        const int kbps = (p.num_k_blocks + sp - 1) / sp;
        const int se = (p.num_k_blocks + kbps - 1) / kbps;
        if (se != sp) continue;
        const long long waves = (tl[cfg] * sp + slots[cfg] - 1) / slots[cfg];
        const float t = (float)waves * ((float)kbps + c0) * w[cfg];
        if (t < best * 0.97f || best_cfg < 0) { best = t; best_cfg = cfg; best_s = sp; }
      }
    }
    if (best_s > 1) {
      const int kbps = (p.num_k_blocks + best_s - 1) / best_s;
      const int bn_sel = best_cfg == 2 ? 128 : 256;
      if (splitk_fp32() && tail_fits(a, tl[best_cfg], best_cfg == 0 ? 2 : 1, bn_sel, best_s)) {
        // every tile is split; the partial sums meet in fp32 slices and the last unit adds the complete product to the
        // gradient (out = residual = G): ONE bf16 rounding per accumulation instead of one per split, fixed summation order
        set_tail(p, a, 0, best_s, kbps);
      } else {
        p.k_splits = best_s;
        p.kb_per_split = kbps;
        p.atomic_out = 1;
        p.residual = nullptr;
      }
      const int key2 = (a->a_mn_major ? 2 : 0) | (a->b_mn_major ? 1 : 0);
      if (best_cfg == 0) {
        p.num_n_tiles = (p.N + 255) / 256;
        switch (key2) {
          case 0: return launch_tc2<false, false>(a, p, stream);
          case 1: return launch_tc2<false, true>(a, p, stream);
          case 2: return launch_tc2<true, false>(a, p, stream);
          default: return launch_tc2<true, true>(a, p, stream);
        }
      }
      const int bn = best_cfg == 1 ? 256 : 128;
      p.num_n_tiles = (p.N + bn - 1) / bn;
      switch ((best_cfg == 1 ? 4 : 0) | key2) {
        case 0: return launch_tc<128, false, false>(a, p, stream);
        case 1: return launch_tc<128, false, true>(a, p, stream);
        case 2: return launch_tc<128, true, false>(a, p, stream);
        case 3: return launch_tc<128, true, true>(a, p, stream);
        case 4: return launch_tc<256, false, false>(a, p, stream);
        case 5: return launch_tc<256, false, true>(a, p, stream);
        case 6: return launch_tc<256, true, false>(a, p, stream);
        default: return launch_tc<256, true, true>(a, p, stream);
      }
    }
  }
  // CTA-pair kernel (256x256 tiles, staged bf16 epilogue): when the problem fills the 74 SM pairs at least as well as
  // the single-CTA tiling fills the 148 SMs
  if (p.staged_ok && gemm_mode() != 1) {
    const long long t2 = (long long)((p.M + 2 * BM - 1) / (2 * BM)) * ((p.N + 255) / 256);
    const long long t1 = (long long)p.num_m_tiles * ((p.N + 255) / 256);
    const long long r2 = (t2 + sms / 2 - 1) / (sms / 2), r1 = (t1 + sms - 1) / sms;
    const bool big = (long long)p.M * p.N >= 256LL * 256 * 40;
    if (gemm_mode() == 2 || (big && p.N > 128 && r2 <= r1)) {
      p.num_n_tiles = (p.N + 255) / 256;
      plan_tail(p, a, t2, sms / 2, 2, 256);
      const int key2 = (a->a_mn_major ? 2 : 0) | (a->b_mn_major ? 1 : 0);
      switch (key2) {
        case 0: return launch_tc2<false, false>(a, p, stream);
        case 1: return launch_tc2<false, true>(a, p, stream);
        case 2: return launch_tc2<true, false>(a, p, stream);
        default: return launch_tc2<true, true>(a, p, stream);
      }
    }
  }
  // tile-width heuristic: fewer, fatter tiles unless that leaves SMs idle for a whole extra wave
  const long long t128 = (long long)p.num_m_tiles * ((p.N + 127) / 128);
  const long long t256 = (long long)p.num_m_tiles * ((p.N + 255) / 256);
  const long long cost128 = (t128 + sms - 1) / sms;
  const long long cost256 = 2 * ((t256 + sms - 1) / sms);
  const bool use256 = (p.N > 128) && (cost256 <= cost128);
  const int BNsel = use256 ? 256 : 128;
  p.num_n_tiles = (p.N + BNsel - 1) / BNsel;
  plan_tail(p, a, use256 ? t256 : t128, sms, 1, BNsel);
  const int key = (use256 ? 4 : 0) | (a->a_mn_major ? 2 : 0) | (a->b_mn_major ? 1 : 0);
  switch (key) {
    case 0: return launch_tc<128, false, false>(a, p, stream);
    case 1: return launch_tc<128, false, true>(a, p, stream);
    case 2: return launch_tc<128, true, false>(a, p, stream);
    case 3: return launch_tc<128, true, true>(a, p, stream);
    case 4: return launch_tc<256, false, false>(a, p, stream);
    case 5: return launch_tc<256, false, true>(a, p, stream);
    case 6: return launch_tc<256, true, false>(a, p, stream);
    default: return launch_tc<256, true, true>(a, p, stream);
  }
}

int gemm_plan(const dvla_gemm_args* a, dvla_gemm_plan_info* out) {
  if (!out) { set_error("dvla_gemm_plan: null output"); return DVLA_ERR_INVALID; }
  memset(out, 0, sizeof(*out));
  g_plan = out;
  const int rc = gemm_dispatch(a, nullptr);
  g_plan = nullptr;
  return rc;
}
int gemm_plan_unit(const dvla_gemm_plan_info* plan, int unit, int* tile, int* kb0, int* kb1, int* slot, int* split) {
  if (!plan || plan->kernel < 2 || unit < 0 || unit >= plan->units) { set_error("dvla_gemm_plan_unit: bad plan / unit"); return DVLA_ERR_INVALID; }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.num_k_blocks = plan->k_blocks; p.k_splits = plan->k_splits; p.kb_per_split = plan->kb_per_split;
  p.tail_first = plan->tail_first; p.tail_splits = plan->tail_splits; p.tail_kbps = plan->tail_kbps;
  const Unit u = map_unit(unit, plan->m_tiles * plan->n_tiles, p);
  if (tile) *tile = u.tile;
  if (kb0) *kb0 = u.kb0;
  if (kb1) *kb1 = u.kb1;
  if (slot) *slot = u.slot;
  if (split) *split = u.split;
  return DVLA_OK;
}

}  // namespace dvla
