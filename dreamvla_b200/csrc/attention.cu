// Flash-style multi-head attention for head_dim = 64 with a bit-matrix visibility mask and tile skipping.
//
// Data path of THIS file: cp.async -> XOR-swizzled smem tiles -> ldmatrix -> mma.sync.m16n8k16 (bf16, fp32 accumulate),
// online softmax in registers, no [Lq,Lk] tensor in HBM.  These kernels serve the short sequences of the path (DiT: 4
// tokens, resampler: 16 queries) and strides a tensor map cannot express; sequences of >= 96 tokens go to the
// warp-specialised tcgen05 / TMEM kernels of attention_fwd_ws.cu / attention_bwd_ws.cu through the dispatchers below.
//
// Kernels:
//   attn_fwd_kernel     grid (ceil(Lq/64), H, B), 4 warps x 16 query rows, K/V tiles of 64 keys double-buffered
//   attn_delta_kernel   delta[b,h,i] = sum_d dO*O
//   attn_bwd_dkv_kernel grid (ceil(Lk/64), H, B): per key tile, loops over query tiles, S^T/dP^T formulation
//   attn_bwd_dq_kernel  grid (ceil(Lq/64), H, B): per query tile, loops over key tiles
// Masking: tile_flags[qt, kt] in {0 skip, 1 partial, 2 full}; partial tiles test bits of mask[i, j/32].
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "../../include/dvla.h"

namespace dvla {
void set_error(const char* fmt, ...);
void count_launch();
int num_sms();

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct AttnParams {
  const bf16 *q, *k, *v, *o, *d_o;
  bf16 *out, *dq, *dk, *dv;
  float* lse;
  float* delta;
  const uint32_t* mask;
  const uint8_t* tile_flags;
  int B, H, Lq, Lk, nqt, nkt, mask_words;
  int qt0, kt0;          // first q / kv 64-row tile of this launch (tail launches next to the tcgen05 kernels)
  long long q_sb, q_ss, q_sh, k_sb, k_ss, k_sh, v_sb, v_ss, v_sh, o_sb, o_ss, o_sh;
  long long do_sb, do_ss, do_sh, dq_sb, dq_ss, dq_sh, dk_sb, dk_ss, dk_sh, dv_sb, dv_ss, dv_sh;
  float scale;
  float drop_scale; uint32_t drop_thresh; uint64_t drop_seed; const uint64_t* drop_seed_ptr;
};

// ---- smem tile helpers: 64 rows x 64 bf16 (128 B per row), 16-byte chunk c of row r stored at chunk c ^ (r & 7) --------
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
  return base + row * 128 + ((chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// rows [r0, r0+64) of a strided [L, 64] matrix -> swizzled tile; rows >= L are zero-filled.  128 threads.
__device__ __forceinline__ void load_tile_async(uint32_t sbase, const bf16* g, long long row_stride, int r0, int L) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int i = threadIdx.x + it * 128;
    const int row = i >> 3, ch = i & 7;
    const int gr = r0 + row;
    const bf16* src = g + static_cast<long long>(gr < L ? gr : (L - 1)) * row_stride + ch * 8;
    cp_async16(tile_addr(sbase, row, ch), src, gr < L ? 16 : 0);
  }
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragments (16 rows x 64 cols) of a tile for this warp's 16 rows starting at row r0: frag[kk][0..3]
__device__ __forceinline__ void load_a_frags(uint32_t tile, int r0, uint32_t (&f)[4][4]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) ldsm_x4(tile_addr(tile, r0 + (lane & 15), kk * 2 + (lane >> 4)), f[kk]);
}
// C[16 x 64] (+)= A[16 x 64(k)] * T^T where T is a [64(n) x 64(k)] row-major tile (B fragment = plain ldmatrix)
__device__ __forceinline__ void mma_a_tT(float (&c)[8][4], const uint32_t (&a)[4][4], uint32_t tile) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4(tile_addr(tile, np * 16 + (lane & 7) + ((lane >> 4) << 3), kk * 2 + ((lane >> 3) & 1)), b);
      mma16816(c[2 * np], a[kk], b[0], b[1]);
      mma16816(c[2 * np + 1], a[kk], b[2], b[3]);
    }
  }
}
// C[16 x 64(n)] += P[16 x 64(k)] * T where T is a [64(k) x 64(n)] row-major tile (B fragment = ldmatrix.trans);
// P given as packed bf16 A fragments pa[kk][0..3].
__device__ __forceinline__ void mma_p_t(float (&c)[8][4], const uint32_t (&pa)[4][4], uint32_t tile) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldsm_x4_t(tile_addr(tile, kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), np * 2 + (lane >> 4)), b);
      mma16816(c[2 * np], pa[kk], b[0], b[1]);
      mma16816(c[2 * np + 1], pa[kk], b[2], b[3]);
    }
  }
}
// C-fragment layout (16 x 64 fp32) -> A-fragment layout (bf16) for the next MMA
__device__ __forceinline__ void c_to_a(const float (&c)[8][4], uint32_t (&a)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    a[kk][0] = pack_bf16x2(c[2 * kk][0], c[2 * kk][1]);
    a[kk][1] = pack_bf16x2(c[2 * kk][2], c[2 * kk][3]);
    a[kk][2] = pack_bf16x2(c[2 * kk + 1][0], c[2 * kk + 1][1]);
    a[kk][3] = pack_bf16x2(c[2 * kk + 1][2], c[2 * kk + 1][3]);
  }
}
__device__ __forceinline__ void zero_c(float (&c)[8][4]) {
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int e = 0; e < 4; ++e) c[t][e] = 0.f;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// Dropout keep bits for this thread's elements of a 16x64 C fragment whose rows are `grow0 + {g, g+8}` and whose
// columns are `gcol0 + t*8 + 2c + {0,1}`; RNG block = (bh*Lrow + grow) * ceil(Lcol/8) + gcol/8 (8 consecutive columns).
// The 4 lanes of a quad split the 16 (row, n-tile) Philox calls and exchange 8-bit masks by shuffle.
// `transposed`: fragment holds S^T (rows are keys, cols are queries) -> (row,col) swap for the index.
__device__ __forceinline__ void dropout_bits(const AttnParams& p, long long bh, int frag_row0, int frag_col0,
                                             bool transposed, uint32_t (&keep)[8] /* per n-tile: 4 bits e0..e3 */) {
  const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const int nblk = (p.Lk + 7) >> 3;
  const uint64_t seed = p.drop_seed + (p.drop_seed_ptr ? __ldg(p.drop_seed_ptr) : 0ull);
  if (!transposed) {
    // rows = queries, cols = keys: one RNG block == one n-tile of one row.
    uint32_t m[2][2];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int t = c + half * 4;
        const long long row = frag_row0 + g + r * 8;
        m[r][half] = dropout_keep8(seed, (bh * p.Lq + row) * nblk + ((frag_col0 >> 3) + t), p.drop_thresh);
      }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int src = (lane & ~3) | (t & 3);
      const uint32_t m0 = __shfl_sync(0xffffffffu, m[0][t >> 2], src);
      const uint32_t m1 = __shfl_sync(0xffffffffu, m[1][t >> 2], src);
      keep[t] = ((m0 >> (2 * c)) & 3u) | (((m1 >> (2 * c)) & 3u) << 2);
    }
  } else {
    // rows = keys (g, g+8 -> same 8-key RNG block iff ... no: keys g and g+8 are in different blocks), cols = queries.
    // element (key kr, query qc): block = (bh*Lq + qc) * nblk + kr/8, bit kr%8.
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      uint32_t bits = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kr = frag_row0 + g + (e >> 1) * 8;
        const long long qc = frag_col0 + t * 8 + 2 * c + (e & 1);
        const uint32_t mk = dropout_keep8(seed, (bh * p.Lq + qc) * nblk + (kr >> 3), p.drop_thresh);
        bits |= ((mk >> (kr & 7)) & 1u) << e;
      }
      keep[t] = bits;
    }
  }
}

// ===================================================== forward ====================================================
__global__ void __launch_bounds__(128) attn_fwd_kernel(const AttnParams p) {
  __shared__ __align__(1024) uint8_t sm[5 * 8192];
  const uint32_t sQ = smem_u32(sm), sK0 = sQ + 8192, sV0 = sQ + 2 * 8192, sK1 = sQ + 3 * 8192, sV1 = sQ + 4 * 8192;
  const int qt = blockIdx.x + p.qt0, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const int q0 = qt * 64;
  const bf16* qg = p.q + b * p.q_sb + h * p.q_sh;
  const bf16* kg = p.k + b * p.k_sb + h * p.k_sh;
  const bf16* vg = p.v + b * p.v_sb + h * p.v_sh;
  const uint8_t* flags = p.tile_flags ? p.tile_flags + static_cast<long long>(qt) * p.nkt : nullptr;

  auto next_tile = [&](int kt) {
    while (kt < p.nkt && flags && flags[kt] == 0) ++kt;
    return kt;
  };

  float o[8][4];
  zero_c(o);
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float sc = p.scale * LOG2E;
  const long long bh = static_cast<long long>(b) * p.H + h;

  int kt = next_tile(0);
  load_tile_async(sQ, qg, p.q_ss, q0, p.Lq);
  if (kt < p.nkt) {
    load_tile_async(sK0, kg, p.k_ss, kt * 64, p.Lk);
    load_tile_async(sV0, vg, p.v_ss, kt * 64, p.Lk);
  }
  cp_async_commit();
  uint32_t qf[4][4];
  bool q_loaded = false;
  int buf = 0;
  while (kt < p.nkt) {
    const int kt_next = next_tile(kt + 1);
    if (kt_next < p.nkt) {
      load_tile_async(buf ? sK0 : sK1, kg, p.k_ss, kt_next * 64, p.Lk);
      load_tile_async(buf ? sV0 : sV1, vg, p.v_ss, kt_next * 64, p.Lk);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (!q_loaded) { load_a_frags(sQ, warp * 16, qf); q_loaded = true; }
    const uint32_t sK = buf ? sK1 : sK0, sV = buf ? sV1 : sV0;

    float s[8][4];
    zero_c(s);
    mma_a_tT(s, qf, sK);

    // ---- mask + scale (log2 domain) ----
    const int flag = flags ? flags[kt] : 2;
    const int k0 = kt * 64;
    const bool tail = (k0 + 64 > p.Lk);
    uint32_t w[2][2] = {{0xffffffffu, 0xffffffffu}, {0xffffffffu, 0xffffffffu}};
    if (flag == 1) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = q0 + warp * 16 + g + r * 8;
#pragma unroll
        for (int ww = 0; ww < 2; ++ww) {
          const int wi = kt * 2 + ww;
          w[r][ww] = (row < p.Lq && wi < p.mask_words) ? p.mask[static_cast<long long>(row) * p.mask_words + wi] : 0u;
        }
      }
    }
    float tmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int t = 0; t < 8; ++t) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        const int kk = t * 8 + 2 * c + (e & 1);
        bool vis = true;
        if (flag == 1) vis = (w[r][kk >> 5] >> (kk & 31)) & 1u;
        if (tail) vis = vis && (k0 + kk < p.Lk);
        const float v = vis ? s[t][e] * sc : -INFINITY;
        s[t][e] = v;
        tmax[r] = fmaxf(tmax[r], v);
      }
    }
    float alpha[2], muse[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float mn = fmaxf(m_run[r], quad_max(tmax[r]));
      muse[r] = (mn == -INFINITY) ? 0.f : mn;
      alpha[r] = exp2f(m_run[r] - muse[r]);
      m_run[r] = mn;
      l_run[r] *= alpha[r];
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        const float pv = exp2f(s[t][e] - muse[r]);
        s[t][e] = pv;
        l_run[r] += pv;
        o[t][e] *= alpha[r];
      }
    }
    if (p.drop_scale != 0.f) {
      uint32_t keep[8];
      dropout_bits(p, bh, q0 + warp * 16, k0, false, keep);
#pragma unroll
      for (int t = 0; t < 8; ++t)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[t][e] = ((keep[t] >> e) & 1u) ? s[t][e] * p.drop_scale : 0.f;
    }
    uint32_t pa[4][4];
    c_to_a(s, pa);
    mma_p_t(o, pa, sV);
    __syncthreads();
    kt = kt_next;
    buf ^= 1;
  }
  cp_async_wait<0>();

  // ---- finalize ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float l = quad_sum(l_run[r]);
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    const int row = q0 + warp * 16 + g + r * 8;
    if (row < p.Lq) {
      bf16* dst = p.out + b * p.o_sb + static_cast<long long>(row) * p.o_ss + h * p.o_sh;
#pragma unroll
      for (int t = 0; t < 8; ++t)
        *reinterpret_cast<uint32_t*>(dst + t * 8 + 2 * c) = pack_bf16x2(o[t][2 * r] * inv, o[t][2 * r + 1] * inv);
      if (p.lse && c == 0) p.lse[(bh * p.Lq) + row] = (l > 0.f) ? (m_run[r] + log2f(l)) * LN2 : -INFINITY;
    }
  }
}

// ===================================================== delta ======================================================
__global__ void __launch_bounds__(256) attn_delta_kernel(const AttnParams p) {
  // 8 lanes per (b, h, i) row of 64 elements
  const long long idx = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  const long long total = static_cast<long long>(p.B) * p.H * p.Lq;
  float acc = 0.f;
  if (idx < total) {
    const int i = static_cast<int>(idx % p.Lq);
    const int h = static_cast<int>((idx / p.Lq) % p.H);
    const int b = static_cast<int>(idx / (static_cast<long long>(p.Lq) * p.H));
    const uint4 ou = *reinterpret_cast<const uint4*>(p.o + b * p.o_sb + static_cast<long long>(i) * p.o_ss + h * p.o_sh + sub * 8);
    const uint4 du = *reinterpret_cast<const uint4*>(p.d_o + b * p.do_sb + static_cast<long long>(i) * p.do_ss + h * p.do_sh + sub * 8);
    float2 a, d;
    a = unpack_bf16x2(ou.x); d = unpack_bf16x2(du.x); acc += a.x * d.x + a.y * d.y;
    a = unpack_bf16x2(ou.y); d = unpack_bf16x2(du.y); acc += a.x * d.x + a.y * d.y;
    a = unpack_bf16x2(ou.z); d = unpack_bf16x2(du.z); acc += a.x * d.x + a.y * d.y;
    a = unpack_bf16x2(ou.w); d = unpack_bf16x2(du.w); acc += a.x * d.x + a.y * d.y;
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (idx < total && sub == 0) p.delta[idx] = acc;  // layout [B, H, Lq] == idx ordering (b, h, i)? see dispatch
}

// ===================================================== dK / dV ====================================================
// One CTA per (key tile, h, b); warp w owns keys [k0 + 16w, k0 + 16w + 16).  Works on S^T = K Q^T (rows = keys).
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const AttnParams p) {
  __shared__ __align__(1024) uint8_t sm[4 * 8192];
  __shared__ float s_lse[64], s_delta[64];
  __shared__ uint32_t s_mask[64][2];
  const uint32_t sK = smem_u32(sm), sV = sK + 8192, sQ = sK + 2 * 8192, sDO = sK + 3 * 8192;
  const int kt = blockIdx.x + p.kt0, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const int k0 = kt * 64;
  const bf16* qg = p.q + b * p.q_sb + h * p.q_sh;
  const bf16* kg = p.k + b * p.k_sb + h * p.k_sh;
  const bf16* vg = p.v + b * p.v_sb + h * p.v_sh;
  const bf16* dog = p.d_o + b * p.do_sb + h * p.do_sh;
  const long long bh = static_cast<long long>(b) * p.H + h;
  const float sc = p.scale * LOG2E;

  load_tile_async(sK, kg, p.k_ss, k0, p.Lk);
  load_tile_async(sV, vg, p.v_ss, k0, p.Lk);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  uint32_t kf[4][4], vf[4][4];
  load_a_frags(sK, warp * 16, kf);
  load_a_frags(sV, warp * 16, vf);

  float dk[8][4], dv[8][4];
  zero_c(dk);
  zero_c(dv);

  for (int qt = 0; qt < p.nqt; ++qt) {
    const int flag = p.tile_flags ? p.tile_flags[static_cast<long long>(qt) * p.nkt + kt] : 2;
    if (flag == 0) continue;
    const int q0 = qt * 64;
    __syncthreads();  // previous iteration's readers are done with sQ / sDO / s_*
    load_tile_async(sQ, qg, p.q_ss, q0, p.Lq);
    load_tile_async(sDO, dog, p.do_ss, q0, p.Lq);
    cp_async_commit();
    if (threadIdx.x < 64) {
      const int qi = q0 + threadIdx.x;
      s_lse[threadIdx.x] = (qi < p.Lq) ? p.lse[bh * p.Lq + qi] * LOG2E : INFINITY;  // +inf => p = 0 for padded rows
      s_delta[threadIdx.x] = (qi < p.Lq) ? p.delta[bh * p.Lq + qi] : 0.f;
      if (flag == 1) {
#pragma unroll
        for (int ww = 0; ww < 2; ++ww) {
          const int wi = kt * 2 + ww;
          s_mask[threadIdx.x][ww] = (qi < p.Lq && wi < p.mask_words) ? p.mask[static_cast<long long>(qi) * p.mask_words + wi] : 0u;
        }
      }
    }
    cp_async_wait<0>();
    __syncthreads();

    // S^T[16 keys x 64 q] = K_w Q^T
    float st[8][4];
    zero_c(st);
    mma_a_tT(st, kf, sQ);
    // dP^T[16 keys x 64 q] = V_w dO^T
    float dpt[8][4];
    zero_c(dpt);
    mma_a_tT(dpt, vf, sDO);

    uint32_t keep[8];
    if (p.drop_scale != 0.f) dropout_bits(p, bh, k0 + warp * 16, q0, true, keep);

    float pt[8][4];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kr = warp * 16 + g + (e >> 1) * 8;      // key row within tile
        const int qc = t * 8 + 2 * c + (e & 1);           // query col within tile
        bool vis = (k0 + kr < p.Lk);
        if (flag == 1) vis = vis && ((s_mask[qc][kr >> 5] >> (kr & 31)) & 1u);
        const float pv = vis ? exp2f(st[t][e] * sc - s_lse[qc]) : 0.f;
        float dpv = dpt[t][e];
        float pdrop = pv;
        if (p.drop_scale != 0.f) {
          const bool kp = (keep[t] >> e) & 1u;
          pdrop = kp ? pv * p.drop_scale : 0.f;
          dpv = kp ? dpv * p.drop_scale : 0.f;
        }
        pt[t][e] = pdrop;                                  // for dV
        st[t][e] = pv * (dpv - s_delta[qc]);               // dS^T
      }
    }
    uint32_t pa[4][4];
    c_to_a(pt, pa);
    mma_p_t(dv, pa, sDO);   // dV[16 keys x 64 hd] += P^T[16 x 64 q] * dO[64 q x 64 hd]
    c_to_a(st, pa);
    mma_p_t(dk, pa, sQ);    // dK += dS^T * Q
  }

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int key = k0 + warp * 16 + g + r * 8;
    if (key < p.Lk) {
      bf16* dkd = p.dk + b * p.dk_sb + static_cast<long long>(key) * p.dk_ss + h * p.dk_sh;
      bf16* dvd = p.dv + b * p.dv_sb + static_cast<long long>(key) * p.dv_ss + h * p.dv_sh;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        *reinterpret_cast<uint32_t*>(dkd + t * 8 + 2 * c) = pack_bf16x2(dk[t][2 * r] * p.scale, dk[t][2 * r + 1] * p.scale);
        *reinterpret_cast<uint32_t*>(dvd + t * 8 + 2 * c) = pack_bf16x2(dv[t][2 * r], dv[t][2 * r + 1]);
      }
    }
  }
}

// ===================================================== dQ =========================================================
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const AttnParams p) {
  __shared__ __align__(1024) uint8_t sm[4 * 8192];
  const uint32_t sQ = smem_u32(sm), sDO = sQ + 8192, sK = sQ + 2 * 8192, sV = sQ + 3 * 8192;
  const int qt = blockIdx.x + p.qt0, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const int q0 = qt * 64;
  const bf16* qg = p.q + b * p.q_sb + h * p.q_sh;
  const bf16* kg = p.k + b * p.k_sb + h * p.k_sh;
  const bf16* vg = p.v + b * p.v_sb + h * p.v_sh;
  const bf16* dog = p.d_o + b * p.do_sb + h * p.do_sh;
  const long long bh = static_cast<long long>(b) * p.H + h;
  const float sc = p.scale * LOG2E;
  const uint8_t* flags = p.tile_flags ? p.tile_flags + static_cast<long long>(qt) * p.nkt : nullptr;

  load_tile_async(sQ, qg, p.q_ss, q0, p.Lq);
  load_tile_async(sDO, dog, p.do_ss, q0, p.Lq);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  uint32_t qf[4][4], dof[4][4];
  load_a_frags(sQ, warp * 16, qf);
  load_a_frags(sDO, warp * 16, dof);
  float lse2[2], dlt[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + warp * 16 + g + r * 8;
    lse2[r] = (row < p.Lq) ? p.lse[bh * p.Lq + row] * LOG2E : INFINITY;
    dlt[r] = (row < p.Lq) ? p.delta[bh * p.Lq + row] : 0.f;
  }
  float dq[8][4];
  zero_c(dq);

  for (int kt = 0; kt < p.nkt; ++kt) {
    const int flag = flags ? flags[kt] : 2;
    if (flag == 0) continue;
    const int k0 = kt * 64;
    __syncthreads();
    load_tile_async(sK, kg, p.k_ss, k0, p.Lk);
    load_tile_async(sV, vg, p.v_ss, k0, p.Lk);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    float s[8][4];
    zero_c(s);
    mma_a_tT(s, qf, sK);       // S = Q K^T
    float dp[8][4];
    zero_c(dp);
    mma_a_tT(dp, dof, sV);     // dP = dO V^T

    uint32_t w[2][2] = {{0xffffffffu, 0xffffffffu}, {0xffffffffu, 0xffffffffu}};
    if (flag == 1) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = q0 + warp * 16 + g + r * 8;
#pragma unroll
        for (int ww = 0; ww < 2; ++ww) {
          const int wi = kt * 2 + ww;
          w[r][ww] = (row < p.Lq && wi < p.mask_words) ? p.mask[static_cast<long long>(row) * p.mask_words + wi] : 0u;
        }
      }
    }
    uint32_t keep[8];
    if (p.drop_scale != 0.f) dropout_bits(p, bh, q0 + warp * 16, k0, false, keep);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        const int kk = t * 8 + 2 * c + (e & 1);
        bool vis = (k0 + kk < p.Lk);
        if (flag == 1) vis = vis && ((w[r][kk >> 5] >> (kk & 31)) & 1u);
        const float pv = vis ? exp2f(s[t][e] * sc - lse2[r]) : 0.f;
        float dpv = dp[t][e];
        if (p.drop_scale != 0.f) dpv = ((keep[t] >> e) & 1u) ? dpv * p.drop_scale : 0.f;
        s[t][e] = pv * (dpv - dlt[r]);   // dS
      }
    }
    uint32_t pa[4][4];
    c_to_a(s, pa);
    mma_p_t(dq, pa, sK);       // dQ += dS K
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + warp * 16 + g + r * 8;
    if (row < p.Lq) {
      bf16* dst = p.dq + b * p.dq_sb + static_cast<long long>(row) * p.dq_ss + h * p.dq_sh;
#pragma unroll
      for (int t = 0; t < 8; ++t)
        *reinterpret_cast<uint32_t*>(dst + t * 8 + 2 * c) = pack_bf16x2(dq[t][2 * r] * p.scale, dq[t][2 * r + 1] * p.scale);
    }
  }
}

// ===================================================== mask tiles =================================================
__global__ void attn_mask_tiles_kernel(const uint32_t* __restrict__ mask, int mask_words, int Lq, int Lk, int nkt,
                                       uint8_t* __restrict__ flags) {
  // one warp per (qt, kt)
  const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int nqt = (Lq + 63) / 64;
  if (tile >= nqt * nkt) return;
  const int qt = tile / nkt, kt = tile % nkt;
  int any = 0, all = 1;
  for (int i = lane; i < 128; i += 32) {  // 64 rows x 2 words
    const int row = qt * 64 + (i >> 1), wi = kt * 2 + (i & 1);
    if (row >= Lq) continue;
    const int kbase = wi * 32;
    if (kbase >= Lk) continue;
    const uint32_t valid = (Lk - kbase >= 32) ? 0xffffffffu : ((1u << (Lk - kbase)) - 1u);
    const uint32_t w = (wi < mask_words ? mask[static_cast<long long>(row) * mask_words + wi] : 0u) & valid;
    any |= (w != 0u);
    all &= (w == valid);
  }
  any = __any_sync(0xffffffffu, any);
  all = __all_sync(0xffffffffu, all);
  if (lane == 0) flags[tile] = any ? (all ? 2 : 1) : 0;
}

// ===================================================== host =======================================================
#define DVLA_CHECK_LAUNCH(name)                                                                 \
  do {                                                                                          \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess) { set_error("%s launch: %s", name, cudaGetErrorString(e__)); return DVLA_ERR_CUDA; } \
    count_launch();                                                                             \
  } while (0)

static bool strides_ok(const void* ptr, long long sb, long long ss, long long sh) {
  return ptr && (reinterpret_cast<uintptr_t>(ptr) % 16 == 0) && sb % 8 == 0 && ss % 8 == 0 && sh % 8 == 0;
}

bool attn_small_applicable(int64_t Lq, int64_t Lk, const void* mask, float dropout_p);   // attention_small.cu
int attn_small_fwd_dispatch(const dvla_attn_fwd_args* a, cudaStream_t s);
int attn_small_bwd_dispatch(const dvla_attn_bwd_args* a, cudaStream_t s);
static bool attn_small_enabled() {
  static const bool on = [] { const char* e = getenv("DVLA_ATTN_SMALL"); return !(e && e[0] == '0'); }();   // thread-safe init
  return on;
}
int attn_fwd_ws_dispatch(const dvla_attn_fwd_args* a, cudaStream_t s, long long q_rows);   // attention_fwd_ws.cu (tcgen05, warp-specialised)

// 0 = auto (warp-specialised tcgen05 kernel for Lq >= 96, SIMT row kernel for Lq <= 32 without mask / dropout, mma.sync
// kernel otherwise), 1 = mma.sync everywhere, 3 = warp-specialised tcgen05 kernel whenever expressible
static int attn_fwd_mode() {
  static const int mode = [] {
    const char* e = getenv("DVLA_ATTN_FWD");
    return (e && !strcmp(e, "legacy")) ? 1 : (e && !strcmp(e, "ws")) ? 3 : 0;
  }();
  return mode;
}

int attn_fwd_dispatch(const dvla_attn_fwd_args* a, cudaStream_t s) {
  if (!a) { set_error("attn_fwd: null args"); return DVLA_ERR_INVALID; }
  if (!strides_ok(a->q, a->q_sb, a->q_ss, a->q_sh) || !strides_ok(a->k, a->k_sb, a->k_ss, a->k_sh) ||
      !strides_ok(a->v, a->v_sb, a->v_ss, a->v_sh) || !strides_ok(a->o, a->o_sb, a->o_ss, a->o_sh)) {
    set_error("attn_fwd: q/k/v/o must be non-null, 16-byte aligned with strides multiple of 8 elements");
    return DVLA_ERR_INVALID;
  }
  if (a->B <= 0 || a->H <= 0 || a->Lq <= 0 || a->Lk <= 0) { set_error("attn_fwd: non-positive dims"); return DVLA_ERR_INVALID; }
  if (a->H > 65535 || a->B > 65535) { set_error("attn_fwd: B,H must be <= 65535"); return DVLA_ERR_UNSUPPORTED; }
  if (a->mask && a->mask_words * 32 < a->Lk) { set_error("attn_fwd: mask_words too small"); return DVLA_ERR_INVALID; }
  if (a->mask && !a->tile_flags) { set_error("attn_fwd: mask given without tile_flags"); return DVLA_ERR_INVALID; }
  const int mode = attn_fwd_mode();
  if (a->key_bias) {
    if (!attn_small_applicable(a->Lq, a->Lk, a->mask, a->dropout_p)) {
      set_error("attn_fwd: key_bias needs the short-sequence kernels (Lq <= 32, Lk <= 64, no mask, no dropout)");
      return DVLA_ERR_UNSUPPORTED;
    }
    return attn_small_fwd_dispatch(a, s);
  }
  if (mode == 0 && attn_small_enabled() && attn_small_applicable(a->Lq, a->Lk, a->mask, a->dropout_p))
    return attn_small_fwd_dispatch(a, s);
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.q = (const bf16*)a->q; p.k = (const bf16*)a->k; p.v = (const bf16*)a->v; p.out = (bf16*)a->o; p.lse = a->lse;
  p.mask = a->mask; p.tile_flags = a->mask ? a->tile_flags : nullptr;
  p.B = (int)a->B; p.H = (int)a->H; p.Lq = (int)a->Lq; p.Lk = (int)a->Lk;
  p.nqt = (p.Lq + 63) / 64; p.nkt = (p.Lk + 63) / 64; p.mask_words = a->mask_words;
  p.q_sb = a->q_sb; p.q_ss = a->q_ss; p.q_sh = a->q_sh; p.k_sb = a->k_sb; p.k_ss = a->k_ss; p.k_sh = a->k_sh;
  p.v_sb = a->v_sb; p.v_ss = a->v_ss; p.v_sh = a->v_sh; p.o_sb = a->o_sb; p.o_ss = a->o_ss; p.o_sh = a->o_sh;
  p.scale = a->scale;
  if (a->dropout_p > 0.f) {
    p.drop_thresh = (uint32_t)(a->dropout_p * 65536.0f + 0.5f);
    p.drop_scale = 1.0f / (1.0f - (float)p.drop_thresh / 65536.0f);
    p.drop_seed = a->dropout_seed;
    p.drop_seed_ptr = a->dropout_seed_ptr;
  }
  if (mode == 3 || (mode == 0 && a->Lq >= 96)) {
    // warp-specialised tcgen05 kernel on 256-query CTAs.  (Sending a short remainder of <= 64 rows to the mma.sync kernel
    // instead of a latency-bound CTA was measured and lost: profiles/r1_notes.md.)
    const int rc = attn_fwd_ws_dispatch(a, s, a->Lq);
    if (rc != DVLA_ERR_UNSUPPORTED) return rc;
  }
  attn_fwd_kernel<<<dim3(p.nqt, p.H, p.B), 128, 0, s>>>(p);
  DVLA_CHECK_LAUNCH("attn_fwd");
  return DVLA_OK;
}

int attn_bwd_ws_dispatch(const dvla_attn_bwd_args* a, const uint32_t* mask_t, int mask_t_words, cudaStream_t s, long long q_rows,
                         long long k_rows);    // attention_bwd_ws.cu

// 0 = auto (warp-specialised tcgen05 kernels when both sequences are >= 96 long, SIMT row kernels for Lq <= 32 without
// mask / dropout, mma.sync kernels otherwise), 1 = mma.sync kernels everywhere, 4 = warp-specialised whenever expressible
static int attn_bwd_mode() {
  static const int mode = [] {
    const char* e = getenv("DVLA_ATTN_BWD");
    return (e && !strcmp(e, "legacy")) ? 1 : (e && !strcmp(e, "ws")) ? 4 : 0;
  }();
  return mode;
}

int attn_bwd_dispatch(const dvla_attn_bwd_args* a, cudaStream_t s) {
  if (!a) { set_error("attn_bwd: null args"); return DVLA_ERR_INVALID; }
  if (!strides_ok(a->q, a->q_sb, a->q_ss, a->q_sh) || !strides_ok(a->k, a->k_sb, a->k_ss, a->k_sh) ||
      !strides_ok(a->v, a->v_sb, a->v_ss, a->v_sh) || !strides_ok(a->o, a->o_sb, a->o_ss, a->o_sh) ||
      !strides_ok(a->d_o, a->do_sb, a->do_ss, a->do_sh) || !strides_ok(a->dq, a->dq_sb, a->dq_ss, a->dq_sh) ||
      !strides_ok(a->dk, a->dk_sb, a->dk_ss, a->dk_sh) || !strides_ok(a->dv, a->dv_sb, a->dv_ss, a->dv_sh)) {
    set_error("attn_bwd: tensors must be non-null, 16-byte aligned with strides multiple of 8 elements");
    return DVLA_ERR_INVALID;
  }
  if (!a->lse || !a->delta) { set_error("attn_bwd: lse/delta null"); return DVLA_ERR_INVALID; }
  if (a->B <= 0 || a->H <= 0 || a->Lq <= 0 || a->Lk <= 0) { set_error("attn_bwd: non-positive dims"); return DVLA_ERR_INVALID; }
  if (a->mask && !a->tile_flags) { set_error("attn_bwd: mask given without tile_flags"); return DVLA_ERR_INVALID; }
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.q = (const bf16*)a->q; p.k = (const bf16*)a->k; p.v = (const bf16*)a->v; p.o = (const bf16*)a->o; p.d_o = (const bf16*)a->d_o;
  p.dq = (bf16*)a->dq; p.dk = (bf16*)a->dk; p.dv = (bf16*)a->dv;
  p.lse = const_cast<float*>(a->lse); p.delta = a->delta;
  p.mask = a->mask; p.tile_flags = a->mask ? a->tile_flags : nullptr;
  p.B = (int)a->B; p.H = (int)a->H; p.Lq = (int)a->Lq; p.Lk = (int)a->Lk;
  p.nqt = (p.Lq + 63) / 64; p.nkt = (p.Lk + 63) / 64; p.mask_words = a->mask_words;
  p.q_sb = a->q_sb; p.q_ss = a->q_ss; p.q_sh = a->q_sh; p.k_sb = a->k_sb; p.k_ss = a->k_ss; p.k_sh = a->k_sh;
  p.v_sb = a->v_sb; p.v_ss = a->v_ss; p.v_sh = a->v_sh; p.o_sb = a->o_sb; p.o_ss = a->o_ss; p.o_sh = a->o_sh;
  p.do_sb = a->do_sb; p.do_ss = a->do_ss; p.do_sh = a->do_sh;
  p.dq_sb = a->dq_sb; p.dq_ss = a->dq_ss; p.dq_sh = a->dq_sh; p.dk_sb = a->dk_sb; p.dk_ss = a->dk_ss; p.dk_sh = a->dk_sh;
  p.dv_sb = a->dv_sb; p.dv_ss = a->dv_ss; p.dv_sh = a->dv_sh;
  p.scale = a->scale;
  if (a->dropout_p > 0.f) {
    p.drop_thresh = (uint32_t)(a->dropout_p * 65536.0f + 0.5f);
    p.drop_scale = 1.0f / (1.0f - (float)p.drop_thresh / 65536.0f);
    p.drop_seed = a->dropout_seed;
    p.drop_seed_ptr = a->dropout_seed_ptr;
  }
  const int bmode = attn_bwd_mode();
  if (a->key_bias) {
    if (!attn_small_applicable(a->Lq, a->Lk, a->mask, a->dropout_p)) {
      set_error("attn_bwd: key_bias needs the short-sequence kernels (Lq <= 32, Lk <= 64, no mask, no dropout)");
      return DVLA_ERR_UNSUPPORTED;
    }
    return attn_small_bwd_dispatch(a, s);
  }
  if (bmode == 0 && attn_small_enabled() && attn_small_applicable(a->Lq, a->Lk, a->mask, a->dropout_p))
    return attn_small_bwd_dispatch(a, s);          // writes delta itself
  const long long rows = (long long)p.B * p.H * p.Lq;
  attn_delta_kernel<<<(unsigned)((rows * 8 + 255) / 256), 256, 0, s>>>(p);
  DVLA_CHECK_LAUNCH("attn_delta");
  if (bmode == 4 || (bmode == 0 && a->Lq >= 96 && a->Lk >= 96)) {
    // warp-specialised tcgen05 kernels, 2 CTAs / SM (attention_bwd_ws.cu)
    const int rc = attn_bwd_ws_dispatch(a, a->mask_t, a->mask_t_words, s, a->Lq, a->Lk);
    if (rc != DVLA_ERR_UNSUPPORTED) return rc;
  }
  attn_bwd_dkv_kernel<<<dim3(p.nkt, p.H, p.B), 128, 0, s>>>(p);
  DVLA_CHECK_LAUNCH("attn_bwd_dkv");
  attn_bwd_dq_kernel<<<dim3(p.nqt, p.H, p.B), 128, 0, s>>>(p);
  DVLA_CHECK_LAUNCH("attn_bwd_dq");
  return DVLA_OK;
}

int attn_mask_tiles_dispatch(const uint32_t* mask, int32_t mask_words, int64_t Lq, int64_t Lk, uint8_t* flags,
                             cudaStream_t s) {
  if (!mask || !flags) { set_error("attn_mask_tiles: null pointer"); return DVLA_ERR_INVALID; }
  if (mask_words * 32 < Lk) { set_error("attn_mask_tiles: mask_words too small"); return DVLA_ERR_INVALID; }
  const int nqt = (int)((Lq + 63) / 64), nkt = (int)((Lk + 63) / 64);
  const int tiles = nqt * nkt;
  attn_mask_tiles_kernel<<<(tiles + 3) / 4, 128, 0, s>>>(mask, mask_words, (int)Lq, (int)Lk, nkt, flags);
  DVLA_CHECK_LAUNCH("attn_mask_tiles");
  return DVLA_OK;
}

}  // namespace dvla
