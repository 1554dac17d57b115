// Whole-sampler kernel for action inference: the 10-step DDIM loop of the DiT action head with classifier-free guidance
// (reference models/dreamvla_model.py:935-987 -> action_model.py:76-83 -> gaussian_diffusion.py:522-569,609-689 ->
// models.py:234-268) as ONE persistent cooperative kernel.
//
// Why: at batch 1 the sampler is 10 x (12 DiT blocks + embedders + final layer + DDIM algebra) ~ 1100 kernel launches on
// 12 token rows (2 guidance branches x 6 tokens) -- every launch is latency, none is work: the whole DiT-B is 170 MB of bf16
// weights, i.e. ~30 us per pass at HBM speed.  Here one CTA per SM stays resident for the entire sample; every linear layer
// is a "skinny" GEMM (M <= 24 rows) in which a warp owns one output column, streams that column's weight row from HBM/L2
// with 16-byte loads and keeps the activations (bf16) in shared memory; the phases of a block (LN+QKV | attention+proj+res |
// LN+fc1+GELU | fc2+res) are separated by grid barriers (cooperative launch) instead of kernel boundaries.  Timestep and
// condition embeddings are hoisted out of the loop (they do not depend on x); x itself (T x 7 numbers per sample) lives in
// shared memory of every CTA and is advanced redundantly, so no barrier is needed for the final layer / guidance / DDIM update.
// Buffers exchanged between CTAs (h, qkv, u, te, ze) are read with ld.global.cg: L1 is not coherent across SMs.
//
// Numerics follow the module path where it rounds: GEMM inputs are bf16 (activations are rounded when staged in shared
// memory), accumulation fp32; the residual stream, LayerNorm statistics and softmax stay fp32 (the module path keeps the
// residual stream in bf16), so results agree with it to bf16 rounding, not bit for bit (tests/test_rollout_gpu.py).
#include <cooperative_groups.h>

#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "../../include/dvla.h"

namespace cg = cooperative_groups;

namespace dvla {
void set_error(const char* fmt, ...);
void count_launch();
int num_sms();

namespace {

constexpr int DS_THREADS = 256;
constexpr int DS_WARPS = DS_THREADS / 32;
constexpr int DS_MAXM = 24;          // rows: 2 guidance branches x bs x 2T tokens
constexpr int DS_MAX_DEPTH = 28;
constexpr int DS_MAX_STEPS = 16;

struct DitBlockW { const bf16 *qkv_w, *qkv_b, *proj_w, *proj_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b; };
struct DitSamplerParams {
  DitBlockW blk[DS_MAX_DEPTH];
  int depth;
  const bf16 *x_w, *x_b, *t0_w, *t0_b, *t2_w, *t2_b, *z_w, *z_b, *uncond, *pos, *fin_w, *fin_b;
  const bf16* z;            // [bs, T, token]
  const float* noise;       // [bs, T, C]
  float* out;               // [bs, T, C]
  int tmap[DS_MAX_STEPS];
  float sqrt_recip[DS_MAX_STEPS], sqrt_recipm1[DS_MAX_STEPS], acp_prev[DS_MAX_STEPS];
  int n_steps;
  float cfg_scale;
  int bs, T, C, H, heads, token, mlp, freq;
  float *h0, *h1, *qkv, *u, *te, *t1, *ze;     // fp32 scratch (global, L2 resident)
};

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}

// out(m, n) = sum_k Xs[m][k] * W[n][k] (+ bias[n]) for every column n owned by this warp; Xs: shared, bf16, [M][K], K % 8 == 0.
// epi(m, n, value) is called by lane m (m < M) once per owned column.  A warp works on TWO columns at a time (the staged
// activations are read once for both) and issues the weight loads of four 8-element chunks per column before using them:
// the kernel is a weight-streaming loop whose only latency hiding is loads in flight.
template <typename Epi>
__device__ __forceinline__ void skinny_gemm(const bf16* __restrict__ Xs, int M, int K, const bf16* __restrict__ W,
                                            const bf16* __restrict__ bias, int N, Epi epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * DS_WARPS + warp, nw = gridDim.x * DS_WARPS;
  const int chunks = K >> 3;
  for (int n0 = gw; n0 < N; n0 += 2 * nw) {
    const int n1 = n0 + nw;
    const bool two = n1 < N;
    float acc0[DS_MAXM], acc1[DS_MAXM];
#pragma unroll
    for (int m = 0; m < DS_MAXM; ++m) { acc0[m] = 0.f; acc1[m] = 0.f; }
    const uint4* w0 = reinterpret_cast<const uint4*>(W + static_cast<long long>(n0) * K);
    const uint4* w1 = reinterpret_cast<const uint4*>(W + static_cast<long long>(two ? n1 : n0) * K);
    for (int c0 = lane; c0 < chunks; c0 += 128) {
      uint4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + 32 * u;
        a[u] = c < chunks ? __ldg(w0 + c) : make_uint4(0, 0, 0, 0);
        b[u] = c < chunks ? __ldg(w1 + c) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + 32 * u;
        if (c < chunks) {
          float wa[8], wb[8];
          unpack8(a[u], wa);
          unpack8(b[u], wb);
#pragma unroll
          for (int m = 0; m < DS_MAXM; ++m) {
            if (m < M) {
              float x[8];
              unpack8(*reinterpret_cast<const uint4*>(Xs + m * K + c * 8), x);
              float s0 = 0.f, s1 = 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) { s0 = fmaf(x[j], wa[j], s0); s1 = fmaf(x[j], wb[j], s1); }
              acc0[m] += s0;
              acc1[m] += s1;
            }
          }
        }
      }
    }
    float mine0 = 0.f, mine1 = 0.f;
#pragma unroll
    for (int m = 0; m < DS_MAXM; ++m) {
      if (m < M) {
        const float t0 = warp_sum_f(acc0[m]), t1 = warp_sum_f(acc1[m]);
        if (lane == m) { mine0 = t0; mine1 = t1; }
      }
    }
    if (lane < M) {
      epi(lane, n0, mine0 + (bias ? __bfloat162float(bias[n0]) : 0.f));
      if (two) epi(lane, n1, mine1 + (bias ? __bfloat162float(bias[n1]) : 0.f));
    }
  }
}

// Xs[m][:] = bf16(LayerNorm(h[m][:])) without affine, eps 1e-6 (timm Block with elementwise_affine=False), warp per row
__device__ __forceinline__ void stage_layernorm(bf16* Xs, const float* __restrict__ h, int M, int H) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int m = warp; m < M; m += DS_WARPS) {
    const float* row = h + static_cast<long long>(m) * H;
    float s = 0.f;
    for (int c = lane; c < H; c += 32) s += __ldcg(row + c);
    const float mean = warp_sum_f(s) / H;
    float v = 0.f;
    for (int c = lane; c < H; c += 32) { const float d = __ldcg(row + c) - mean; v += d * d; }
    const float rstd = rsqrtf(warp_sum_f(v) / H + 1e-6f);
    for (int c = lane; c < H; c += 32) Xs[m * H + c] = __float2bfloat16((__ldcg(row + c) - mean) * rstd);
  }
}
__device__ __forceinline__ void stage_rows(bf16* Xs, const float* __restrict__ src, int n) {   // fp32 global -> bf16 shared, n % 4 == 0
  const float4* s4 = reinterpret_cast<const float4*>(src);
  uint2* d2 = reinterpret_cast<uint2*>(Xs);
  const int n4 = n >> 2;
  for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * DS_THREADS) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * DS_THREADS;
      v[u] = i < n4 ? __ldcg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * DS_THREADS;
      if (i < n4) d2[i] = make_uint2(pack_bf16x2(v[u].x, v[u].y), pack_bf16x2(v[u].z, v[u].w));
    }
  }
}

__global__ void __launch_bounds__(DS_THREADS, 1) dit_ddim_sample_kernel(const __grid_constant__ DitSamplerParams p) {
  extern __shared__ __align__(16) uint8_t ds_smem[];
  cg::grid_group grid = cg::this_grid();
  const int H = p.H, T = p.T, C = p.C, bs = p.bs;
  const int L = 2 * T;                 // tokens per sequence: T condition tokens + T action tokens (models.py:240-244)
  const int nseq = 2 * bs;             // guidance: sequences [0, bs) conditional, [bs, 2bs) unconditional (models.py:253-257)
  const int M = nseq * L;
  bf16* Xs = reinterpret_cast<bf16*>(ds_smem);                                // staged GEMM input, up to [M][mlp]
  float* xcur = reinterpret_cast<float*>(ds_smem + static_cast<size_t>(M) * p.mlp * 2);     // [bs][T][C] current sample
  float* eps_s = xcur + bs * T * C;                                           // [nseq][T][C] final-layer outputs
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gtid = blockIdx.x * DS_THREADS + tid, gthreads = gridDim.x * DS_THREADS;

  // ---- prologue A: timestep-frequency features -> t1 = silu(W0 f + b0) for all steps;  ze = Wz z + bz for the bs*T condition
  //      rows and for the `uncondition` vector (row bs*T)
  for (int i = tid; i < p.n_steps * p.freq; i += DS_THREADS) {
    const int s = i / p.freq, k = i - s * p.freq, half = p.freq >> 1;
    const float fr = __expf(-9.210340371976184f * static_cast<float>(k < half ? k : k - half) / half);   // ln(10000)
    const float a = static_cast<float>(p.tmap[s]) * fr;
    Xs[i] = __float2bfloat16(k < half ? cosf(a) : sinf(a));
  }
  __syncthreads();
  skinny_gemm(Xs, p.n_steps, p.freq, p.t0_w, p.t0_b, H, [&](int m, int n, float v) {
    p.t1[m * H + n] = v / (1.0f + __expf(-v));
  });
  __syncthreads();
  {
    const int rows = bs * T + 1;
    for (int i = tid; i < rows * p.token; i += DS_THREADS) {
      const int r = i / p.token, k = i - r * p.token;
      Xs[i] = r < bs * T ? p.z[static_cast<long long>(r) * p.token + k] : p.uncond[k];
    }
    __syncthreads();
    skinny_gemm(Xs, rows, p.token, p.z_w, p.z_b, H, [&](int m, int n, float v) { p.ze[m * H + n] = v; });
  }
  for (int i = tid; i < bs * T * C; i += DS_THREADS) xcur[i] = __bfloat162float(__float2bfloat16(p.noise[i]));
  grid.sync();
  // ---- prologue B: te = W2 t1 + b2
  stage_rows(Xs, p.t1, p.n_steps * H);
  __syncthreads();
  skinny_gemm(Xs, p.n_steps, H, p.t2_w, p.t2_b, H, [&](int m, int n, float v) { p.te[m * H + n] = v; });
  grid.sync();

  for (int step = p.n_steps - 1, it = 0; step >= 0; --step, ++it) {
    float* h = (it & 1) ? p.h1 : p.h0;
    // ---- H0: token rows.  row = seq * L + tok; tok < T: te + ze (+ pos); tok >= T: x_embedder(x) (+ pos)
    for (int i = gtid; i < M * H; i += gthreads) {
      const int row = i / H, c = i - row * H;
      const int seq = row / L, tok = row - seq * L;
      const int b = seq < bs ? seq : seq - bs;
      float v;
      if (tok < T) {
        const int zr = seq < bs ? b * T + tok : bs * T;
        v = __ldcg(p.te + step * H + c) + __ldcg(p.ze + zr * H + c);
      } else {
        v = __bfloat162float(p.x_b[c]);
        const float* xr = xcur + (b * T + tok - T) * C;
        for (int k = 0; k < C; ++k)
          v = fmaf(__bfloat162float(__float2bfloat16(xr[k])), __bfloat162float(p.x_w[c * C + k]), v);
      }
      h[i] = v + __bfloat162float(p.pos[tok * H + c]);
    }
    grid.sync();
    for (int l = 0; l < p.depth; ++l) {
      const DitBlockW& w = p.blk[l];
      // ---- P1: qkv = LN(h) Wqkv^T + b
      stage_layernorm(Xs, h, M, H);
      __syncthreads();
      skinny_gemm(Xs, M, H, w.qkv_w, w.qkv_b, 3 * H, [&](int m, int n, float v) { p.qkv[m * 3 * H + n] = v; });
      grid.sync();
      // ---- P2: attention (every CTA computes all (sequence, head) pairs: 6x6 scores each) -> O in shared; h += O Wproj^T + b
      {
        bf16* Os = Xs;                                                     // [M][H] bf16
        const float scale = 0.125f;                                       // head_dim 64
        for (int pr = warp; pr < nseq * p.heads; pr += DS_WARPS) {
          const int seq = pr / p.heads, hd = pr - seq * p.heads;
          float q[8][2], k[8][2], v[8][2];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (i < L) {
              const float* r = p.qkv + static_cast<long long>(seq * L + i) * 3 * H + hd * 64 + 2 * lane;
              q[i][0] = __ldcg(r); q[i][1] = __ldcg(r + 1);
              k[i][0] = __ldcg(r + H); k[i][1] = __ldcg(r + H + 1);
              v[i][0] = __ldcg(r + 2 * H); v[i][1] = __ldcg(r + 2 * H + 1);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (i < L) {
              float s[8], mx = -INFINITY;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (j < L) {
                  s[j] = warp_sum_f(q[i][0] * k[j][0] + q[i][1] * k[j][1]) * scale;
                  mx = fmaxf(mx, s[j]);
                }
              }
              float den = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (j < L) {
                  const float e = __expf(s[j] - mx);
                  den += e; o0 = fmaf(e, v[j][0], o0); o1 = fmaf(e, v[j][1], o1);
                }
              }
              const float inv = 1.0f / den;
              *reinterpret_cast<uint32_t*>(Os + (seq * L + i) * H + hd * 64 + 2 * lane) = pack_bf16x2(o0 * inv, o1 * inv);
            }
          }
        }
        __syncthreads();
        skinny_gemm(Os, M, H, w.proj_w, w.proj_b, H, [&](int m, int n, float v) { __stcg(h + m * H + n, __ldcg(h + m * H + n) + v); });
      }
      grid.sync();
      // ---- P3: u = gelu_tanh(LN(h) Wfc1^T + b)
      stage_layernorm(Xs, h, M, H);
      __syncthreads();
      skinny_gemm(Xs, M, H, w.fc1_w, w.fc1_b, p.mlp, [&](int m, int n, float v) { p.u[m * p.mlp + n] = gelu_tanh_f(v); });
      grid.sync();
      // ---- P4: h += u Wfc2^T + b
      stage_rows(Xs, p.u, M * p.mlp);
      __syncthreads();
      skinny_gemm(Xs, M, p.mlp, w.fc2_w, w.fc2_b, H, [&](int m, int n, float v) { __stcg(h + m * H + n, __ldcg(h + m * H + n) + v); });
      grid.sync();
    }
    // ---- final layer on the action rows + guidance + DDIM update: redundantly in every CTA (x stays in shared memory).
    //      h of this step is not written again before two more grid barriers (the next step uses the other h buffer).
    {
      const int arows = nseq * T;
      for (int r = warp; r < arows; r += DS_WARPS) {           // Xs[r][:] = bf16(LN(h[action row r]))
        const int seq = r / T, tok = T + (r - seq * T);
        const float* row = h + static_cast<long long>(seq * L + tok) * H;
        float s = 0.f;
        for (int c = lane; c < H; c += 32) s += __ldcg(row + c);
        const float mean = warp_sum_f(s) / H;
        float vv = 0.f;
        for (int c = lane; c < H; c += 32) { const float d = __ldcg(row + c) - mean; vv += d * d; }
        const float rstd = rsqrtf(warp_sum_f(vv) / H + 1e-6f);
        for (int c = lane; c < H; c += 32) Xs[r * H + c] = __float2bfloat16((__ldcg(row + c) - mean) * rstd);
      }
      __syncthreads();
      for (int o = warp; o < arows * C; o += DS_WARPS) {       // eps_s[r][cc] = bf16(LN row . fin_w[cc] + fin_b[cc])
        const int r = o / C, cc = o - r * C;
        float s = 0.f;
        for (int c = lane; c < H; c += 32) s = fmaf(__bfloat162float(Xs[r * H + c]), __bfloat162float(p.fin_w[cc * H + c]), s);
        s = warp_sum_f(s);
        if (lane == 0) eps_s[o] = __bfloat162float(__float2bfloat16(s + __bfloat162float(p.fin_b[cc])));
      }
      __syncthreads();
      for (int i = tid; i < bs * T * C; i += DS_THREADS) {
        const int b = i / (T * C), rem = i - b * T * C;
        const float cond = eps_s[b * T * C + rem], unc = eps_s[(bs + b) * T * C + rem];
        const float eps = __bfloat162float(__float2bfloat16(unc + p.cfg_scale * (cond - unc)));     // models.py:262-266
        const float x = xcur[i];
        const float x0 = p.sqrt_recip[step] * x - p.sqrt_recipm1[step] * eps;                       // gaussian_diffusion.py:328,345
        const float e2 = (p.sqrt_recip[step] * x - x0) / p.sqrt_recipm1[step];                      // :554
        const float ab = p.acp_prev[step];
        xcur[i] = x0 * sqrtf(ab) + sqrtf(1.0f - ab) * e2;                                            // :560-565, eta = 0
      }
      __syncthreads();
    }
  }
  if (blockIdx.x == 0)
    for (int i = tid; i < bs * T * C; i += DS_THREADS) p.out[i] = xcur[i];
}

}  // namespace

int dit_ddim_sample_dispatch(const dvla_dit_sampler_args* a, cudaStream_t s) {
  if (!a || !a->blocks || !a->z || !a->noise || !a->out || !a->workspace) { set_error("dit_ddim_sample: null pointer"); return DVLA_ERR_INVALID; }
  const int H = (int)a->hidden, T = (int)a->T, C = (int)a->channels, bs = (int)a->batch;
  const int M = 2 * bs * 2 * T;
  if (a->depth <= 0 || a->depth > DS_MAX_DEPTH || a->n_steps <= 0 || a->n_steps > DS_MAX_STEPS || M > DS_MAXM || 2 * T > 8 ||
      a->heads * 64 != H || H % 8 || a->token % 8 || a->mlp % 8 || a->freq % 8 || bs * T + 1 > DS_MAXM) {
    set_error("dit_ddim_sample: unsupported shape (rows %d <= %d, tokens per sequence <= 8, head_dim 64, dims %% 8)", M, DS_MAXM);
    return DVLA_ERR_UNSUPPORTED;
  }
  if (dvla_dit_sampler_workspace_bytes(a->batch, a->T, a->hidden, a->mlp, a->n_steps) > a->workspace_bytes) {
    set_error("dit_ddim_sample: workspace too small"); return DVLA_ERR_INVALID;
  }
  DitSamplerParams p;
  memset(&p, 0, sizeof(p));
  for (int l = 0; l < a->depth; ++l) {
    const dvla_dit_block_weights& b = a->blocks[l];
    p.blk[l] = DitBlockW{(const bf16*)b.qkv_w, (const bf16*)b.qkv_b, (const bf16*)b.proj_w, (const bf16*)b.proj_b,
                         (const bf16*)b.fc1_w, (const bf16*)b.fc1_b, (const bf16*)b.fc2_w, (const bf16*)b.fc2_b};
  }
  p.depth = (int)a->depth;
  p.x_w = (const bf16*)a->x_w; p.x_b = (const bf16*)a->x_b; p.t0_w = (const bf16*)a->t0_w; p.t0_b = (const bf16*)a->t0_b;
  p.t2_w = (const bf16*)a->t2_w; p.t2_b = (const bf16*)a->t2_b; p.z_w = (const bf16*)a->z_w; p.z_b = (const bf16*)a->z_b;
  p.uncond = (const bf16*)a->uncondition; p.pos = (const bf16*)a->pos; p.fin_w = (const bf16*)a->final_w; p.fin_b = (const bf16*)a->final_b;
  p.z = (const bf16*)a->z; p.noise = a->noise; p.out = a->out;
  for (int i = 0; i < a->n_steps; ++i) {
    p.tmap[i] = a->timestep_map[i]; p.sqrt_recip[i] = a->sqrt_recip_alphas_cumprod[i];
    p.sqrt_recipm1[i] = a->sqrt_recipm1_alphas_cumprod[i]; p.acp_prev[i] = a->alphas_cumprod_prev[i];
  }
  p.n_steps = (int)a->n_steps; p.cfg_scale = a->cfg_scale;
  p.bs = bs; p.T = T; p.C = C; p.H = H; p.heads = (int)a->heads; p.token = (int)a->token; p.mlp = (int)a->mlp; p.freq = (int)a->freq;
  float* ws = reinterpret_cast<float*>(a->workspace);
  p.h0 = ws; ws += M * H;
  p.h1 = ws; ws += M * H;
  p.qkv = ws; ws += M * 3 * H;
  p.u = ws; ws += M * p.mlp;
  p.te = ws; ws += p.n_steps * H;
  p.t1 = ws; ws += p.n_steps * H;
  p.ze = ws; ws += (bs * T + 1) * H;
  int widest = p.mlp > p.token ? p.mlp : p.token;
  if (p.freq > widest) widest = p.freq;
  int rows = M > bs * T + 1 ? M : bs * T + 1;
  if (p.n_steps > rows) rows = p.n_steps;
  size_t smem = static_cast<size_t>(rows) * widest * 2;
  const size_t xs_bytes = static_cast<size_t>(M) * p.mlp * 2;          // xcur / eps_s sit behind the [M][mlp] staging area
  if (smem < xs_bytes) smem = xs_bytes;
  smem += static_cast<size_t>(bs * T * C + 2 * bs * T * C) * sizeof(float) + 16;
  if (static_cast<size_t>(bs * T + 1) * p.token * 2 > xs_bytes || static_cast<size_t>(p.n_steps) * (p.freq > H ? p.freq : H) * 2 > xs_bytes) {
    set_error("dit_ddim_sample: prologue staging does not fit behind the block staging area"); return DVLA_ERR_UNSUPPORTED;
  }
  if (smem > 227 * 1024) { set_error("dit_ddim_sample: %zu bytes of shared memory needed", smem); return DVLA_ERR_UNSUPPORTED; }
  static const cudaError_t attr_err = cudaFuncSetAttribute(dit_ddim_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (attr_err != cudaSuccess) { set_error("dit_ddim_sample smem attr: %s", cudaGetErrorString(attr_err)); return DVLA_ERR_CUDA; }
  int dev = 0, coop = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  if (!coop) { set_error("dit_ddim_sample: device has no cooperative launch"); return DVLA_ERR_UNSUPPORTED; }
  static const int env_ctas = [] { const char* v = getenv("DVLA_DIT_CTAS"); return v ? atoi(v) : 0; }();
  int ctas = num_sms();
  if (env_ctas > 0 && env_ctas < ctas) ctas = env_ctas;
  void* args[] = {const_cast<DitSamplerParams*>(&p)};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(dit_ddim_sample_kernel), dim3(ctas), dim3(DS_THREADS), args, smem, s);
  if (e != cudaSuccess) { set_error("dit_ddim_sample launch: %s", cudaGetErrorString(e)); return DVLA_ERR_CUDA; }
  count_launch();
  return DVLA_OK;
}

}  // namespace dvla

extern "C" int64_t dvla_dit_sampler_workspace_bytes(int64_t batch, int64_t T, int64_t hidden, int64_t mlp, int64_t n_steps) {
  const int64_t M = 2 * batch * 2 * T;
  return 4 * (2 * M * hidden + M * 3 * hidden + M * mlp + 2 * n_steps * hidden + (batch * T + 1) * hidden);
}
