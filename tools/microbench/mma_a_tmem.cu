// Probe for the next attention round: tcgen05.mma with the A operand in TENSOR MEMORY (what lets the softmax warps hand P to
// the P.V MMA without the st.shared + fence.proxy.async round trip).  Computes D[128,64] = A[128,64] . B[64,64]^T with
//   A written by tcgen05.st as packed bf16 pairs (thread = row, 32-bit column j holds A[row][2j], A[row][2j+1]),
//   B in shared memory, K-major, 128B swizzle (the layout the attention kernels already use for K / P tiles),
// and checks it against the host.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Idreamvla_b200/csrc
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "common.cuh"

using namespace dvla;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
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
// D[tmem] (+)= A[tmem] . B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                 int a_cols_per_kstep) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 256);
  if (tid < 64) {                                   // B[n = tid][k]: 64 bf16 = 128 B per row, chunk c at c ^ (n & 7)
    for (int c = 0; c < 8; ++c) {
      uint32_t w[4];
      for (int j = 0; j < 4; ++j) w[j] = pack_bf16x2(B[tid * 64 + c * 8 + 2 * j], B[tid * 64 + c * 8 + 2 * j + 1]);
      *reinterpret_cast<uint4*>(smem + tid * 128 + ((c ^ (tid & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
  uint32_t r[32];
  for (int j = 0; j < 32; ++j) r[j] = pack_bf16x2(A[tid * 64 + 2 * j], A[tid * 64 + 2 * j + 1]);
  tmem_st32(tmem + 128 + lane_off, r);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, false, false);
    const uint32_t b = smem_u32(smem);
    for (int k = 0; k < 4; ++k)
      umma_bf16_ts(tmem, tmem + 128 + k * a_cols_per_kstep, make_smem_desc_sw128(b + k * 32, 16, 1024), idesc, k != 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t o0[32], o1[32];
  tmem_ld_32x32(tmem + lane_off, o0);
  tmem_ld_32x32(tmem + lane_off + 32, o1);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) { D[tid * 64 + j] = __uint_as_float(o0[j]); D[tid * 64 + 32 + j] = __uint_as_float(o1[j]); }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int main() {
  std::vector<float> A(128 * 64), B(64 * 64), D(128 * 64), R(128 * 64);
  for (int i = 0; i < 128 * 64; ++i) A[i] = (float)((i * 7 + (i >> 6)) % 9 - 4);          // small integers: exact in bf16/fp32
  for (int i = 0; i < 64 * 64; ++i) B[i] = (float)((i * 5 + (i >> 5)) % 7 - 3);
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) { float s = 0; for (int k = 0; k < 64; ++k) s += A[m * 64 + k] * B[n * 64 + k]; R[m * 64 + n] = s; }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 + 64);
  for (int cols : {8, 16}) {               // 32-bit TMEM columns of A consumed per K=16 step: 8 if two bf16 share a column
    cudaMemset(dD, 0, D.size() * 4);
    probe<<<1, 128, 8192 + 64>>>(dA, dB, dD, cols);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cols/kstep=%d: %s\n", cols, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0; for (size_t i = 0; i < D.size(); ++i) err = fmax(err, fabs(D[i] - R[i]));
    printf("A in TMEM, %2d columns per K=16 step: max |D - ref| = %g  %s\n", cols, err, err == 0 ? "EXACT" : "mismatch");
  }
  return 0;
}
