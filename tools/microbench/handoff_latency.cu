// Micro-benchmark for the next attention round (build like tmem_mufu.cu).  One CTA per SM; warp 1 lane 0 issues MMAs, warp 2
// plays a softmax warp.  Measures the round-trip cycles of the hand-offs that make the attention kernels latency bound:
//   A. mbarrier ping-pong between two warps (arrive -> try_wait wake-up), no tensor work
//   B. MMA 128x64x64 (4 x tcgen05.mma) -> tcgen05.commit -> consumer wake-up -> arrive -> issuer wake-up
//   C. B + consumer tcgen05.ld of the 64 result columns (x32 twice) before it arrives
//   D. C + consumer writes a 128x64 bf16 tile row (8 x st.shared.v4) + fence.proxy.async before it arrives
// and the same with `busy` extra warps spinning on ex2 (issue-slot / MUFU contention as in the real kernel).
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"

using namespace dvla;

template <int MODE>
__global__ void __launch_bounds__(512, 1) handoff_kernel(int iters, int busy_warps, long long* cycles, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * 16384);
  uint64_t* full = bars;        // issuer -> consumer
  uint64_t* done = bars + 1;    // consumer -> issuer
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  volatile int* stop = reinterpret_cast<volatile int*>(bars + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 3 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(full, 1); mbar_init(done, 1); *stop = 0; fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 128);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  float acc = 0.f;
  if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, false, false);
      const uint32_t a = smem_u32(smem), b = smem_u32(smem + 16384);
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (MODE >= 1) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tmem, make_smem_desc_sw128(a + k * 32, 16, 1024), make_smem_desc_sw128(b + k * 32, 16, 1024), idesc, k != 0);
          umma_commit(full);
        } else {
          mbar_arrive(full);
        }
        mbar_wait(done, it & 1);
        tc_fence_after();
      }
      cycles[blockIdx.x] = clock64() - t0;
      *stop = 1;
    }
  } else if (warp == 2) {
    const uint32_t t_lane = tmem;     // lane quarter 2 of the accumulator (warp 2 % 4 == 2)
    uint8_t* prow = smem + 2 * 16384 + (64 + lane) * 128;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(full, it & 1);
      tc_fence_after();
      if (MODE >= 2) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(t_lane + (64u << 16), r0);
        tmem_ld_32x32(t_lane + (64u << 16) + 32, r1);
        tmem_ld_wait();
        acc += __uint_as_float(r0[lane & 31]) + __uint_as_float(r1[0]);
        if (MODE >= 3) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<uint4*>(prow + ((g ^ (lane & 7)) << 4)) = make_uint4(r0[4 * g], r0[4 * g + 1], r1[4 * g], r1[4 * g + 1]);
          fence_proxy_async_smem();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(done);
    }
  } else if (warp >= 3 && warp < 3 + busy_warps) {
    float x = 0.01f * lane;
    while (!*stop) {
#pragma unroll
      for (int i = 0; i < 32; ++i) x = ex2_approx(x * 0.5f - 1.0f);
      acc += x;
    }
  }
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

template <int MODE>
static void run(const char* name, int sms) {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, sizeof(long long) * sms); cudaMalloc(&sink, 4);
  const int smem_bytes = 3 * 16384 + 64;
  cudaFuncSetAttribute(handoff_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int iters = 2000;
  for (int busy : {0, 4, 12}) {
    handoff_kernel<MODE><<<sms, 512, smem_bytes>>>(iters, busy, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
    printf("%-44s busy_warps=%2d  cycles per round trip = %8.1f\n", name, busy, avg / iters);
  }
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs=%d\n", sms);
  run<0>("A mbarrier ping-pong", sms);
  run<1>("B 4xMMA(128x64x16) + commit -> wake -> arrive", sms);
  run<2>("C B + tcgen05.ld 64 columns", sms);
  run<3>("D C + 8x st.shared.v4 + fence.proxy.async", sms);
  return 0;
}
