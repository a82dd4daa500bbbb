// Microbenchmark of tcgen05 / mbarrier primitive latencies on sm_100a (design input for the attention pipeline).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hsenet_b200/csrc -o /tmp/umma_latency tools/umma_latency.cu
#include <cstdio>
#include "common.cuh"
using namespace hs;

__global__ void __launch_bounds__(128, 1) lat_kernel(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(&tmem_slot, 256); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem));
  const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem + 16384));
  uint32_t ph0 = 0, ph1 = 0;
  if (warp == 0 && lane == 0) {
    // warm up
    umma_ss(tm, adesc, bdesc, make_idesc_bf16(128, 64, 0, 0), 0); tc_commit(&bar[0]); mbar_wait(&bar[0], ph0); ph0 ^= 1;
    for (int rep = 0; rep < 4; ++rep) {
      // T0: commit only
      long long t0 = clock64();
      tc_commit(&bar[0]); mbar_wait(&bar[0], ph0); ph0 ^= 1;
      long long t1 = clock64();
      // T1: 1 MMA N=64 + commit
      umma_ss(tm, adesc, bdesc, make_idesc_bf16(128, 64, 0, 0), 0); tc_commit(&bar[0]); mbar_wait(&bar[0], ph0); ph0 ^= 1;
      long long t2 = clock64();
      // T2: 4 MMA N=64 + commit
      for (int k = 0; k < 4; ++k) umma_ss(tm, adesc + 2 * k, bdesc + 2 * k, make_idesc_bf16(128, 64, 0, 0), k);
      tc_commit(&bar[0]); mbar_wait(&bar[0], ph0); ph0 ^= 1;
      long long t3 = clock64();
      // T3: 4 MMA N=128 + commit
      for (int k = 0; k < 4; ++k) umma_ss(tm, adesc + 2 * k, bdesc + 2 * k, make_idesc_bf16(128, 128, 0, 0), k);
      tc_commit(&bar[0]); mbar_wait(&bar[0], ph0); ph0 ^= 1;
      long long t4 = clock64();
      // T4: 16 MMA N=64 + commit (issue rate)
      for (int k = 0; k < 16; ++k) umma_ss(tm, adesc + 2 * (k & 3), bdesc + 2 * (k & 3), make_idesc_bf16(128, 64, 0, 0), k);
      long long t4i = clock64();
      tc_commit(&bar[0]); mbar_wait(&bar[0], ph0); ph0 ^= 1;
      long long t5 = clock64();
      // T5: 8 TS MMA (A from TMEM) N=64 + commit
      for (int k = 0; k < 8; ++k) umma_ts(tm + 64, tm + 128 + 8 * (k & 3), bdesc + 128 * (k & 3), make_idesc_bf16(128, 64, 0, 1), k);
      tc_commit(&bar[0]); mbar_wait(&bar[0], ph0); ph0 ^= 1;
      long long t6 = clock64();
      // T6: ping-pong with warp 2: arrive bar1 -> (warp2 waits bar1, arrives bar2) -> wait bar2, 8 round trips
      for (int i = 0; i < 8; ++i) { mbar_arrive(&bar[1]); mbar_wait(&bar[2], ph1); ph1 ^= 1; }
      long long t7 = clock64();
      // T7: MMA + commit -> warp 2 waits bar3 then arrives bar2 -> wait (8 round trips)
      for (int i = 0; i < 8; ++i) {
        umma_ss(tm, adesc, bdesc, make_idesc_bf16(128, 64, 0, 0), 0); tc_commit(&bar[3]);
        mbar_wait(&bar[2], ph1); ph1 ^= 1;
      }
      long long t8 = clock64();
      if (rep == 3) {
        out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4; out[5] = t4i - t4;
        out[6] = t6 - t5; out[7] = (t7 - t6) / 8; out[8] = (t8 - t7) / 8;
      }
    }
  } else if (warp == 2) {
    uint32_t p1 = 0, p3 = 0;
    for (int rep = 0; rep < 4; ++rep) {
      for (int i = 0; i < 8; ++i) { mbar_wait(&bar[1], p1); p1 ^= 1; __syncwarp(); if (lane == 0) mbar_arrive(&bar[2]); __syncwarp(); }
      for (int i = 0; i < 8; ++i) { mbar_wait(&bar[3], p3); p3 ^= 1; tc_fence_after(); __syncwarp(); if (lane == 0) mbar_arrive(&bar[2]); __syncwarp(); }
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tm, 256); }
}

int main() {
  long long* d; cudaMalloc(&d, 16 * sizeof(long long)); cudaMemset(d, 0, 128);
  cudaFuncSetAttribute(lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024);
  lat_kernel<<<1, 128, 34 * 1024>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("status %s\n", cudaGetErrorString(e));
  const char* names[] = {"commit only -> wait", "1 MMA(N=64)+commit -> wait", "4 MMA(N=64)+commit -> wait", "4 MMA(N=128)+commit -> wait",
                         "16 MMA(N=64)+commit -> wait", "  of which issue of 16 MMA", "8 TS-MMA(N=64)+commit -> wait",
                         "mbarrier ping-pong round trip", "MMA+commit -> other warp -> arrive round trip"};
  for (int i = 0; i < 9; ++i) printf("%-48s %6lld cycles\n", names[i], h[i]);
  return 0;
}
