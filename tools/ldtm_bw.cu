// TMEM read/write throughput microbenchmark (sm_100a): cycles per tcgen05.ld/st 32x32b.x32 (4 KB per warp).
#include <cstdio>
#include "common.cuh"
using namespace hs;
__global__ void __launch_bounds__(128, 1) k(long long* out, int nwarps_active) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot + (static_cast<uint32_t>(warp * 32) << 16);
  uint32_t x[32];
  for (int i = 0; i < 32; ++i) x[i] = i;
  tmem_st32(tm, x); tmem_st32(tm + 32, x); tmem_st32(tm + 64, x); tmem_st32(tm + 96, x); tmem_st_wait();
  __syncthreads();
  long long t0 = 0, t1 = 0, t2 = 0;
  uint32_t acc = 0;
  if (warp < nwarps_active) {
    t0 = clock64();
    for (int it = 0; it < 64; ++it) {
      tmem_ld32(tm + (it & 3) * 32, x);
      tmem_ld_wait();
      acc += x[it & 31];
    }
    t1 = clock64();
    for (int it = 0; it < 16; ++it) {     // 4 loads in flight before the wait (as the attention kernel does)
      uint32_t a[32], b[32], c[32], d[32];
      tmem_ld32(tm, a); tmem_ld32(tm + 32, b); tmem_ld32(tm + 64, c); tmem_ld32(tm + 96, d);
      tmem_ld_wait();
      acc += a[it & 31] + b[it & 31] + c[it & 31] + d[it & 31];
    }
    t2 = clock64();
  }
  __syncthreads();
  if (lane == 0) { out[warp * 4 + 0] = t1 - t0; out[warp * 4 + 1] = t2 - t1; out[warp * 4 + 2] = acc; }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 256); }
}
int main() {
  long long* d; cudaMalloc(&d, 64 * sizeof(long long));
  for (int nw = 1; nw <= 4; nw *= 2) {
    cudaMemset(d, 0, 512);
    k<<<1, 128>>>(d, nw);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("active warps %d (%s): serial ld+wait: %lld cycles per x32 load (4 KB); 4-deep: %lld cycles per 4 loads (16 KB)\n", nw,
           cudaGetErrorString(e), h[0] / 64, h[1] / 16);
  }
  return 0;
}
