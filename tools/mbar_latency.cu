// Latency of mbarrier operations on sm_100a, one warp (optionally with other warps hammering the same SM).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/mbar_latency.bin tools/mbar_latency.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(long long* out, int mode) {
  __shared__ uint64_t bar[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar[0])) : "memory");   // phase 0 of bar0 complete
  }
  __syncthreads();
  if (warp == 0) {
    const int N = 200;
    uint32_t acc = 0, ok_prev = 1;
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) {
      uint32_t ok;
      // dependent chain: the address of the next probe depends on the previous result
      asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2; selp.b32 %0, 1, 0, P;}"
                   : "=r"(ok) : "r"(s32(&bar[0]) + ((ok_prev ^ 1u) << 3)), "r"(0) : "memory");
      acc += ok; ok_prev = ok;
    }
    long long t1 = clock64();
    if (lane == 0) { out[0] = (t1 - t0) / N; out[1] = acc; }
    // test_wait
    acc = 0; t0 = clock64();
    for (int i = 0; i < N; ++i) {
      uint32_t ok;
      asm volatile("{.reg .pred P; mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2; selp.b32 %0, 1, 0, P;}"
                   : "=r"(ok) : "r"(s32(&bar[0]) + ((ok_prev ^ 1u) << 3)), "r"(0) : "memory");
      acc += ok; ok_prev = ok;
    }
    t1 = clock64();
    if (lane == 0) { out[2] = (t1 - t0) / N; out[3] = acc; }
    // arrive (lane 0) followed by a satisfied try_wait on ANOTHER barrier
    acc = 0; t0 = clock64();
    for (int i = 0; i < N; ++i) {
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar[1])) : "memory");
      uint32_t ok;
      asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2; selp.b32 %0, 1, 0, P;}"
                   : "=r"(ok) : "r"(s32(&bar[0]) + ((ok_prev ^ 1u) << 3)), "r"(0) : "memory");
      acc += ok; ok_prev = ok;
    }
    t1 = clock64();
    if (lane == 0) { out[4] = (t1 - t0) / N; out[5] = acc; }
    // plain shared-memory load chain for scale
    volatile uint64_t* vb = bar;
    acc = 0; t0 = clock64();
    for (int i = 0; i < N; ++i) { uint32_t v = (uint32_t)vb[(ok_prev ^ 1u)]; acc += v; ok_prev = (v | 1u) & 1u; }
    t1 = clock64();
    if (lane == 0) { out[6] = (t1 - t0) / N; out[7] = acc; }
  } else if (mode == 1) {
    // other warps poll a never-completing barrier phase (like waiting control / softmax warps)
    uint32_t ok = 0; int spins = 0;
    while (!ok && spins < 4000) {
      asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2; selp.b32 %0, 1, 0, P;}"
                   : "=r"(ok) : "r"(s32(&bar[2])), "r"(0) : "memory");
      ++spins;
    }
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  for (int mode = 0; mode < 2; ++mode) {
    k<<<1, mode == 0 ? 32 : 352>>>(d, mode);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("%s (%s): try_wait(satisfied) %lld clk, test_wait %lld clk, arrive+try_wait %lld clk, ld.shared chain %lld clk\n",
           mode == 0 ? "one warp alone" : "with 10 polling warps", cudaGetErrorString(e), h[0], h[2], h[4], h[6]);
  }
  return 0;
}
