// tcgen05.mma rate on sm_100a as a function of shape, operand source / layout and accumulator dependence.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hsenet_b200/csrc -o /tmp/umma_tp tools/umma_throughput.cu && /tmp/umma_tp
// One CTA, one issuing thread, REP MMAs back to back, then one commit + wait.  "dep" accumulates every MMA into the
// same TMEM tile (what Q K^T's four k-slices and P V's key slices do), "alt2"/"alt4" rotate over 2 / 4 tiles.
#include <cstdio>
#include "common.cuh"
using namespace hs;

constexpr int REP = 256;

template <int N, int TS, int BMN, int NACC>
__device__ __forceinline__ void run_case(uint32_t tm, uint64_t adesc, uint64_t bdesc, uint64_t* bar, uint32_t& ph,
                                         long long* out) {
  constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, BMN);
  for (int pass = 0; pass < 2; ++pass) {              // pass 0 = warm-up
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REP; i += 4 * NACC) {
#pragma unroll
      for (int a = 0; a < NACC; ++a) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {                 // accumulator rotates fastest: consecutive MMAs hit different tiles
          const uint32_t d = tm + static_cast<uint32_t>(((k + a) % NACC) * N);
          const uint64_t bd = bdesc + (BMN ? 128 * k : 2 * k);
          if (TS) umma_ts(d, tm + 448 + 8 * k, bd, idesc, 1);
          else umma_ss(d, adesc + 2 * k, bd, idesc, 1);
        }
      }
    }
    const long long t1 = clock64();
    tc_commit(bar); mbar_wait(bar, ph); ph ^= 1;
    const long long t2 = clock64();
    if (pass == 1) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
}

// N = 64 SS MMAs with a tcgen05.commit (to a barrier nobody waits on during the run) after every EVERY MMAs
template <int EVERY>
__device__ __forceinline__ void run_commit_case(uint32_t tm, uint64_t adesc, uint64_t bdesc, uint64_t* bar,
                                                uint64_t* dummy, uint32_t& ph, long long* out) {
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
  for (int pass = 0; pass < 2; ++pass) {
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < REP; i += EVERY) {
#pragma unroll
      for (int k = 0; k < EVERY; ++k) umma_ss(tm + (k & 1) * 64, adesc + 2 * (k & 3), bdesc + 2 * (k & 3), idesc, 1);
      tc_commit(dummy);
    }
    const long long t1 = clock64();
    tc_commit(bar); mbar_wait(bar, ph); ph ^= 1;
    const long long t2 = clock64();
    if (pass == 1) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
}

__global__ void __launch_bounds__(128, 1) tp_kernel(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, dummy;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&dummy, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem));                 // A: 128 rows x 64 (K-major)
  const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem + 16384));         // B: up to 256 rows x 64
  uint32_t ph = 0;
  if (warp == 0 && lane == 0) {
    run_case<64, 0, 0, 1>(tm, adesc, bdesc, &bar, ph, out + 0);
    run_case<64, 0, 0, 2>(tm, adesc, bdesc, &bar, ph, out + 2);
    run_case<64, 0, 0, 4>(tm, adesc, bdesc, &bar, ph, out + 4);
    run_case<128, 0, 0, 1>(tm, adesc, bdesc, &bar, ph, out + 6);
    run_case<128, 0, 0, 2>(tm, adesc, bdesc, &bar, ph, out + 8);
    run_case<256, 0, 0, 1>(tm, adesc, bdesc, &bar, ph, out + 10);
    run_case<64, 0, 1, 1>(tm, adesc, bdesc, &bar, ph, out + 12);
    run_case<64, 0, 1, 2>(tm, adesc, bdesc, &bar, ph, out + 14);
    run_case<64, 1, 0, 1>(tm, adesc, bdesc, &bar, ph, out + 16);
    run_case<64, 1, 0, 2>(tm, adesc, bdesc, &bar, ph, out + 18);
    run_case<64, 1, 1, 1>(tm, adesc, bdesc, &bar, ph, out + 20);
    run_case<64, 1, 1, 2>(tm, adesc, bdesc, &bar, ph, out + 22);
    run_case<64, 1, 1, 4>(tm, adesc, bdesc, &bar, ph, out + 24);
    run_commit_case<1>(tm, adesc, bdesc, &bar, &dummy, ph, out + 26);
    run_commit_case<2>(tm, adesc, bdesc, &bar, &dummy, ph, out + 28);
    run_commit_case<4>(tm, adesc, bdesc, &bar, &dummy, ph, out + 30);
    run_commit_case<8>(tm, adesc, bdesc, &bar, &dummy, ph, out + 32);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

// W issuing warps (lane 0 of each) run the attention pattern concurrently: 4 N=64 MMAs into their own accumulator + one
// commit, REP/4 times.  If MMAs / commits of different threads serialise on one pipe the time grows with W.
__global__ void __launch_bounds__(256, 1) multi_kernel(long long* out, int W) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[8], dummy[8];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) { mbar_init(&bar[i], 1); mbar_init(&dummy[i], 1); } fence_mbar_init(); }
  if (warp == 7) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem));
  const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem + 16384));
  constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
  if (warp < W && lane == 0) {
    uint32_t ph = 0;
    for (int pass = 0; pass < 2; ++pass) {
      const long long t0 = clock64();
#pragma unroll 1
      for (int i = 0; i < REP; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tm + warp * 64, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
        tc_commit(&dummy[warp]);
      }
      tc_commit(&bar[warp]); mbar_wait(&bar[warp], ph); ph ^= 1;
      const long long t2 = clock64();
      if (pass == 1) out[warp] = t2 - t0;
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 7) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

struct Case { const char* name; int n; };

int main() {
  const Case h_cases[] = {
      {"SS N=64  B K-major   dep ", 64},  {"SS N=64  B K-major   alt2", 64},  {"SS N=64  B K-major   alt4", 64},
      {"SS N=128 B K-major   dep ", 128}, {"SS N=128 B K-major   alt2", 128}, {"SS N=256 B K-major   dep ", 256},
      {"SS N=64  B MN-major  dep ", 64},  {"SS N=64  B MN-major  alt2", 64},  {"TS N=64  B K-major   dep ", 64},
      {"TS N=64  B K-major   alt2", 64},  {"TS N=64  B MN-major  dep ", 64},  {"TS N=64  B MN-major  alt2", 64},
      {"TS N=64  B MN-major  alt4", 64},  {"SS N=64 alt2, commit every 1", 64}, {"SS N=64 alt2, commit every 2", 64},
      {"SS N=64 alt2, commit every 4", 64}, {"SS N=64 alt2, commit every 8", 64},
  };
  const int n = sizeof(h_cases) / sizeof(h_cases[0]);
  long long* d_out;
  cudaMalloc(&d_out, 2 * n * sizeof(long long));
  cudaFuncSetAttribute(tp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 99328 + 1024);
  tp_kernel<<<1, 128, 99328 + 1024>>>(d_out);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
  long long h_out[64];
  cudaMemcpy(h_out, d_out, 2 * n * sizeof(long long), cudaMemcpyDeviceToHost);
  printf("%d MMAs (M=128, K=16, bf16) per case, one issuing thread, one CTA\n", REP);
  printf("%-28s %14s %16s %10s\n", "case", "issue cyc/MMA", "complete cyc/MMA", "math cyc");
  for (int i = 0; i < n; ++i)
    printf("%-28s %14.1f %16.1f %10d\n", h_cases[i].name, double(h_out[2 * i]) / REP, double(h_out[2 * i + 1]) / REP,
           h_cases[i].n / 2);
  cudaFuncSetAttribute(multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 99328 + 1024);
  for (int W = 1; W <= 4; W *= 2) {
    cudaMemset(d_out, 0, 8 * sizeof(long long));
    multi_kernel<<<1, 256, 99328 + 1024>>>(d_out, W);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("multi kernel failed\n"); return 1; }
    cudaMemcpy(h_out, d_out, 8 * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < W; ++i) mx = h_out[i] > mx ? h_out[i] : mx;
    printf("%d issuing warps x (4 MMA N=64 + commit) x %d: %lld cycles = %.1f cycles per MMA per thread, %.1f aggregate\n", W,
           REP / 4, mx, double(mx) / REP, double(mx) / (REP * W));
  }
  return 0;
}
