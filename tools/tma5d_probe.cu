// Probe: how does a 5-D TMA box with a 64-byte inner extent land in shared memory under SWIZZLE_128B?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma5d_probe tools/tma5d_probe.cu ; run on the GPU box
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int c1, int c3, int c4) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  uint32_t sdst = (uint32_t)__cvta_generic_to_shared(smem), sbar = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sbar));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sbar), "r"(16384));
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(sdst), "l"((uint64_t)&tm), "r"(sbar), "r"(0), "r"(c1), "r"(0), "r"(c3), "r"(c4) : "memory");
  }
  __syncthreads();
  uint32_t ok = 0;
  while (!ok) asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0; selp.b32 %0,1,0,P;}" : "=r"(ok) : "r"(sbar));
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}
int main() {
  const int B = 1;
  size_t n = (size_t)B * 32 * 256 * 256;
  std::vector<float> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (float)i;      // value = flat voxel index (exact below 2^24)
  float *d, *o;
  cudaMalloc(&d, n * 4); cudaMalloc(&o, 4096 * 4);
  cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn fn = (EncodeTiledFn)p;
  for (int swz = 0; swz < 2; ++swz) {
    CUtensorMap tm;
    const cuuint64_t dims[5] = {16, 16, 16, 16, (cuuint64_t)B * 32};
    const cuuint64_t strides[4] = {256 * 4, 16 * 4, 4096 * 4, 65536 * 4};
    const cuuint32_t box[5] = {16, 2, 16, 8, 1};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("swizzle=%d encode rc=%d\n", swz, (int)r);
    const int c1 = 4, c3 = 8, c4 = 5;    // p2 = 4..5, hy = 8..15, z = 5
    probe<<<1, 128, 16384 + 2048>>>(tm, o, c1, c3, c4);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run: %s\n", cudaGetErrorString(e));
    std::vector<float> s(4096);
    cudaMemcpy(s.data(), o, 4096 * 4, cudaMemcpyDeviceToHost);
    // expected (my assumption): linear float index = p3 + 16*p2l + 32*wx + 512*hyl ; swizzled: 16-byte chunk ^= (row % 8), row = idx/32
    int bad_lin = 0, bad_swz = 0;
    for (int hyl = 0; hyl < 8; ++hyl) for (int wx = 0; wx < 16; ++wx) for (int p2l = 0; p2l < 2; ++p2l) for (int p3 = 0; p3 < 16; ++p3) {
      float want = (float)((size_t)c4 * 65536 + (size_t)((c3 + hyl) * 16 + (c1 + p2l)) * 256 + wx * 16 + p3);
      int lin = p3 + 16 * p2l + 32 * wx + 512 * hyl;
      int row = lin / 32, chunk = (lin % 32) / 4, within = lin % 4;
      int sw = row * 32 + ((chunk ^ (row % 8)) * 4) + within;
      if (s[lin] != want) ++bad_lin;
      if (s[sw] != want) ++bad_swz;
    }
    printf("  mismatches vs linear layout: %d, vs 128B-swizzled layout: %d (of 4096)\n", bad_lin, bad_swz);
    printf("  first 40 floats (as voxel index - base):");
    size_t base = (size_t)c4 * 65536 + (size_t)(c3 * 16 + c1) * 256;
    for (int i = 0; i < 40; ++i) printf(" %ld", (long)s[i] - (long)base);
    printf("\n");
    if (swz) {   // where does each element land?  print the smem float index of elements (hyl=0,wx=0..1,p2l,p3=0,4,8,12)
      for (int wx = 0; wx < 3; ++wx) for (int p2l = 0; p2l < 2; ++p2l) for (int p3 = 0; p3 < 16; p3 += 4) {
        float want = (float)((size_t)c4 * 65536 + (size_t)((c3 + 0) * 16 + (c1 + p2l)) * 256 + wx * 16 + p3);
        int at = -1;
        for (int i = 0; i < 4096; ++i) if (s[i] == want) { at = i; break; }
        printf("  (wx=%d,p2l=%d,p3=%d) -> smem float %d (row %d, chunk %d)\n", wx, p2l, p3, at, at / 32, (at % 32) / 4);
      }
    }
  }
  return 0;
}
