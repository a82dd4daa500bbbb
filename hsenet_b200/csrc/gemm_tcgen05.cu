// bf16 GEMM  C[M,N] = A[M,K] * W[N,K]^T  with fused epilogues, for sm_100a.
//
// This is the kernel behind every dense projection of the HSENet visual path (SURVEY.md K1/K3/K5/K7/K8/K10/K14/K16/
// K17): MONAI SABlock.qkv / out_proj, MLPBlock.linear1/2, the perceptron patch embedding, regular_attention.W*,
// resolution_attention_v3.W*, VisualPacker_3d_phi_v3.proj_mpls.
//
// Design (B200-first, not a translation of anything in the reference, which only calls torch.nn.Linear):
//   * persistent CTAs (one per SM), static tile schedule with N fastest so concurrently running CTAs share A in L2;
//   * warp 0 = TMA producer (128B-swizzled K-major boxes), warp 1 = single-thread tcgen05.mma issuer,
//     warps 2..9 = epilogue; a 4-stage smem ring (full/empty mbarriers) feeds the tensor cores;
//   * fp32 accumulators live in TMEM, double buffered (2 x 256 columns) so the epilogue of tile i overlaps the
//     mainloop of tile i+1;
//   * epilogue: tcgen05.ld (thread = row) -> padded smem transpose -> coalesced 128-bit global accesses, fusing
//     bias, positional-embedding add, residual add (fp32 residual stream), exact-erf GELU, bf16 and/or fp32 output
//     and an output row remap (used to write patch tokens behind the cls row and packer tokens into their
//     [B,256,3072] slot, replacing the reference's torch.cat calls).
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "kernels.h"

namespace hs {

extern void count_launch();

namespace {

#ifndef HSENET_GEMM_PARK
#define HSENET_GEMM_PARK 1
#endif
__device__ __forceinline__ void gwait(uint64_t* bar, uint32_t parity) {
  if (HSENET_GEMM_PARK) mbar_wait_parked(bar, parity); else mbar_wait_nocall(bar, parity);
}

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;            // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;            // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int EPI_BYTES = EPI_WARPS * EPI_WARP_BYTES;       // 32 KB
constexpr int NUM_THREADS = 64 + EPI_WARPS * 32;            // 320
constexpr int TMEM_COLS = 512;                              // 2 accumulator stages x 256 fp32 columns
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct Barriers {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

template <int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmEpilogue ep, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  const uint32_t smem_epi = smem_u32(smem + STAGES * STAGE_BYTES);
  Barriers* bars = reinterpret_cast<Barriers*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int n_tiles = N / BN;
  const int m_tiles = (M + BM - 1) / BM;
  const int total_tiles = n_tiles * m_tiles;
  const int k_blocks = K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_base, 0);
  pdl_prologue_done();      // everything above is independent of the previous kernel's output

  // producer / issuer: convergent warps with elect_one()-guarded issue (see gemm_tcgen05_2cta.cu)
  if (warp == 0) {
    // ===================== TMA producer =====================
    int s = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * BM;
      const int n0 = (tile % n_tiles) * BN;
      for (int kb = 0; kb < k_blocks; ++kb) {
        gwait(&bars->empty[s], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->full[s], STAGE_BYTES);
          tma_load_2d(smem_a + s * A_STAGE_BYTES, &tmA, &bars->full[s], kb * BK, m0);
          tma_load_2d_hint(smem_b + s * B_STAGE_BYTES, &tmB, &bars->full[s], kb * BK, n0, kEvictLast);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
    int s = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      gwait(&bars->tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < k_blocks; ++kb) {
        gwait(&bars->full[s], phase);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem_a + s * A_STAGE_BYTES));
        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem_b + s * B_STAGE_BYTES));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in 16-byte units
            umma_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(&bars->empty[s]);
          if (kb == k_blocks - 1) tc_commit(&bars->tmem_full[acc]);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;                 // 0..7
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int col_half = ew >> 2;            // which 128-column half of the tile
    const uint32_t stage = smem_epi + ew * EPI_WARP_BYTES;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int tile_step = gridDim.x;
    auto next_m0 = [&](int t) { return (t / n_tiles) * BM; };
    float2 ln_sq[8];
    if (static_cast<int>(blockIdx.x) < total_tiles)
      epilogue_ln_load<MODE>(ep, next_m0(blockIdx.x) + quarter * 32, M, lane, ln_sq);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * BM;
      const int n0 = (tile % n_tiles) * BN;
      uint64_t ln_a[8], ln_b[8];
      epilogue_ln_coeffs<MODE>(ep, ln_sq, lane, ln_a, ln_b);                    // statistics requested one tile ago
      if (tile + tile_step < total_tiles)
        epilogue_ln_load<MODE>(ep, next_m0(tile + tile_step) + quarter * 32, M, lane, ln_sq);
      gwait(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      uint64_t* empty_bar = &bars->tmem_empty[acc];
      epilogue_slab<MODE>(ep, tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + col_half * 128, stage,
                    m0 + quarter * 32, n0 + col_half * 128, M, N, lane, ln_a, ln_b, [&]() {
                      // all TMEM reads of this accumulator stage by this warp are done: hand it back to the MMA warp
                      tc_fence_before();
                      __syncwarp();
                      if (lane == 0) mbar_arrive(empty_bar);
                    });
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
int gemm_bf16_1cta(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
                   cudaStream_t stream) {
  if (M <= 0) return HS_OK;
  if (N % BN != 0 || K % BK != 0) return HS_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15) || (lda % 8) || (ldw % 8))
    return HS_ERR_ALIGN;
  if (ep.out_f32 && ((reinterpret_cast<uintptr_t>(ep.out_f32) & 15) || (ep.ld_f32 % 4))) return HS_ERR_ALIGN;
  if (ep.out_bf16 && ((reinterpret_cast<uintptr_t>(ep.out_bf16) & 7) || (ep.ld_bf16 % 4))) return HS_ERR_ALIGN;
  if (ep.resid && ((reinterpret_cast<uintptr_t>(ep.resid) & 15) || (ep.ld_resid % 4))) return HS_ERR_ALIGN;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d_bf16(&tmA, A, K, M, lda, BK, BM);
  if (rc != HS_OK) return rc;
  rc = make_tmap_2d_bf16(&tmB, W, K, N, ldw, BK, BN);
  if (rc != HS_OK) return rc;
  static unsigned char attr_set[kMaxDevices] = {0};
  if (first_use_on_device(attr_set)) {
    bool ok = true;
    ok &= cudaFuncSetAttribute(gemm_bf16_kernel<EPI_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess;
    ok &= cudaFuncSetAttribute(gemm_bf16_kernel<EPI_BF16_GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess;
    ok &= cudaFuncSetAttribute(gemm_bf16_kernel<EPI_F32_RESID>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess;
    ok &= cudaFuncSetAttribute(gemm_bf16_kernel<EPI_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess;
    ok &= cudaFuncSetAttribute(gemm_bf16_kernel<EPI_BF16_LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess;
    ok &= cudaFuncSetAttribute(gemm_bf16_kernel<EPI_BF16_GELU_LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess;
    ok &= cudaFuncSetAttribute(gemm_bf16_kernel<EPI_F32_RESID_LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess;
    if (!ok) return HS_ERR_CUDA;
  }
  const int tiles = (N / BN) * ((M + BM - 1) / BM);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  ProfScope prof(PROF_GEMM, 2.0 * M * N * K, 2.0 * (double(M) * K + double(N) * K + double(M) * N), stream);
  if (!epilogue_stats_ok(ep, N)) return HS_ERR_ARG;
  switch (epilogue_mode(ep)) {
    case EPI_BF16: launch_pdl(gemm_bf16_kernel<EPI_BF16>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, tmA, tmB, ep, M, N, K); break;
    case EPI_BF16_GELU: launch_pdl(gemm_bf16_kernel<EPI_BF16_GELU>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, tmA, tmB, ep, M, N, K); break;
    case EPI_F32_RESID: launch_pdl(gemm_bf16_kernel<EPI_F32_RESID>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, tmA, tmB, ep, M, N, K); break;
    case EPI_BF16_LN: launch_pdl(gemm_bf16_kernel<EPI_BF16_LN>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, tmA, tmB, ep, M, N, K); break;
    case EPI_BF16_GELU_LN: launch_pdl(gemm_bf16_kernel<EPI_BF16_GELU_LN>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, tmA, tmB, ep, M, N, K); break;
    case EPI_F32_RESID_LN: launch_pdl(gemm_bf16_kernel<EPI_F32_RESID_LN>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, tmA, tmB, ep, M, N, K); break;
    default: launch_pdl(gemm_bf16_kernel<EPI_GENERIC>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, tmA, tmB, ep, M, N, K); break;
  }
  count_launch();
  return cudaGetLastError() == cudaSuccess ? HS_OK : HS_ERR_CUDA;
}

}  // namespace hs
