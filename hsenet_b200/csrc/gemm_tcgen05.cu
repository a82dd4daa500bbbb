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
#include "kernels.h"

namespace hs {

extern void count_launch();

namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;            // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;            // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int EPI_PITCH = 33;
constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_PITCH * 4;   // 33,792 B
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

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmEpilogue ep, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  float* smem_epi = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  Barriers* bars = reinterpret_cast<Barriers*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);

  const int warp = threadIdx.x >> 5;
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
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BM;
        const int n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&bars->empty[s], phase ^ 1);
          mbar_arrive_expect_tx(&bars->full[s], STAGE_BYTES);
          tma_load_2d(smem_a + s * A_STAGE_BYTES, &tmA, &bars->full[s], kb * BK, m0);
          tma_load_2d_hint(smem_b + s * B_STAGE_BYTES, &tmB, &bars->full[s], kb * BK, n0, kEvictLast);
          if (++s == STAGES) { s = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      int s = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&bars->full[s], phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem_a + s * A_STAGE_BYTES));
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem_b + s * B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in 16-byte units
            umma_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(&bars->empty[s]);
          if (++s == STAGES) { s = 0; phase ^= 1; }
        }
        tc_commit(&bars->tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;                 // 0..7
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int col_half = ew >> 2;            // which 128-column half of the tile
    float* stage = smem_epi + ew * 32 * EPI_PITCH;
    const int sub_row = lane >> 3;           // 0..3   (readback: 4 rows per pass)
    const int sub_col = (lane & 7) * 4;      // 0..28  (4 consecutive columns per lane)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * BM;
      const int n0 = (tile % n_tiles) * BN;
      mbar_wait(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < 4; ++chunk) {
        const int c0 = col_half * 128 + chunk * 32;
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + c0, v);
        tmem_ld_wait();
        if (chunk == 3) {
          // all TMEM reads of this accumulator stage by this warp are done: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) stage[lane * EPI_PITCH + j] = __uint_as_float(v[j]);
        __syncwarp();
        const int col = n0 + c0 + sub_col;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ep.bias != nullptr) bias4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + sub_row;
          const int row = m0 + quarter * 32 + r;
          if (row < M) {
            float4 x;
            x.x = stage[r * EPI_PITCH + sub_col + 0] + bias4.x;
            x.y = stage[r * EPI_PITCH + sub_col + 1] + bias4.y;
            x.z = stage[r * EPI_PITCH + sub_col + 2] + bias4.z;
            x.w = stage[r * EPI_PITCH + sub_col + 3] + bias4.w;
            long orow = row;
            if (ep.rows_per_group > 0) {
              const int g = row / ep.rows_per_group;
              const int rr = row - g * ep.rows_per_group;
              orow = static_cast<long>(g) * ep.group_stride + ep.group_offset + rr;
              if (ep.row_add != nullptr) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(ep.row_add + static_cast<long>(rr) * N + col));
                x.x += a.x; x.y += a.y; x.z += a.z; x.w += a.w;
              }
            }
            if (ep.resid != nullptr) {
              const float4 a = *reinterpret_cast<const float4*>(ep.resid + orow * ep.ld_resid + col);
              x.x += a.x; x.y += a.y; x.z += a.z; x.w += a.w;
            }
            if (ep.gelu) {
              x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w);
            }
            if (ep.out_f32 != nullptr) *reinterpret_cast<float4*>(ep.out_f32 + orow * ep.ld_f32 + col) = x;
            if (ep.out_bf16 != nullptr) {
              uint2 pk;
              pk.x = pack_bf16x2(x.x, x.y);
              pk.y = pack_bf16x2(x.z, x.w);
              *reinterpret_cast<uint2*>(ep.out_bf16 + orow * ep.ld_bf16 + col) = pk;
            }
          }
        }
        __syncwarp();
      }
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
int gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
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
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) !=
        cudaSuccess)
      return HS_ERR_CUDA;
    attr_set = true;
  }
  const int tiles = (N / BN) * ((M + BM - 1) / BM);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  ProfScope prof(PROF_GEMM, 2.0 * M * N * K, 2.0 * (double(M) * K + double(N) * K + double(M) * N), stream);
  gemm_bf16_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmA, tmB, ep, M, N, K);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? HS_OK : HS_ERR_CUDA;
}

}  // namespace hs
