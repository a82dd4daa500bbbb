// bf16 GEMM  C[M,N] = A[M,K] * W[N,K]^T  on CTA pairs (tcgen05 cta_group::2), sm_100a.
//
// Why pairs: with one CTA per 128x256 tile the mainloop is bound by L2->SM operand traffic (48 KB per 4.2 MFLOP k-block,
// measured ~10 TB/s aggregate on B200 = ~45 % of the tensor peak).  A CTA pair computes a 256x256 tile with
// UMMA 256x256x16: each CTA stages only its own 128 rows of A and HALF of the B tile (32 KB per k-block), the tensor
// cores of both SMs read the peer's half of B through the pair's shared window.  Operand traffic per FLOP drops by a
// third.  The smem ring has 3 stages of K = 128 (64 KB per CTA per stage).
//
//   cluster (2,1,1); persistent: one pair per SM pair, static tile order (N fastest)
//   warp 0   TMA producer of THIS CTA's operand halves (cp.async.bulk.tensor ... .cta_group::2, signalling the
//            LEADER CTA's full barrier, which expects the bytes of both CTAs)
//   warp 1   leader CTA only: single-thread tcgen05.mma.cta_group::2 issuer; tcgen05.commit multicasts the
//            stage-free / accumulator-ready arrivals to both CTAs
//   warps 2..9  epilogue of this CTA's 128 rows (shared with the 1-CTA kernel: gemm_epilogue.cuh); the peer's epilogue
//            warps release the accumulator stage by a remote arrive on the leader's barrier
//   TMEM     2 x 256 fp32 columns per CTA (double-buffered accumulator), allocated with cta_group::2
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "kernels.h"

#include <cstdlib>

namespace hs {

extern void count_launch();
int gemm_bf16_1cta(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
                   cudaStream_t stream);

namespace {

#ifndef HSENET_GEMM_PARK
#define HSENET_GEMM_PARK 1
#endif
// all waits of the kernel: parked try_wait (suspend-time hint) by default; HSENET_GEMM_PARK=0 builds spin for A/B runs
__device__ __forceinline__ void gwait(uint64_t* bar, uint32_t parity) {
  if (HSENET_GEMM_PARK) mbar_wait_parked(bar, parity); else mbar_wait_nocall(bar, parity);
}

constexpr int PM = 256;                     // pair tile rows   (128 per CTA)
constexpr int PN = 256;                     // pair tile cols   (each CTA stages 128 of the 256 W rows)
constexpr int BK = 128;                     // K per pipeline stage = two 64-column (128-byte swizzle) TMA boxes per operand
constexpr int STAGES = 3;
constexpr int SUB_BYTES = 128 * 64 * 2;     // one [128 rows x 64 cols] swizzled sub-tile, 16 KB
constexpr int A_STAGE_BYTES = 2 * SUB_BYTES; // 32 KB
constexpr int B_STAGE_BYTES = 2 * SUB_BYTES; // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int EPI_BYTES = EPI_WARPS * EPI_WARP_BYTES;
constexpr int NUM_THREADS = 64 + EPI_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 256;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared-window address of the even (leader) CTA of a pair

struct Barriers {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load into this CTA's smem, completing bytes on the LEADER CTA's mbarrier.
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0),
      "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void umma_ss_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (count 1) on the barrier at the same smem offset in BOTH CTAs of the pair once prior MMAs retire
__device__ __forceinline__ void tc_commit_2cta(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}
// arrive on the barrier at this smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n"
      ::"r"(smem_u32(bar)), "r"(rank)
      : "memory");
}

template <int MODE>
// (10 warps on 4 schedulers = 3 warps on the fullest one: 16384 / 96 threads = 170 registers is the real ceiling, which is
// what ptxas derives from the launch bounds -- __maxnreg__(200) compiles but cannot launch.)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const GemmEpilogue ep_in, int M, int N, int K, int ksplit, int kb_per, long split_stride) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  const uint32_t smem_epi = smem_u32(smem + STAGES * STAGE_BYTES);
  Barriers* bars = reinterpret_cast<Barriers*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();       // 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int n_tiles = N / PN;
  const int m_tiles = (M + PM - 1) / PM;
  const int mn_tiles = n_tiles * m_tiles;
  const int total_tiles = mn_tiles * ksplit;      // split-K: tile = (k slice, m tile, n tile); slices write separate partials
  const int k_blocks = K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars->full[s], 1);        // leader: one arrive.expect_tx covering both CTAs' bytes
      mbar_init(&bars->empty[s], 1);       // multicast tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);                 // multicast tcgen05.commit
      mbar_init(&bars->tmem_empty[a], 2 * EPI_WARPS);    // leader only: epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta(&bars->tmem_base, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_base, 0);
  pdl_prologue_done();      // everything above is independent of the previous kernel's output

  // The producer and the MMA issuer are whole WARPS that stay convergent, with the single-thread instructions inside
  // elect_one() regions: as single lanes of a diverged warp ("if (lane == 0)", round 1) every instruction that takes
  // uniform-register operands (UTMALDG, UTCHMMA, UTCBAR, SYNCS) was wrapped by ptxas in an ELECT / R2UR.BROADCAST /
  // BRA.U.ANY loop.
  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int s = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < total_tiles; tile += num_pairs) {
      const int ks = tile / mn_tiles, mn = tile - ks * mn_tiles;
      const int m0 = (mn / n_tiles) * PM + static_cast<int>(cta_rank) * 128;
      const int n0 = (mn % n_tiles) * PN + static_cast<int>(cta_rank) * 128;
      const int kb0 = ks * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
      for (int kb = kb0; kb < kb1; ++kb) {
        gwait(&bars->empty[s], phase ^ 1);
        if (elect_one()) {
          if (cta_rank == 0) mbar_arrive_expect_tx(&bars->full[s], 2 * STAGE_BYTES);
          tma_load_2d_2cta(smem_a + s * A_STAGE_BYTES, &tmA, &bars->full[s], kb * BK, m0, kEvictNormal);
          tma_load_2d_2cta(smem_b + s * B_STAGE_BYTES, &tmB, &bars->full[s], kb * BK, n0, kEvictLast);
          tma_load_2d_2cta(smem_a + s * A_STAGE_BYTES + SUB_BYTES, &tmA, &bars->full[s], kb * BK + 64, m0,
                           kEvictNormal);
          tma_load_2d_2cta(smem_b + s * B_STAGE_BYTES + SUB_BYTES, &tmB, &bars->full[s], kb * BK + 64, n0,
                           kEvictLast);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(PM, PN, 0, 0);
      int s = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < total_tiles; tile += num_pairs) {
        gwait(&bars->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * PN;
        const int kb0 = (tile / mn_tiles) * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          gwait(&bars->full[s], phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem_a + s * A_STAGE_BYTES));
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem_b + s * B_STAGE_BYTES));
          // 8 MMAs per barrier round trip (an mbarrier wait costs ~250 cycles, as much as two of these MMAs):
          // k = 0..3 walk the first swizzled sub-tile (+32 B each), k = 4..7 the second (+16 KB = 1024 x 16 B)
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint32_t off = (k >> 2) * (SUB_BYTES >> 4) + 2 * (k & 3);
              umma_ss_2cta(tmem_d, adesc + off, bdesc + off, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
            }
            tc_commit_2cta(&bars->empty[s]);
            if (kb == kb1 - 1) tc_commit_2cta(&bars->tmem_full[acc]);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int col_half = ew >> 2;
    const uint32_t stage = smem_epi + ew * EPI_WARP_BYTES;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int tile_step = num_pairs;
    auto next_m0 = [&](int t) { return ((t % mn_tiles) / n_tiles) * PM + static_cast<int>(cta_rank) * 128; };
    const GemmEpilogue& ep = ep_in;
    float2 ln_sq[8];
    if (pair < total_tiles) epilogue_ln_load<MODE>(ep, next_m0(pair) + quarter * 32, M, lane, ln_sq);
    for (int tile = pair; tile < total_tiles; tile += num_pairs) {
      const int ks = tile / mn_tiles, mn = tile - ks * mn_tiles;
      const int m0 = (mn / n_tiles) * PM + static_cast<int>(cta_rank) * 128;
      const int n0 = (mn % n_tiles) * PN;
      GemmEpilogue ept = ep_in;                 // split-K: slice ks writes its own fp32 partial
      if (ksplit > 1) ept.out_f32 = ep_in.out_f32 + ks * split_stride;
      uint64_t ln_a[8], ln_b[8];
      epilogue_ln_coeffs<MODE>(ep, ln_sq, lane, ln_a, ln_b);                    // statistics requested one tile ago
      if (tile + tile_step < total_tiles)
        epilogue_ln_load<MODE>(ep, next_m0(tile + tile_step) + quarter * 32, M, lane, ln_sq);
      gwait(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      uint64_t* empty_bar = &bars->tmem_empty[acc];
      epilogue_slab<MODE>(ept, tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * PN + col_half * 128, stage,
                    m0 + quarter * 32, n0 + col_half * 128, M, N, lane, ln_a, ln_b, [&]() {
                      tc_fence_before();
                      __syncwarp();
                      if (lane == 0) mbar_arrive_cluster(empty_bar, 0);   // the leader's MMA thread waits on it
                    });
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();        // neither CTA may leave (or free TMEM) while its peer can still touch it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
}

}  // namespace

int gemm_bf16_2cta(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
                   cudaStream_t stream, int ksplit = 1, long split_stride = 0) {
  if (M <= 0) return HS_OK;
  if (N % PN != 0 || K % BK != 0) return HS_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15) || (lda % 8) || (ldw % 8))
    return HS_ERR_ALIGN;
  if (ep.out_f32 && ((reinterpret_cast<uintptr_t>(ep.out_f32) & 15) || (ep.ld_f32 % 4))) return HS_ERR_ALIGN;
  if (ep.out_bf16 && ((reinterpret_cast<uintptr_t>(ep.out_bf16) & 7) || (ep.ld_bf16 % 4))) return HS_ERR_ALIGN;
  if (ep.resid && ((reinterpret_cast<uintptr_t>(ep.resid) & 15) || (ep.ld_resid % 4))) return HS_ERR_ALIGN;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d_bf16(&tmA, A, K, M, lda, 64, 128);
  if (rc != HS_OK) return rc;
  rc = make_tmap_2d_bf16(&tmB, W, K, N, ldw, 64, 128);
  if (rc != HS_OK) return rc;
  static unsigned char attr_set[kMaxDevices] = {0};
  if (first_use_on_device(attr_set)) {
    bool ok = true;
    ok &= cudaFuncSetAttribute(gemm_bf16_2cta_kernel<EPI_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess;
    ok &= cudaFuncSetAttribute(gemm_bf16_2cta_kernel<EPI_BF16_GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess;
    ok &= cudaFuncSetAttribute(gemm_bf16_2cta_kernel<EPI_F32_RESID>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess;
    ok &= cudaFuncSetAttribute(gemm_bf16_2cta_kernel<EPI_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess;
    ok &= cudaFuncSetAttribute(gemm_bf16_2cta_kernel<EPI_BF16_LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess;
    ok &= cudaFuncSetAttribute(gemm_bf16_2cta_kernel<EPI_BF16_GELU_LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess;
    ok &= cudaFuncSetAttribute(gemm_bf16_2cta_kernel<EPI_F32_RESID_LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) == cudaSuccess;
    if (!ok) return HS_ERR_CUDA;
  }
  const int k_blocks = K / BK;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > k_blocks) ksplit = k_blocks;
  const int kb_per = (k_blocks + ksplit - 1) / ksplit;
  ksplit = (k_blocks + kb_per - 1) / kb_per;            // no empty slice
  const int tiles = (N / PN) * ((M + PM - 1) / PM) * ksplit;
  const int max_pairs = num_sms() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  const int grid = 2 * pairs;
  ProfScope prof(PROF_GEMM, 2.0 * M * N * K, 2.0 * (double(M) * K + double(N) * K + double(M) * N), stream);
  if (!epilogue_stats_ok(ep, N)) return HS_ERR_ARG;
  switch (epilogue_mode(ep)) {
    case EPI_BF16: launch_pdl(gemm_bf16_2cta_kernel<EPI_BF16>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, tmA, tmB, ep, M, N, K, ksplit, kb_per, split_stride); break;
    case EPI_BF16_GELU: launch_pdl(gemm_bf16_2cta_kernel<EPI_BF16_GELU>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, tmA, tmB, ep, M, N, K, ksplit, kb_per, split_stride); break;
    case EPI_F32_RESID: launch_pdl(gemm_bf16_2cta_kernel<EPI_F32_RESID>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, tmA, tmB, ep, M, N, K, ksplit, kb_per, split_stride); break;
    case EPI_BF16_LN: launch_pdl(gemm_bf16_2cta_kernel<EPI_BF16_LN>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, tmA, tmB, ep, M, N, K, ksplit, kb_per, split_stride); break;
    case EPI_BF16_GELU_LN: launch_pdl(gemm_bf16_2cta_kernel<EPI_BF16_GELU_LN>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, tmA, tmB, ep, M, N, K, ksplit, kb_per, split_stride); break;
    case EPI_F32_RESID_LN: launch_pdl(gemm_bf16_2cta_kernel<EPI_F32_RESID_LN>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, tmA, tmB, ep, M, N, K, ksplit, kb_per, split_stride); break;
    default: launch_pdl(gemm_bf16_2cta_kernel<EPI_GENERIC>, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, tmA, tmB, ep, M, N, K, ksplit, kb_per, split_stride); break;
  }
  count_launch();
  return cudaGetLastError() == cudaSuccess ? HS_OK : HS_ERR_CUDA;
}

// Split-K GEMM for the weight gradients (M, N small, K = the token dimension): `ksplit` K slices run as independent tiles and
// write fp32 partials [ksplit][M][N] (finish with a fixed-order sum).  Returns the number of slices actually used.
int gemm_bf16_splitk(const void* A, int lda, const void* W, int ldw, int M, int N, int K, float* partials, int ksplit,
                     cudaStream_t stream, int* used) {
  const int k_blocks = K / BK;
  if (ksplit > k_blocks) ksplit = k_blocks;
  if (ksplit < 1) ksplit = 1;
  const int kb_per = (k_blocks + ksplit - 1) / ksplit;
  *used = (k_blocks + kb_per - 1) / kb_per;
  GemmEpilogue ep;
  ep.out_f32 = partials; ep.ld_f32 = N;
  return gemm_bf16_2cta(A, lda, W, ldw, M, N, K, ep, stream, ksplit, static_cast<long>(M) * N);
}

// Dispatcher: CTA pairs for everything with more than one 128-row tile; HSENET_GEMM_1CTA=1 forces the 1-CTA kernel.
int gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
              cudaStream_t stream) {
  const char* e = std::getenv("HSENET_GEMM_1CTA");     // read per call so tests can flip it in-process
  const bool force_1cta = e != nullptr && e[0] == '1';
  GemmEpilogue ep2 = ep;
  const char* g = std::getenv("HSENET_GELU_ERF");      // 1: A&S exact-erf form also for bf16 outputs (A/B switch)
  if (ep2.gelu != 0 && g != nullptr && g[0] == '1') ep2.gelu = 2;
  if (force_1cta || M <= 128 || (K % 128) != 0) return gemm_bf16_1cta(A, lda, W, ldw, M, N, K, ep2, stream);
  return gemm_bf16_2cta(A, lda, W, ldw, M, N, K, ep2, stream);
}

}  // namespace hs
