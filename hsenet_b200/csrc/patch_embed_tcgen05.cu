// Patch embedding of BOTH towers as ONE implicit-im2col GEMM (BASELINE north_star (1); the op is MONAI's perceptron
// PatchEmbeddingBlock, vit.py:428-437 / 290-299, called at vit.py:455 / 325):
//     X[b, t, :] = W patch(b, t) + bias + pos[t],   patch(b,t)[p1*256 + p2*16 + p3] = vol[b, dz*4+p1, hy*16+p2, wx*16+p3]
// with t = (dz*16 + hy)*16 + wx -- exactly the gather map hsenet_patch_gather_map dumps.  No patch matrix is ever written:
//   * the fp32 volume [B,32,256,256] is a 5-D tensor (p3:16, p2:16, wx:16, hy:16, z:B*32) with byte strides
//     (4, 1024, 64, 16384, 262144); one TMA box (16, 1, 16, 8, 1) with SWIZZLE_64B lands in shared memory as a K-major
//     [128 tokens x 16 features] fp32 operand tile (64-byte rows) of 128 consecutive tokens (8 hy x 16 wx) and the 16
//     features of one (p1, p2).  A pipeline stage holds two such tiles (p2 = 2j, 2j+1) against one 32-feature weight tile.
//     (A single box with two p2 values and SWIZZLE_128B would give 128-byte rows, but the TMA unit faults on a swizzle
//     span wider than the box's inner extent -- tools/tma5d_probe.cu.)
//   * the MMAs are tcgen05.mma kind::tf32 (M = 128, N = 256, K = 8) straight from the fp32 tiles -- no conversion pass,
//     10-bit mantissa instead of bf16's 8;
//   * the weight operand is the two towers' patch matrices stacked to [1536,1024] fp32, so the volume is read ONCE for
//     both encoders (the reference reads it once per tower and materialises an 8.4 MB permuted copy each time);
//   * the epilogue (gemm_epilogue.cuh, generic mode) adds bias and the positional embedding and writes tower 1's rows
//     behind the cls row of its fp32 residual stream and tower 2's into its fp32 + activation-dtype patch buffers.
// Same warp roles as gemm_tcgen05.cu: warp 0 TMA producer, warp 1 MMA issuer (both convergent, elect_one-guarded issue),
// warps 2..9 epilogue; 3-stage ring of 48 KB (A 16 KB + W 32 KB); two 256-column TMEM accumulators.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "kernels.h"

namespace hs {

extern void count_launch();

namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BKF = 32;                                // fp32 elements per k-block = one 128-byte swizzle row
constexpr int STAGES = 3;
constexpr int A_STAGE_BYTES = BM * BKF * 4;            // 16 KB
constexpr int B_STAGE_BYTES = BN * BKF * 4;            // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int EPI_BYTES = EPI_WARPS * EPI_WARP_BYTES;
constexpr int NUM_THREADS = 64 + EPI_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 256;
constexpr int K_BLOCKS = kPatchDim / BKF;              // 32

struct Barriers {
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// kind::tf32: a_format = b_format = 2 (TF32), fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void umma_ss_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared-memory matrix descriptor for a K-major tile of 64-byte rows under SWIZZLE_64B: 8-row groups 512 B apart
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;          // SBO
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;                 // SWIZZLE_64B
  return d;
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}

// ep0: tower 1 (weight rows 0..767), ep1: tower 2 (rows 768..1535); n_tiles = 3 per tower present
__global__ void __launch_bounds__(NUM_THREADS, 1)
patch_embed_tf32_kernel(const __grid_constant__ CUtensorMap tmVol, const __grid_constant__ CUtensorMap tmW,
                        const GemmEpilogue ep0, const GemmEpilogue ep1, int B, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  const uint32_t smem_epi = smem_u32(smem + STAGES * STAGE_BYTES);
  Barriers* bars = reinterpret_cast<Barriers*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int M = B * kNPatch;
  const int m_tiles = M / BM;                 // 16 per volume
  const int total_tiles = n_tiles * m_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmVol);
    tma_prefetch_desc(&tmW);
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

  if (warp == 0) {
    // ===================== TMA producer: implicit im2col =====================
    int s = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile / n_tiles, n0 = (tile % n_tiles) * BN;
      const int b = mt >> 4, ti = mt & 15;             // 16 token tiles per volume: dz = ti / 2, hy0 = (ti % 2) * 8
      const int dz = ti >> 1, hy0 = (ti & 1) * 8;
      for (int kb = 0; kb < K_BLOCKS; ++kb) {
        mbar_wait_parked(&bars->empty[s], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars->full[s], STAGE_BYTES);
          // features 32*kb .. +31: p1 = kb / 8, p2 = 2 * (kb % 8) .. +1, p3 = 0 .. 15
          tma_load_5d(smem_a + s * A_STAGE_BYTES, &tmVol, &bars->full[s], 0, 2 * (kb & 7), 0, hy0,
                      b * 32 + dz * 4 + (kb >> 3));
          tma_load_5d(smem_a + s * A_STAGE_BYTES + A_STAGE_BYTES / 2, &tmVol, &bars->full[s], 0, 2 * (kb & 7) + 1, 0, hy0,
                      b * 32 + dz * 4 + (kb >> 3));
          tma_load_2d_hint(smem_b + s * B_STAGE_BYTES, &tmW, &bars->full[s], kb * BKF, n0, kEvictLast);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
    int s = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait_parked(&bars->tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < K_BLOCKS; ++kb) {
        mbar_wait_parked(&bars->full[s], phase);
        tc_fence_after();
        const uint64_t adesc0 = make_smem_desc_sw64(smem_u32(smem_a + s * A_STAGE_BYTES));
        const uint64_t adesc1 = make_smem_desc_sw64(smem_u32(smem_a + s * A_STAGE_BYTES + A_STAGE_BYTES / 2));
        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem_b + s * B_STAGE_BYTES));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BKF / 8; ++k) {    // 8 tf32 = 32 bytes along K: +2 (16-byte units) inside a swizzle row;
            // features 0..15 of the k-block come from the first 64-byte-row A tile, 16..31 from the second
            const uint64_t adesc = (k < 2 ? adesc0 : adesc1) + 2 * (k & 1);
            umma_ss_tf32(tmem_d, adesc, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(&bars->empty[s]);
          if (kb == K_BLOCKS - 1) tc_commit(&bars->tmem_full[acc]);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int col_half = ew >> 2;
    const uint32_t stage = smem_epi + ew * EPI_WARP_BYTES;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint64_t ln_a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ln_b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * BM;
      const int n0 = (tile % n_tiles) * BN;
      const bool second = n0 >= kHidden;
      const GemmEpilogue& ep = second ? ep1 : ep0;
      const int ncol = (second ? n0 - kHidden : n0) + col_half * 128;
      mbar_wait_parked(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      uint64_t* empty_bar = &bars->tmem_empty[acc];
      epilogue_slab<EPI_GENERIC>(ep, tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + col_half * 128,
                                 stage, m0 + quarter * 32, ncol, M, kHidden, lane, ln_a, ln_b, [&]() {
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn_local() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

}  // namespace

// images fp32 [B,1,32,256,256]; w_stack fp32 [n_towers*768, 1024] (tower 1 rows first); ep0 / ep1: the two towers'
// epilogues (bias, row_add = positional embedding, row remap, outputs), ep1 ignored when n_towers == 1.
int patch_embed_tf32(const float* images, const float* w_stack, int B, int n_towers, const GemmEpilogue& ep0,
                     const GemmEpilogue& ep1, cudaStream_t stream) {
  if (B <= 0) return HS_OK;
  if (n_towers != 1 && n_towers != 2) return HS_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(images) & 15) || (reinterpret_cast<uintptr_t>(w_stack) & 15)) return HS_ERR_ALIGN;
  EncodeTiledFn fn = encode_fn_local();
  if (fn == nullptr) return HS_ERR_DRIVER;
  CUtensorMap tmVol, tmW;
  {
    const cuuint64_t dims[5] = {16, 16, 16, 16, static_cast<cuuint64_t>(B) * 32};          // p3, p2, wx, hy, z
    const cuuint64_t strides[4] = {256 * 4, 16 * 4, 4096 * 4, 65536 * 4};                  // bytes, dims 1..4
    const cuuint32_t box[5] = {16, 1, 16, 8, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (fn(&tmVol, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(images), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return HS_ERR_DRIVER;
  }
  {
    const cuuint64_t dims[2] = {kPatchDim, static_cast<cuuint64_t>(n_towers) * kHidden};
    const cuuint64_t strides[1] = {kPatchDim * 4};
    const cuuint32_t box[2] = {BKF, BN};
    const cuuint32_t estr[2] = {1, 1};
    if (fn(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(w_stack), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return HS_ERR_DRIVER;
  }
  static unsigned char attr_set[kMaxDevices] = {0};
  if (first_use_on_device(attr_set) &&
      cudaFuncSetAttribute(patch_embed_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
    return HS_ERR_CUDA;
  const int n_tiles = n_towers * (kHidden / BN);
  const int tiles = n_tiles * (B * kNPatch / BM);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  ProfScope prof(PROF_GEMM, 2.0 * B * kNPatch * double(n_towers) * kHidden * kPatchDim,
                 4.0 * (double(B) * kNPatch * kPatchDim + double(n_towers) * kHidden * kPatchDim), stream);
  patch_embed_tf32_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmVol, tmW, ep0, ep1, B, n_tiles);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? HS_OK : HS_ERR_CUDA;
}

}  // namespace hs
