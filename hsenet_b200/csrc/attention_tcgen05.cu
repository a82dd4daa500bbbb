// Fused flash-style self attention for the 3D token sequence (MONAI SABlock core as used by vit.py:438-443):
//     out[b, :, h*64:(h+1)*64] = softmax(Q_h K_h^T / sqrt(64)) V_h,   S = 2049 tokens, 12 heads x 64.
// The reference materialises the [B,12,2049,2049] score tensor (201 MB/layer/volume in fp32); here scores never
// leave the SM: S = Q K^T accumulates in TMEM, the softmax warps turn it into bf16 P (also in TMEM), and
// O += P V runs with P as the TMEM-resident A operand.
//
// CTA = one 128-query tile of one (batch, head); 2 CTAs are co-resident per SM (256 TMEM columns each).
// Keys are processed in 64-key sub-tiles with DOUBLE-BUFFERED score / probability tiles in TMEM:
//     columns   0.. 63  S0      64..127  S1     128..159  P0     160..191  P1     192..255  O
// so Q K^T of sub-tile t+1 (and t+2) is already in flight while the softmax warps work on sub-tile t, and the
// softmax is single pass (64 fp32 scores per thread live in registers).
//   warp 0      TMA producer: Q once, then K / V 128-key tiles straight out of the fused qkv activation
//               [B*S, 2304] (column offsets 0 / 768 / 1536 + h*64) -- no head-split copies are ever made;
//   warp 1      single-thread tcgen05.mma issuer: S = Q K^T (SS, both K-major, N = 64),
//               O += P V (TS: P from TMEM, V is the MN-major B operand straight from its row-major tile);
//   warps 2..5  online softmax in fp32 (exp2 domain, lazy rescale), O correction, final normalise + bf16 store.
// Keys past the end of the sequence (2049 = 32*64 + 1) are masked in the last sub-tile; softmax warps whose 32 query
// rows are all past the end (last query tile) skip the exponentials.
#include "common.cuh"
#include "kernels.h"

namespace hs {

extern void count_launch();

namespace {

constexpr int QT = 128;                    // queries per CTA
constexpr int KT = 128;                    // keys per TMA tile
constexpr int KS = 64;                     // keys per softmax / MMA sub-tile
constexpr int K_STAGES = 3;
constexpr int V_STAGES = 2;
constexpr int TILE_BYTES = 128 * 64 * 2;   // 16 KB: 128 rows x 64 bf16, 128B-swizzled
constexpr int SUB_BYTES = KS * 128;        // 64 rows x 128 B
constexpr int ATT_THREADS = 192;
constexpr int TMEM_COLS = 256;
constexpr uint32_t COL_S = 0, COL_P = 128, COL_O = 192;
constexpr int ATT_SMEM = (1 + K_STAGES + V_STAGES) * TILE_BYTES + 1024 + 256;

struct AttBarriers {
  uint64_t q_full;
  uint64_t k_full[K_STAGES], k_empty[K_STAGES];
  uint64_t v_full[V_STAGES], v_empty[V_STAGES];
  uint64_t s_full[2], p_full[2], pv_done[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out, int S) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;
  uint8_t* sV = sK + K_STAGES * TILE_BYTES;
  AttBarriers* bars = reinterpret_cast<AttBarriers*>(sV + V_STAGES * TILE_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QT;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int row0 = b * S;                      // first row of this volume in the [B*S, 2304] activation
  const int ntiles = (S + KT - 1) / KT;        // 128-key TMA tiles
  const int nsub = (S + KS - 1) / KS;          // 64-key sub-tiles

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(&bars->q_full, 1);
    for (int s = 0; s < K_STAGES; ++s) { mbar_init(&bars->k_full[s], 1); mbar_init(&bars->k_empty[s], 1); }
    for (int s = 0; s < V_STAGES; ++s) { mbar_init(&bars->v_full[s], 1); mbar_init(&bars->v_empty[s], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->p_full[i], 4);
      mbar_init(&bars->pv_done[i], 1);
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
      mbar_arrive_expect_tx(&bars->q_full, TILE_BYTES);
      tma_load_2d(sQ, &tmQKV, &bars->q_full, h * kHeadDim, row0 + q0);
      for (int j = 0; j < ntiles; ++j) {
        const int ks = j % K_STAGES, vs = j % V_STAGES;
        mbar_wait(&bars->k_empty[ks], ((j / K_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->k_full[ks], TILE_BYTES);
        tma_load_2d_hint(sK + ks * TILE_BYTES, &tmQKV, &bars->k_full[ks], kHidden + h * kHeadDim, row0 + j * KT,
                         kEvictLast);
        mbar_wait(&bars->v_empty[vs], ((j / V_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->v_full[vs], TILE_BYTES);
        tma_load_2d_hint(sV + vs * TILE_BYTES, &tmQKV, &bars->v_full[vs], 2 * kHidden + h * kHeadDim,
                         row0 + j * KT, kEvictLast);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(QT, KS, 0, 0);          // S[128q x 64k] = Q K^T
      constexpr uint32_t idesc_pv = make_idesc_bf16(QT, kHeadDim, 0, 1);    // O[128q x 64d] += P V (V MN-major)
      const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ));
      auto issue_qk = [&](int t) {
        const int j = t >> 1, ks = j % K_STAGES;
        if ((t & 1) == 0) {                         // first sub-tile of a new K tile: wait for its TMA
          mbar_wait(&bars->k_full[ks], (j / K_STAGES) & 1);
          tc_fence_after();
        }
        const uint64_t kdesc =
            make_smem_desc_sw128(smem_u32(sK + ks * TILE_BYTES + (t & 1) * SUB_BYTES));
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k)
          umma_ss(tmem_base + COL_S + (t & 1) * KS, qdesc + 2 * k, kdesc + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        if ((t & 1) == 1 || t == nsub - 1) tc_commit(&bars->k_empty[ks]);   // K tile fully consumed
        tc_commit(&bars->s_full[t & 1]);
      };
      mbar_wait(&bars->q_full, 0);
      tc_fence_after();
      issue_qk(0);
      if (nsub > 1) issue_qk(1);
      for (int t = 0; t < nsub; ++t) {
        const int bsel = t & 1;
        mbar_wait(&bars->p_full[bsel], (t >> 1) & 1);   // softmax t done: S[bsel] consumed, P[bsel] written
        tc_fence_after();
        if (t + 2 < nsub) issue_qk(t + 2);              // refill the score buffer first
        const int j = t >> 1, vs = j % V_STAGES;
        if (bsel == 0) {
          mbar_wait(&bars->v_full[vs], (j / V_STAGES) & 1);
          tc_fence_after();
        }
        const uint64_t vdesc = make_smem_desc_sw128(smem_u32(sV + vs * TILE_BYTES + bsel * SUB_BYTES));
#pragma unroll
        for (int k = 0; k < KS / 16; ++k) {
          // A: 16 keys = 8 TMEM columns of packed bf16 pairs; B: 16 key rows = 2048 bytes = +128 (16-byte units)
          umma_ts(tmem_base + COL_O, tmem_base + COL_P + bsel * 32 + 8 * k, vdesc + 128 * k, idesc_pv,
                  (t | k) != 0 ? 1u : 0u);
        }
        if (bsel == 1 || t == nsub - 1) tc_commit(&bars->v_empty[vs]);      // V tile fully consumed
        tc_commit(&bars->pv_done[bsel]);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue warps =====================
    const int quarter = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const int qi = q0 + quarter * 32 + lane;             // query index inside the volume
    const bool warp_live = (q0 + quarter * 32) < S;      // warp-uniform: any valid query row in this warp?
    const float c = 0.125f * 1.4426950408889634f;       // head_dim^-0.5 * log2(e)
    float m = -INFINITY;                                 // running reference max (log2 domain, already scaled)
    float l = 0.f;
    for (int t = 0; t < nsub; ++t) {
      const int bsel = t & 1;
      mbar_wait(&bars->s_full[bsel], (t >> 1) & 1);
      tc_fence_after();
      uint32_t p[32];
      float alpha = 1.f;
      if (warp_live) {
        uint32_t x[64];
        tmem_ld32(tmem_base + lane_base + COL_S + bsel * KS, *reinterpret_cast<uint32_t(*)[32]>(&x[0]));
        tmem_ld32(tmem_base + lane_base + COL_S + bsel * KS + 32, *reinterpret_cast<uint32_t(*)[32]>(&x[32]));
        tmem_ld_wait();
        const int kbase = t * KS;
        if (kbase + KS > S) {                            // last sub-tile: mask keys past the sequence end
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (kbase + i >= S) x[i] = 0xff800000u;      // -inf
        }
        float tmax = __uint_as_float(x[0]);
#pragma unroll
        for (int i = 1; i < 64; ++i) tmax = fmaxf(tmax, __uint_as_float(x[i]));
        // running max with lazy rescale: only move the reference max when it grows by more than 2^8
        const float tm = tmax * c;
        if (tm > m + 8.0f) {
          alpha = ex2(m - tm);                           // m = -inf on the first sub-tile -> 0
          m = tm;
        }
        float rsum = 0.f;
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
          const float e0 = ex2(fmaf(__uint_as_float(x[i]), c, -m));
          const float e1 = ex2(fmaf(__uint_as_float(x[i + 1]), c, -m));
          rsum += e0 + e1;
          p[i >> 1] = pack_bf16x2(e0, e1);
        }
        l = l * alpha + rsum;
      }
      // O correction (rare): needs P V (t-1) retired
      if (t > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        mbar_wait(&bars->pv_done[(t - 1) & 1], ((t - 1) >> 1) & 1);
        tc_fence_after();
        uint32_t o[32];
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          tmem_ld32(tmem_base + lane_base + COL_O + ch * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(tmem_base + lane_base + COL_O + ch * 32, o);
        }
      }
      // P[bsel] was last read by P V (t-2): it must have retired before we overwrite it
      if (t >= 2) {
        mbar_wait(&bars->pv_done[bsel], ((t >> 1) - 1) & 1);
        tc_fence_after();
      }
      tmem_st32(tmem_base + lane_base + COL_P + bsel * 32, p);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[bsel]);
    }
    // ---- epilogue: O / l -> bf16 -> out[b*S + qi, h*64 .. h*64+63] -----------------------------------------------
    if (nsub >= 2) mbar_wait(&bars->pv_done[(nsub - 2) & 1], ((nsub - 2) >> 1) & 1);
    mbar_wait(&bars->pv_done[(nsub - 1) & 1], ((nsub - 1) >> 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l;
    uint32_t x[32];
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      tmem_ld32(tmem_base + lane_base + COL_O + ch * 32, x);
      tmem_ld_wait();
      if (qi < S) {
        uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<long>(row0) + qi) * kHidden + h * kHeadDim + ch * 32);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(x[g * 8 + 0]) * inv, __uint_as_float(x[g * 8 + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(x[g * 8 + 2]) * inv, __uint_as_float(x[g * 8 + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(x[g * 8 + 4]) * inv, __uint_as_float(x[g * 8 + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(x[g * 8 + 6]) * inv, __uint_as_float(x[g * 8 + 7]) * inv);
          dst[g] = u;
        }
      }
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

int attention_bf16(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int S, cudaStream_t stream) {
  if (B <= 0 || S <= 0) return HS_OK;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return HS_ERR_ALIGN;
  CUtensorMap tm;
  const int rc = make_tmap_2d_bf16(&tm, qkv, 3 * kHidden, static_cast<uint64_t>(B) * S, 3 * kHidden, kHeadDim, 128);
  if (rc != HS_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM) != cudaSuccess)
      return HS_ERR_CUDA;
    attr_set = true;
  }
  ProfScope prof(PROF_ATTENTION, 4.0 * B * kHeads * double(S) * S * kHeadDim, 2.0 * B * double(S) * 4 * kHidden,
                 stream);
  attention_kernel<<<dim3((S + QT - 1) / QT, kHeads, B), ATT_THREADS, ATT_SMEM, stream>>>(tm, out, S);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? HS_OK : HS_ERR_CUDA;
}

}  // namespace hs
