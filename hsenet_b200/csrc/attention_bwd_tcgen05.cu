// Backward of the fused self attention (the op is MONAI SABlock's einsum / softmax / einsum
// as instantiated at vit.py:438-443), recompute style: nothing but the [B,12,S] log-sum-exp vector is kept from the
// forward, P is rebuilt from Q K^T on the tensor cores.
//     P = 2^(s c - lse),   dP = dO V^T,   D = rowsum(dO o O),   dS = P o (dP - D) / sqrt(64)
//     dQ = dS K,           dK = dS^T Q,   dV = P^T dO
// Two kernels, both built like the forward one (TMA -> swizzled smem tiles of the fused qkv activation, tcgen05.mma
// into TMEM, softmax-role warps turn the score tile into a bf16 operand that goes back to the tensor core as the
// TMEM-resident A operand) and both WITHOUT atomics, so gradients are bit-reproducible:
//   attn_bwd_dq_kernel   CTA = 128 queries of one (volume, head); loops over 64-key steps:
//                          S = Q K^T, dP = dO V^T (SS)  ->  dS (bf16, aliased onto S)  ->  dQ += dS K (TS, K MN-major)
//   attn_bwd_dkv_kernel  CTA = 128 keys; loops over 64-query steps with the TRANSPOSED tiles, so that keys are TMEM lanes:
//                          S^T = K Q^T, dP^T = V dO^T (SS)  ->  P^T, dS^T (bf16, aliased)  ->
//                          dV += P^T dO, dK += dS^T Q (TS, dO / Q as MN-major B operands)
// S and dP are recomputed in both kernels (7 tile GEMMs instead of the 5 of a dQ-atomics design).
// The same swizzled tile serves as a K-major operand in one MMA and as an MN-major operand in the other (only the
// descriptor differs), so every tile is loaded once per use.  Roles are convergent warps with elect_one()-guarded issue.
// Single-buffered score tiles (two CTAs per SM overlap each other); this is the correctness-first version.
#include "common.cuh"
#include "kernels.h"

#include <type_traits>

namespace hs {

extern void count_launch();

namespace {

constexpr int QT = 128;                    // rows per CTA (queries for dQ, keys for dK/dV) and per TMA tile
constexpr int KS = 64;                     // columns per step
constexpr int TILE_BYTES = 128 * 64 * 2;   // 16 KB
constexpr int SUB_BYTES = KS * 128;        // 64 rows x 128 B
constexpr int BWD_THREADS = 192;           // producer warp, issuer warp, four softmax-role warps
constexpr int BWD_TMEM_COLS = 256;
constexpr int BWD_SMEM = 6 * TILE_BYTES + 1024 + 256;
constexpr float kC = 0.125f * 1.4426950408889634f;   // head_dim^-0.5 * log2(e)

struct BwdBarriers {
  uint64_t res_full;            // the CTA-resident tiles (Q + dO, or K + V)
  uint64_t a_full[2], a_empty[2];
  uint64_t b_full[2], b_empty[2];
  uint64_t s_full, p_full, done;
  uint32_t tmem_base;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_ld32u(uint32_t taddr, uint32_t* r) {
  tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(r));
}

// ---------------------------------------------------------------------------------------------------------------------
// dQ
// ---------------------------------------------------------------------------------------------------------------------
//   TMEM columns   0..63 S (dS packed at 0..31 after the read)   64..127 dP   128..191 dQ
__global__ void __launch_bounds__(BWD_THREADS, 2)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const float* __restrict__ lse, const float* __restrict__ dvec, __nv_bfloat16* __restrict__ d_qkv,
                   int S) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sDO = smem + TILE_BYTES;
  uint8_t* sK = smem + 2 * TILE_BYTES;      // 2 stages
  uint8_t* sV = smem + 4 * TILE_BYTES;      // 2 stages
  BwdBarriers* bars = reinterpret_cast<BwdBarriers*>(smem + 6 * TILE_BYTES);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QT;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int row0 = b * S;
  const int ntiles = (S + QT - 1) / QT;
  const int nsub = (S + KS - 1) / KS;
  const int sp = ntiles * QT;
  constexpr uint32_t COL_S = 0, COL_DP = 64, COL_DQ = 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(&bars->res_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->a_full[s], 1); mbar_init(&bars->a_empty[s], 1);
      mbar_init(&bars->b_full[s], 1); mbar_init(&bars->b_empty[s], 1);
    }
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->p_full, 4);
    mbar_init(&bars->done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, BWD_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_base, 0);

  if (warp == 0) {
    // ---- TMA producer: Q and dO tiles once, then the K (a) / V (b) rings ----
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars->res_full, 2 * TILE_BYTES);
      tma_load_2d(sQ, &tmQKV, &bars->res_full, h * kHeadDim, row0 + q0);
      tma_load_2d(sDO, &tmDO, &bars->res_full, h * kHeadDim, row0 + q0);
    }
    __syncwarp();
    for (int j = 0; j < ntiles; ++j) {
      const int st = j & 1;
      const uint32_t par = ((j >> 1) & 1) ^ 1;
      mbar_wait_parked(&bars->a_empty[st], par);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->a_full[st], TILE_BYTES);
        tma_load_2d_hint(sK + st * TILE_BYTES, &tmQKV, &bars->a_full[st], kHidden + h * kHeadDim, row0 + j * QT, kEvictLast);
      }
      __syncwarp();
      mbar_wait_parked(&bars->b_empty[st], par);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->b_full[st], TILE_BYTES);
        tma_load_2d_hint(sV + st * TILE_BYTES, &tmQKV, &bars->b_full[st], 2 * kHidden + h * kHeadDim, row0 + j * QT,
                         kEvictLast);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----
    constexpr uint32_t idesc_s = make_idesc_bf16(QT, KS, 0, 0);            // [128 x 64] = A[128 x 64d] B[64 x 64d]^T
    constexpr uint32_t idesc_acc = make_idesc_bf16(QT, kHeadDim, 0, 1);    // [128 x 64d] += A[128 x 64 keys] B (MN-major)
    const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ));
    const uint64_t dodesc = make_smem_desc_sw128(smem_u32(sDO));
    mbar_wait_parked(&bars->res_full, 0);
    for (int t = 0; t < nsub; ++t) {
      const int j = t >> 1, st = j & 1, sub = t & 1;
      const bool last_of_tile = sub == 1 || t == nsub - 1;
      if (sub == 0) {
        mbar_wait_parked(&bars->a_full[st], (j >> 1) & 1);
        mbar_wait_parked(&bars->b_full[st], (j >> 1) & 1);
      }
      tc_fence_after();
      const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK + st * TILE_BYTES + sub * SUB_BYTES));
      const uint64_t vdesc = make_smem_desc_sw128(smem_u32(sV + st * TILE_BYTES + sub * SUB_BYTES));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k)
          umma_ss(tmem_base + COL_S, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);       // S = Q K^T
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k)
          umma_ss(tmem_base + COL_DP, dodesc + 2 * k, vdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);     // dP = dO V^T
        tc_commit(&bars->s_full);
        if (last_of_tile) tc_commit(&bars->b_empty[st]);
      }
      __syncwarp();
      mbar_wait_parked(&bars->p_full, t & 1);          // dS stored, S / dP in registers
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < KS / 16; ++k)               // dQ += dS K : A = 16 keys = 8 packed columns, B = 16 key rows
          umma_ts(tmem_base + COL_DQ, tmem_base + COL_S + 8 * k, kdesc + 128 * k, idesc_acc, (t | k) != 0 ? 1u : 0u);
        if (last_of_tile) tc_commit(&bars->a_empty[st]);
        if (t == nsub - 1) tc_commit(&bars->done);
      }
      __syncwarp();
    }
  } else {
    // ---- score warps: thread = query row ----
    const int quarter = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const int qi = q0 + quarter * 32 + lane;
    const long vidx = (static_cast<long>(b) * kHeads + h) * sp + qi;
    const float l2 = lse[vidx];                 // +inf for rows past the sequence end: P = 0 there
    const float dv = dvec[vidx];
    auto step = [&](const int t, auto masked) {
      mbar_wait_parked(&bars->s_full, t & 1);
      tc_fence_after();
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t s[32], dp[32], pk[16];
        tmem_ld32u(tmem_base + lane_base + COL_S + hf * 32, s);
        tmem_ld32u(tmem_base + lane_base + COL_DP + hf * 32, dp);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = ex2f(fmaf(__uint_as_float(s[i]), kC, -l2));
          float p1 = ex2f(fmaf(__uint_as_float(s[i + 1]), kC, -l2));
          if constexpr (decltype(masked)::value) {
            const int key = t * KS + hf * 32 + i;
            if (key >= S) p0 = 0.f;
            if (key + 1 >= S) p1 = 0.f;
          }
          const float d0 = p0 * (__uint_as_float(dp[i]) - dv) * 0.125f;
          const float d1 = p1 * (__uint_as_float(dp[i + 1]) - dv) * 0.125f;
          pk[i >> 1] = pack_bf16x2(d0, d1);
        }
        tmem_st16(tmem_base + lane_base + COL_S + hf * 16, pk);    // aliased: columns this thread has already read
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full);
    };
    const bool ragged = (S % KS) != 0;
    for (int t = 0; t < nsub - (ragged ? 1 : 0); ++t) step(t, std::false_type{});
    if (ragged) step(nsub - 1, std::true_type{});
    mbar_wait_parked(&bars->done, 0);
    tc_fence_after();
    uint32_t o[32];
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      tmem_ld32u(tmem_base + lane_base + COL_DQ + ch * 32, o);
      tmem_ld_wait();
      if (qi < S) {
        uint4* dst = reinterpret_cast<uint4*>(d_qkv + (static_cast<long>(row0) + qi) * (3 * kHidden) + h * kHeadDim + ch * 32);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]), __uint_as_float(o[g * 8 + 1]));
          u.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]), __uint_as_float(o[g * 8 + 3]));
          u.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]), __uint_as_float(o[g * 8 + 5]));
          u.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]), __uint_as_float(o[g * 8 + 7]));
          dst[g] = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BWD_TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// dK, dV
// ---------------------------------------------------------------------------------------------------------------------
//   TMEM columns   0..63 S^T (P^T packed at 0..31)   64..127 dP^T (dS^T packed at 64..95)   128..191 dK   192..255 dV
__global__ void __launch_bounds__(BWD_THREADS, 2)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                    const float* __restrict__ lse, const float* __restrict__ dvec, __nv_bfloat16* __restrict__ d_qkv,
                    int S) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + TILE_BYTES;
  uint8_t* sQ = smem + 2 * TILE_BYTES;      // 2 stages
  uint8_t* sDO = smem + 4 * TILE_BYTES;     // 2 stages
  BwdBarriers* bars = reinterpret_cast<BwdBarriers*>(smem + 6 * TILE_BYTES);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * QT;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int row0 = b * S;
  const int ntiles = (S + QT - 1) / QT;
  const int nsub = (S + KS - 1) / KS;
  const int sp = ntiles * QT;
  constexpr uint32_t COL_ST = 0, COL_DPT = 64, COL_DK = 128, COL_DV = 192;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(&bars->res_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->a_full[s], 1); mbar_init(&bars->a_empty[s], 1);
      mbar_init(&bars->b_full[s], 1); mbar_init(&bars->b_empty[s], 1);
    }
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->p_full, 4);
    mbar_init(&bars->done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, BWD_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_base, 0);

  if (warp == 0) {
    // ---- TMA producer: this CTA's K and V tiles once, then the ring of (Q, dO) tile pairs (a_full / a_empty) ----
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars->res_full, 2 * TILE_BYTES);
      tma_load_2d(sK, &tmQKV, &bars->res_full, kHidden + h * kHeadDim, row0 + k0);
      tma_load_2d(sV, &tmQKV, &bars->res_full, 2 * kHidden + h * kHeadDim, row0 + k0);
    }
    __syncwarp();
    for (int j = 0; j < ntiles; ++j) {
      const int st = j & 1;
      mbar_wait_parked(&bars->a_empty[st], ((j >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->a_full[st], 2 * TILE_BYTES);
        tma_load_2d_hint(sQ + st * TILE_BYTES, &tmQKV, &bars->a_full[st], h * kHeadDim, row0 + j * QT, kEvictLast);
        tma_load_2d_hint(sDO + st * TILE_BYTES, &tmDO, &bars->a_full[st], h * kHeadDim, row0 + j * QT, kEvictLast);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----
    constexpr uint32_t idesc_s = make_idesc_bf16(QT, KS, 0, 0);            // [128 keys x 64 q] = A[128 x 64d] B[64q x 64d]^T
    constexpr uint32_t idesc_acc = make_idesc_bf16(QT, kHeadDim, 0, 1);    // [128 keys x 64d] += A[128 x 64q] B (MN-major)
    const uint64_t kdesc = make_smem_desc_sw128(smem_u32(sK));
    const uint64_t vdesc = make_smem_desc_sw128(smem_u32(sV));
    mbar_wait_parked(&bars->res_full, 0);
    for (int t = 0; t < nsub; ++t) {
      const int j = t >> 1, st = j & 1, sub = t & 1;
      const bool last_of_tile = sub == 1 || t == nsub - 1;
      if (sub == 0) mbar_wait_parked(&bars->a_full[st], (j >> 1) & 1);
      tc_fence_after();
      const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ + st * TILE_BYTES + sub * SUB_BYTES));
      const uint64_t dodesc = make_smem_desc_sw128(smem_u32(sDO + st * TILE_BYTES + sub * SUB_BYTES));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k)
          umma_ss(tmem_base + COL_ST, kdesc + 2 * k, qdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);      // S^T = K Q^T
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k)
          umma_ss(tmem_base + COL_DPT, vdesc + 2 * k, dodesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);    // dP^T = V dO^T
        tc_commit(&bars->s_full);
      }
      __syncwarp();
      mbar_wait_parked(&bars->p_full, t & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < KS / 16; ++k)               // dV += P^T dO
          umma_ts(tmem_base + COL_DV, tmem_base + COL_ST + 8 * k, dodesc + 128 * k, idesc_acc, (t | k) != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < KS / 16; ++k)               // dK += dS^T Q
          umma_ts(tmem_base + COL_DK, tmem_base + COL_DPT + 8 * k, qdesc + 128 * k, idesc_acc, (t | k) != 0 ? 1u : 0u);
        if (last_of_tile) tc_commit(&bars->a_empty[st]);
        if (t == nsub - 1) tc_commit(&bars->done);
      }
      __syncwarp();
    }
  } else {
    // ---- score warps: thread = key row; the per-query lse / D values of a step are the same for every thread ----
    const int quarter = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const int ki = k0 + quarter * 32 + lane;
    const float* lrow = lse + (static_cast<long>(b) * kHeads + h) * sp;     // +inf past the sequence end: P = 0 there
    const float* drow = dvec + (static_cast<long>(b) * kHeads + h) * sp;
    for (int t = 0; t < nsub; ++t) {
      mbar_wait_parked(&bars->s_full, t & 1);
      tc_fence_after();
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t s[32], dp[32], pp[16], pd[16];
        tmem_ld32u(tmem_base + lane_base + COL_ST + hf * 32, s);
        tmem_ld32u(tmem_base + lane_base + COL_DPT + hf * 32, dp);
        const float4* l4 = reinterpret_cast<const float4*>(lrow + t * KS + hf * 32);
        const float4* d4 = reinterpret_cast<const float4*>(drow + t * KS + hf * 32);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 l = __ldg(l4 + g), d = __ldg(d4 + g);
          const float lv[4] = {l.x, l.y, l.z, l.w}, dvv[4] = {d.x, d.y, d.z, d.w};
          float p[4], ds[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            p[u] = ex2f(fmaf(__uint_as_float(s[g * 4 + u]), kC, -lv[u]));
            ds[u] = p[u] * (__uint_as_float(dp[g * 4 + u]) - dvv[u]) * 0.125f;
          }
          pp[g * 2] = pack_bf16x2(p[0], p[1]);
          pp[g * 2 + 1] = pack_bf16x2(p[2], p[3]);
          pd[g * 2] = pack_bf16x2(ds[0], ds[1]);
          pd[g * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
        }
        tmem_st16(tmem_base + lane_base + COL_ST + hf * 16, pp);
        tmem_st16(tmem_base + lane_base + COL_DPT + hf * 16, pd);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full);
    }
    mbar_wait_parked(&bars->done, 0);
    tc_fence_after();
    uint32_t o[32];
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {           // 0,1: dK columns 0..31 / 32..63;  2,3: dV
      tmem_ld32u(tmem_base + lane_base + COL_DK + ch * 32, o);
      tmem_ld_wait();
      if (ki < S) {
        const int col = (ch < 2 ? kHidden : 2 * kHidden) + h * kHeadDim + (ch & 1) * 32;
        uint4* dst = reinterpret_cast<uint4*>(d_qkv + (static_cast<long>(row0) + ki) * (3 * kHidden) + col);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]), __uint_as_float(o[g * 8 + 1]));
          u.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]), __uint_as_float(o[g * 8 + 3]));
          u.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]), __uint_as_float(o[g * 8 + 5]));
          u.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]), __uint_as_float(o[g * 8 + 7]));
          dst[g] = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BWD_TMEM_COLS);
  }
}

}  // namespace

int attention_rowdot_bf16(const __nv_bfloat16* out, const __nv_bfloat16* d_out, float* dvec, int B, int S,
                          cudaStream_t stream);

int attention_bwd_bf16(const __nv_bfloat16* qkv, const __nv_bfloat16* out, const __nv_bfloat16* d_out, const float* lse,
                       float* dvec, __nv_bfloat16* d_qkv, int B, int S, cudaStream_t stream) {
  if (B <= 0 || S <= 0) return HS_OK;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 15) ||
      (reinterpret_cast<uintptr_t>(d_qkv) & 15))
    return HS_ERR_ALIGN;
  int rc = attention_rowdot_bf16(out, d_out, dvec, B, S, stream);
  if (rc != HS_OK) return rc;
  CUtensorMap tmQKV, tmDO;
  rc = make_tmap_2d_bf16(&tmQKV, qkv, 3 * kHidden, static_cast<uint64_t>(B) * S, 3 * kHidden, kHeadDim, 128);
  if (rc != HS_OK) return rc;
  rc = make_tmap_2d_bf16(&tmDO, d_out, kHidden, static_cast<uint64_t>(B) * S, kHidden, kHeadDim, 128);
  if (rc != HS_OK) return rc;
  static unsigned char attr_set[kMaxDevices] = {0};
  if (first_use_on_device(attr_set)) {
    bool ok = true;
    ok &= cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM) == cudaSuccess;
    ok &= cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM) == cudaSuccess;
    if (!ok) return HS_ERR_CUDA;
  }
  const dim3 grid((S + QT - 1) / QT, kHeads, B);
  attn_bwd_dq_kernel<<<grid, BWD_THREADS, BWD_SMEM, stream>>>(tmQKV, tmDO, lse, dvec, d_qkv, S);
  count_launch();
  attn_bwd_dkv_kernel<<<grid, BWD_THREADS, BWD_SMEM, stream>>>(tmQKV, tmDO, lse, dvec, d_qkv, S);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? HS_OK : HS_ERR_CUDA;
}

}  // namespace hs
