// Shared epilogue of the tcgen05 GEMM kernels: one warp drains a 32-row x 128-column slab of an fp32 accumulator tile
// from TMEM (tcgen05.ld, thread = row), transposes it through a 4 KB XOR-swizzled smem block with 128-bit
// st.shared / ld.shared (conflict free both ways) so that every global access of a warp covers 4 rows x 128
// contiguous bytes, and applies the fused epilogue.
//
// The epilogue is compiled per MODE: the first (generic) version of this code cost ~1900 issued instructions per 32x32
// block (null-pointer branches, 64-bit address math and scalar generic LD/ST to the staging buffer repeated per row) and
// made every K = 768 GEMM epilogue-bound (~8 us per tile against a 3.8 us mainloop; profiles/README.md).
//   EPI_BF16       out_bf16 = acc (+bias)                       (qkv, Wk|Wv projections)
//   EPI_BF16_GELU  out_bf16 = gelu(acc + bias)                  (MLP linear1, packer proj_mpls.0)
//   EPI_F32_RESID  out_f32  = acc + bias + resid (may alias)    (attention out_proj, MLP linear2, output_linear)
//   EPI_GENERIC    everything GemmEpilogue can express (row remap, positional embedding, dual outputs)
// All residual loads of a 32x32 block are issued before the first store (the residual aliases the output, so the
// compiler cannot hoist them itself): this keeps the fp32 residual stream HBM-bound instead of latency-bound.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace hs {

enum : int { EPI_BF16 = 0, EPI_BF16_GELU = 1, EPI_F32_RESID = 2, EPI_GENERIC = 3 };

constexpr int EPI_WARP_BYTES = 32 * 128;   // one 32x32 fp32 block per epilogue warp

__host__ inline int epilogue_mode(const GemmEpilogue& ep) {
  const bool plain = ep.rows_per_group == 0 && ep.row_add == nullptr;
  if (plain && ep.out_bf16 != nullptr && ep.out_f32 == nullptr && ep.resid == nullptr)
    return ep.gelu ? EPI_BF16_GELU : EPI_BF16;
  if (plain && ep.out_f32 != nullptr && ep.out_bf16 == nullptr && ep.resid != nullptr && !ep.gelu &&
      ep.bias != nullptr)
    return EPI_F32_RESID;
  return EPI_GENERIC;
}

// Exact-erf GELU, 0.5 x (1 + erf(x / sqrt 2)), with erf from Abramowitz & Stegun 7.1.26 (|abs err| <= 1.5e-7):
// two MUFU ops (rcp, ex2) + ~12 FMA-pipe ops instead of erff()'s ~40-instruction branchy expansion.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erf_abs = fmaf(-p * t, e, 1.0f);          // erf(|x|/sqrt2)
  const float half_x = 0.5f * x;
  return fmaf(fabsf(half_x), erf_abs, half_x);          // 0.5x + 0.5|x| erf(|x|/sqrt2) == 0.5x(1 + erf(x/sqrt2))
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)
               : "memory");
  return v;
}

template <int MODE, typename Release>
__device__ __forceinline__ void epilogue_slab(const GemmEpilogue& ep, uint32_t taddr, uint32_t stage, int row_base,
                                              int col_base, int M, int N, int lane, Release&& release) {
  const int sub_row = lane >> 3;         // 0..3   readback: 4 rows per pass
  const int sub_chunk = lane & 7;        // 16-byte column chunk owned by this lane in the readback
  const uint32_t wr_base = stage + lane * 128;
  const int wr_sw = lane & 7;
#pragma unroll 1
  for (int chunk = 0; chunk < 4; ++chunk) {
    uint32_t v[32];
    tmem_ld32(taddr + chunk * 32, v);
    tmem_ld_wait();
    if (chunk == 3) release();           // last TMEM read of this slab: the accumulator stage can be reused
    __syncwarp();                        // previous block fully read back before it is overwritten
#pragma unroll
    for (int c = 0; c < 8; ++c)
      sts128(wr_base + ((c ^ wr_sw) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    __syncwarp();
    const int col = col_base + chunk * 32 + sub_chunk * 4;
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ep.bias != nullptr) bias4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
    const int row_first = row_base + sub_row;            // rows row_first + 4*it
    const uint32_t rd_base = stage + sub_row * 128;      // + it*512, chunk swizzled by (row & 7)

    if constexpr (MODE == EPI_BF16 || MODE == EPI_BF16_GELU) {
      __nv_bfloat16* o = ep.out_bf16 + static_cast<long>(row_first) * ep.ld_bf16 + col;
      const long step = 4L * ep.ld_bf16;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + sub_row;
        float4 x = lds128(rd_base + it * 512 + ((sub_chunk ^ (r & 7)) << 4));
        x.x += bias4.x; x.y += bias4.y; x.z += bias4.z; x.w += bias4.w;
        if constexpr (MODE == EPI_BF16_GELU) {
          x.x = gelu_fast(x.x); x.y = gelu_fast(x.y); x.z = gelu_fast(x.z); x.w = gelu_fast(x.w);
        }
        if (row_first + 4 * it < M) {
          uint2 pk;
          pk.x = pack_bf16x2(x.x, x.y);
          pk.y = pack_bf16x2(x.z, x.w);
          *reinterpret_cast<uint2*>(o + it * step) = pk;
        }
      }
    } else if constexpr (MODE == EPI_F32_RESID) {
      const float* rp = ep.resid + static_cast<long>(row_first) * ep.ld_resid + col;
      float* o = ep.out_f32 + static_cast<long>(row_first) * ep.ld_f32 + col;
      const long rstep = 4L * ep.ld_resid, ostep = 4L * ep.ld_f32;
      float4 rs[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        rs[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_first + 4 * it < M) rs[it] = *reinterpret_cast<const float4*>(rp + it * rstep);
      }
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + sub_row;
        float4 x = lds128(rd_base + it * 512 + ((sub_chunk ^ (r & 7)) << 4));
        x.x += bias4.x + rs[it].x; x.y += bias4.y + rs[it].y; x.z += bias4.z + rs[it].z; x.w += bias4.w + rs[it].w;
        if (row_first + 4 * it < M) *reinterpret_cast<float4*>(o + it * ostep) = x;
      }
    } else {
      long orow[8];
      float4 ra[8], rs[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = row_first + 4 * it;
        ra[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        rs[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        orow[it] = -1;
        if (row < M) {
          long o = row;
          if (ep.rows_per_group > 0) {
            const int g = row / ep.rows_per_group;
            const int rr = row - g * ep.rows_per_group;
            o = static_cast<long>(g) * ep.group_stride + ep.group_offset + rr;
            if (ep.row_add != nullptr)
              ra[it] = __ldg(reinterpret_cast<const float4*>(ep.row_add + static_cast<long>(rr) * N + col));
          }
          if (ep.resid != nullptr) rs[it] = *reinterpret_cast<const float4*>(ep.resid + o * ep.ld_resid + col);
          orow[it] = o;
        }
      }
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + sub_row;
        float4 x = lds128(rd_base + it * 512 + ((sub_chunk ^ (r & 7)) << 4));
        if (orow[it] >= 0) {
          x.x += bias4.x + ra[it].x + rs[it].x;
          x.y += bias4.y + ra[it].y + rs[it].y;
          x.z += bias4.z + ra[it].z + rs[it].z;
          x.w += bias4.w + ra[it].w + rs[it].w;
          if (ep.gelu) {
            x.x = gelu_fast(x.x); x.y = gelu_fast(x.y); x.z = gelu_fast(x.z); x.w = gelu_fast(x.w);
          }
          if (ep.out_f32 != nullptr) *reinterpret_cast<float4*>(ep.out_f32 + orow[it] * ep.ld_f32 + col) = x;
          if (ep.out_bf16 != nullptr) {
            uint2 pk;
            pk.x = pack_bf16x2(x.x, x.y);
            pk.y = pack_bf16x2(x.z, x.w);
            *reinterpret_cast<uint2*>(ep.out_bf16 + orow[it] * ep.ld_bf16 + col) = pk;
          }
        }
      }
    }
  }
}

}  // namespace hs
