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
// LayerNorm between two GEMMs is folded into their epilogues instead of running as a kernel of its own (which re-read
// the fp32 residual stream, 50 MB per call at batch 8):  LN(x) W^T = rstd (x W'^T) - rstd mu colsum(W') + (W beta + b)
// with W' = gamma (.) W, so
//   EPI_F32_RESID_LN   the producer also writes a bf16 copy of its output rows (the next GEMM's A operand) and stores each
//                      row's (sum, sum of squares) over its 128-column slab -- taken from the fp32 values -- into the
//                      slab's own slot stats_out[(col / 128) * ld_stats + row]: plain stores, no atomics, so the result is
//                      bit-identical from run to run;
//   EPI_BF16_LN, EPI_BF16_GELU_LN   the consumer multiplies by W' , adds the stats_slots (<= 8) partials of every row in a
//                      fixed order and applies the per-row scale / shift.
// All residual loads of a 32x32 block are issued before the first store (the residual aliases the output, so the
// compiler cannot hoist them itself): this keeps the fp32 residual stream HBM-bound instead of latency-bound.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace hs {

enum : int { EPI_BF16 = 0, EPI_BF16_GELU = 1, EPI_F32_RESID = 2, EPI_GENERIC = 3, EPI_BF16_LN = 4, EPI_BF16_GELU_LN = 5,
              EPI_F32_RESID_LN = 6 };

#ifndef HSENET_EPI_PREFETCH
#define HSENET_EPI_PREFETCH 1
#endif
constexpr int EPI_WARP_BYTES = 32 * 128;   // one 32x32 fp32 block per epilogue warp

__host__ inline int epilogue_mode(const GemmEpilogue& ep) {
  const bool plain = ep.rows_per_group == 0 && ep.row_add == nullptr && ep.dgelu_src == nullptr;
  if (plain && ep.out_bf16 != nullptr && ep.out_f32 == nullptr && ep.resid == nullptr && ep.stats_out == nullptr) {
    if (ep.stats_in != nullptr) return ep.gelu ? EPI_BF16_GELU_LN : EPI_BF16_LN;
    return ep.gelu ? EPI_BF16_GELU : EPI_BF16;
  }
  if (plain && ep.out_f32 != nullptr && ep.resid != nullptr && !ep.gelu && ep.bias != nullptr &&
      ep.stats_in == nullptr) {
    if (ep.out_bf16 != nullptr && ep.stats_out != nullptr) return EPI_F32_RESID_LN;
    if (ep.out_bf16 == nullptr && ep.stats_out == nullptr) return EPI_F32_RESID;
  }
  return EPI_GENERIC;     // (does not implement the LayerNorm fields: callers must hit one of the cases above)
}

// LayerNorm-statistics fields a launcher must reject before it picks a kernel
__host__ inline bool epilogue_stats_ok(const GemmEpilogue& ep, int N) {
  if (ep.stats_out != nullptr && (N % 128 != 0 || N > 1024 || ep.ld_stats <= 0)) return false;
  if (ep.stats_in != nullptr && (ep.stats_slots < 1 || ep.stats_slots > 8 || ep.ld_stats <= 0 || ep.colsum == nullptr)) return false;
  return true;
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

// GELU for bf16 outputs, two values at a time on packed fp32x2 ops: 0.5 x (1 + tanh(x (p0 + p1 x^2 + p2 x^4))) with the
// coefficients fitted (minimax over [-8, 8]) to the EXACT erf GELU, max |error| 2.5e-5 -- a sixth of a bf16 ulp at
// |y| = 0.04 and below half an ulp everywhere the output is not already < 1e-2; tanh.approx adds 2^-11 relative.
// Simulated on N(0, 1.5^2) pre-activations the RMS error after bf16 rounding is 1.705e-3 against 1.693e-3 for the
// exact function.  One MUFU + three FMA-pipe issue slots per element instead of two MUFU + ~14: the linear1 epilogue
// (32K elements per CTA tile) was MUFU/issue-bound against its 6144-cycle mainloop.  x^2 is clamped at 64 so that the
// quartic cannot change sign for huge inputs (tanh is saturated there anyway).  fp32 outputs keep gelu_fast.
__device__ __forceinline__ void gelu_tanh2(float& a, float& b) {
  const uint64_t x = pack2(a, b);
  float sa, sb;
  unpack2(mul2(x, x), sa, sb);
  const uint64_t x2 = pack2(fminf(sa, 64.0f), fminf(sb, 64.0f));
  uint64_t q = fma2(x2, pack2(-3.51516790e-4f, -3.51516790e-4f), pack2(3.70056460e-2f, 3.70056460e-2f));
  q = fma2(q, x2, pack2(7.97507884e-1f, 7.97507884e-1f));
  float ua, ub;
  unpack2(mul2(q, x), ua, ub);
  float ta, tb;
  asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(ua));
  asm("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(ub));
  const uint64_t hx = mul2(x, pack2(0.5f, 0.5f));
  unpack2(fma2(hx, pack2(ta, tb), hx), a, b);
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

// LayerNorm-consuming modes: rstd and -mean*rstd of the 8 rows this lane handles in the readback (row_base + (lane>>3)
// + 4*it).  The epilogue runs BEHIND the tensor pipe in the K = 768 GEMMs, so the statistics are loaded one tile ahead
// (epilogue_ln_load for tile i+1 is issued before the slab of tile i is processed) and only turned into coefficients
// (epilogue_ln_coeffs) when their tile starts: an L2 round trip per tile on the critical path cost 11 % of the kernel.
template <int MODE>
__device__ __forceinline__ void epilogue_ln_load(const GemmEpilogue& ep, int row_base, int M, int lane, float2 (&sq)[8]) {
  if constexpr (MODE == EPI_BF16_LN || MODE == EPI_BF16_GELU_LN) {
#pragma unroll
    // lane l fetches the (<= 8) column-slab partials of ITS row of the slab, row_base + l
    const int row = row_base + lane;
    for (int s = 0; s < 8; ++s) {
      sq[s] = make_float2(0.f, 0.f);
      if (row < M && s < ep.stats_slots) sq[s] = ep.stats_in[static_cast<long>(s) * ep.ld_stats + row];
    }
  }
}
// ln_a / ln_b come back as packed (value, value) pairs for the f32x2 FMAs of the readback: entry `it` belongs to row
// row_base + (lane >> 3) + 4 * it.
template <int MODE>
__device__ __forceinline__ void epilogue_ln_coeffs(const GemmEpilogue& ep, const float2 (&sq)[8], int lane,
                                                   uint64_t (&ln_a)[8], uint64_t (&ln_b)[8]) {
  if constexpr (MODE == EPI_BF16_LN || MODE == EPI_BF16_GELU_LN) {
    float sx = 0.f, sy = 0.f;                    // slots in index order: the same bits every run
#pragma unroll
    for (int s = 0; s < 8; ++s) { sx += sq[s].x; sy += sq[s].y; }
    const float mean = sx * ep.ln_inv_dim;
    const float var = fmaxf(sy * ep.ln_inv_dim - mean * mean, 0.f);
    const float a = rsqrtf(var + ep.ln_eps);     // of row row_base + lane
    const float b = -mean * a;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int src = (lane >> 3) + 4 * it;
      const float ai = __shfl_sync(0xffffffffu, a, src), bi = __shfl_sync(0xffffffffu, b, src);
      ln_a[it] = pack2(ai, ai);
      ln_b[it] = pack2(bi, bi);
    }
  }
}

template <int MODE, typename Release>
__device__ __forceinline__ void epilogue_slab(const GemmEpilogue& ep, uint32_t taddr, uint32_t stage, int row_base,
                                              int col_base, int M, int N, int lane, const uint64_t (&ln_a)[8],
                                              const uint64_t (&ln_b)[8], Release&& release) {
  const int sub_row = lane >> 3;         // 0..3   readback: 4 rows per pass
  const int sub_chunk = lane & 7;        // 16-byte column chunk owned by this lane in the readback
  const uint32_t wr_base = stage + lane * 128;
  const int wr_sw = lane & 7;
  constexpr bool LN_IN = MODE == EPI_BF16_LN || MODE == EPI_BF16_GELU_LN;
  constexpr bool LN_OUT = MODE == EPI_F32_RESID_LN;
  // per-lane rows of the readback are row_base + sub_row + 4*it
  uint64_t st_s[LN_OUT ? 8 : 1], st_q[LN_OUT ? 8 : 1];   // producer: packed (even, odd column) partial sum / sum of squares
  if constexpr (LN_OUT) {
#pragma unroll
    for (int it = 0; it < 8; ++it) st_s[it] = st_q[it] = pack2(0.f, 0.f);
  }
  // residual modes: the residual of chunk c+1 is requested before chunk c is processed (different addresses from the
  // stores of chunk c, so the in-place update stays safe); otherwise every chunk exposes a full HBM/L2 round trip
  constexpr bool RESID = MODE == EPI_F32_RESID || LN_OUT;
  float4 rs_next[RESID ? 8 : 1];
  auto load_resid = [&](int chunk, float4 (&dst)[RESID ? 8 : 1]) {
    if constexpr (RESID) {
      const float* rp = ep.resid + static_cast<long>(row_base + sub_row) * ep.ld_resid + col_base + chunk * 32 + sub_chunk * 4;
      const long rstep = 4L * ep.ld_resid;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        dst[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_base + sub_row + 4 * it < M) dst[it] = *reinterpret_cast<const float4*>(rp + it * rstep);
      }
    }
  };
  load_resid(0, rs_next);
  // per-column vectors (bias, colsum) are likewise requested one chunk ahead
  float4 bias_next = make_float4(0.f, 0.f, 0.f, 0.f), cs_next = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load_cols = [&](int chunk) {
    const int c = col_base + chunk * 32 + sub_chunk * 4;
    if (ep.bias != nullptr) bias_next = __ldg(reinterpret_cast<const float4*>(ep.bias + c));
    if constexpr (LN_IN) cs_next = __ldg(reinterpret_cast<const float4*>(ep.colsum + c));
  };
  // The accumulator chunk of step c+1 is requested from TMEM as soon as chunk c has been staged to shared memory (its
  // registers are free again), so the TMEM round trip overlaps the readback / math / global stores of chunk c.
  uint32_t v[32];
  tmem_ld32(taddr, v);
#pragma unroll 1
  for (int chunk = 0; chunk < 4; ++chunk) {
    if (!HSENET_EPI_PREFETCH && chunk > 0) tmem_ld32(taddr + chunk * 32, v);
    load_cols(chunk);
    const float4 bias4 = bias_next;
    [[maybe_unused]] const float4 cs4 = cs_next;
    float4 rs[RESID ? 8 : 1];
    if constexpr (RESID) {
#pragma unroll
      for (int it = 0; it < 8; ++it) rs[it] = rs_next[it];
      if (chunk < 3) load_resid(chunk + 1, rs_next);
    }
    tmem_ld_wait();
    if (chunk == 3) release();           // last TMEM read of this slab: the accumulator stage can be reused
    __syncwarp();                        // previous block fully read back before it is overwritten
#pragma unroll
    for (int c = 0; c < 8; ++c)
      sts128(wr_base + ((c ^ wr_sw) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    if (HSENET_EPI_PREFETCH && chunk < 3) tmem_ld32(taddr + (chunk + 1) * 32, v);
    __syncwarp();
    const int col = col_base + chunk * 32 + sub_chunk * 4;
    const int row_first = row_base + sub_row;            // rows row_first + 4*it
    const uint32_t rd_base = stage + sub_row * 128;      // + it*512, chunk swizzled by (row & 7)

    if constexpr (MODE == EPI_BF16 || MODE == EPI_BF16_GELU || LN_IN) {
      __nv_bfloat16* o = ep.out_bf16 + static_cast<long>(row_first) * ep.ld_bf16 + col;
      const long step = 4L * ep.ld_bf16;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + sub_row;
        float4 x = lds128(rd_base + it * 512 + ((sub_chunk ^ (r & 7)) << 4));
        if constexpr (LN_IN) {      // y = rstd x + (-mean rstd) colsum + bias', two columns per f32x2 FMA
          unpack2(fma2(pack2(x.x, x.y), ln_a[it], fma2(ln_b[it], pack2(cs4.x, cs4.y), pack2(bias4.x, bias4.y))), x.x, x.y);
          unpack2(fma2(pack2(x.z, x.w), ln_a[it], fma2(ln_b[it], pack2(cs4.z, cs4.w), pack2(bias4.z, bias4.w))), x.z, x.w);
        } else {
          x.x += bias4.x; x.y += bias4.y; x.z += bias4.z; x.w += bias4.w;
        }
        if constexpr (MODE == EPI_BF16_GELU || MODE == EPI_BF16_GELU_LN) {
          if (ep.gelu == 2) {          // exact-erf form requested (A/B switch HSENET_GELU_ERF=1)
            x.x = gelu_fast(x.x); x.y = gelu_fast(x.y); x.z = gelu_fast(x.z); x.w = gelu_fast(x.w);
          } else {
            gelu_tanh2(x.x, x.y);
            gelu_tanh2(x.z, x.w);
          }
        }
        if (row_first + 4 * it < M) {
          uint2 pk;
          pk.x = pack_bf16x2(x.x, x.y);
          pk.y = pack_bf16x2(x.z, x.w);
          *reinterpret_cast<uint2*>(o + it * step) = pk;
        }
      }
    } else if constexpr (MODE == EPI_F32_RESID || LN_OUT) {
      float* o = ep.out_f32 + static_cast<long>(row_first) * ep.ld_f32 + col;
      const long ostep = 4L * ep.ld_f32;
      __nv_bfloat16* ob = nullptr;
      long bstep = 0;
      if constexpr (LN_OUT) {
        ob = ep.out_bf16 + static_cast<long>(row_first) * ep.ld_bf16 + col;
        bstep = 4L * ep.ld_bf16;
      }
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + sub_row;
        float4 x = lds128(rd_base + it * 512 + ((sub_chunk ^ (r & 7)) << 4));
        x.x += bias4.x + rs[it].x; x.y += bias4.y + rs[it].y; x.z += bias4.z + rs[it].z; x.w += bias4.w + rs[it].w;
        if (row_first + 4 * it < M) {
          *reinterpret_cast<float4*>(o + it * ostep) = x;
          if constexpr (LN_OUT) {
            uint2 pk;
            pk.x = pack_bf16x2(x.x, x.y);
            pk.y = pack_bf16x2(x.z, x.w);
            *reinterpret_cast<uint2*>(ob + it * bstep) = pk;
            const uint64_t xy = pack2(x.x, x.y), zw = pack2(x.z, x.w);
            st_s[it] = add2(st_s[it], add2(xy, zw));
            st_q[it] = fma2(zw, zw, fma2(xy, xy, st_q[it]));
          }
        }
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
          if (ep.dgelu_src != nullptr) {      // backward of GELU: multiply by gelu'(saved pre-activation)
            const uint2 u = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(ep.dgelu_src) +
                                                            orow[it] * ep.ld_dgelu + col);
            const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
            const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
            x.x *= gelu_erf_grad(a.x); x.y *= gelu_erf_grad(a.y); x.z *= gelu_erf_grad(b.x); x.w *= gelu_erf_grad(b.y);
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
  if constexpr (LN_OUT) {
    // The 8 lanes that share sub_row hold partial sums of the same 8 rows (16 columns each).  Recursive halving: every
    // exchange sends half of the values still held and adds the half received (8 + 4 + 2 = 14 shuffles instead of a
    // 48-shuffle butterfly); lane (b2 b1 b0) ends up with the totals of row `it` = 4*b2 + 2*b1 + b0 and adds them.
    float v[16];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      float lo, hi;
      unpack2(st_s[it], lo, hi);
      v[2 * it] = lo + hi;
      unpack2(st_q[it], lo, hi);
      v[2 * it + 1] = lo + hi;
    }
    const bool b2 = (lane & 4) != 0, b1 = (lane & 2) != 0, b0 = (lane & 1) != 0;
    float w8[8], w4[4], w2[2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = b2 ? v[i] : v[i + 8];
      w8[i] = (b2 ? v[i + 8] : v[i]) + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = b1 ? w8[i] : w8[i + 4];
      w4[i] = (b1 ? w8[i + 4] : w8[i]) + __shfl_xor_sync(0xffffffffu, send, 2);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = b0 ? w4[i] : w4[i + 2];
      w2[i] = (b0 ? w4[i + 2] : w4[i]) + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    const int row = row_base + sub_row + 4 * (lane & 7);
    if (row < M) ep.stats_out[static_cast<long>(col_base >> 7) * ep.ld_stats + row] = make_float2(w2[0], w2[1]);
  }
}

}  // namespace hs
