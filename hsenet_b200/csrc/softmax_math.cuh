// exp2 helpers shared by the attention kernels.
#pragma once
#include "common.cuh"

namespace hs {

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^y for two values on the FMA / ALU pipes (no MUFU): y = n + f with n = round(y), f in [-0.5, 0.5]; cubic for 2^f
// (relative error < 7e-4, inside bf16's 2^-9 rounding); the exponent is patched in with integer adds.  Used for a
// fraction of the exponentials: the MUFU pipe (16 ex2/clk/SM) is what bounds the softmax.
__device__ __forceinline__ void exp2_poly2(float y0, float y1, float& e0, float& e1) {
  y0 = fmaxf(y0, -126.0f);
  y1 = fmaxf(y1, -126.0f);
  const uint64_t magic = pack2(12582912.0f, 12582912.0f);        // 1.5 * 2^23
  const uint64_t nmagic = pack2(-12582912.0f, -12582912.0f);
  const uint64_t y = pack2(y0, y1);
  const uint64_t t = add2(y, magic);                             // integer part lands in the low mantissa bits
  const uint64_t n = add2(t, nmagic);
  const uint64_t f = fma2(n, pack2(-1.0f, -1.0f), y);            // f = y - n
  uint64_t p = fma2(f, pack2(0.0555041f, 0.0555041f), pack2(0.2402265f, 0.2402265f));
  p = fma2(p, f, pack2(0.6931472f, 0.6931472f));
  p = fma2(p, f, pack2(1.0f, 1.0f));
  float p0, p1, t0, t1;
  unpack2(p, p0, p1);
  unpack2(t, t0, t1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

}  // namespace hs
