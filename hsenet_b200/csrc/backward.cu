// Row / elementwise kernels of the training path: everything of the backward pass that is
// not a dense contraction.  All reductions over rows (bias, LayerNorm gain / shift, positional embedding, slice keys)
// are two-stage with a fixed summation order -- no atomics, so gradients are bit-reproducible.
//
//   transpose_pad      X [M,N] -> X^T [N,Mpad] (zero padded to a multiple of 128 rows) in the activation dtype: the weight
//                      gradient dW = dY^T X is run by the same tcgen05 GEMM as the forward (C = A W^T with both operands
//                      K-major), so both factors are needed with the token dimension contiguous.  Optionally applies GELU
//                      on the way (recomputes the MLP hidden activation), writes a straight cast copy, and emits per-tile
//                      column sums (bias gradients).
//   layernorm_bwd      dx (+)= rstd (g - mean(g) - xhat mean(g xhat)), g = dy gamma; per-block partial dgamma / dbeta
//   gelu_rows          H2 = gelu(H1) (training forward keeps the pre-activation)
//   attention_rowdot   D = rowsum(dO o O) per head
//   window_attn_bwd, slice_xattn_bwd_*, score_scale_bwd, pool_bwd, sum_over_batch, combine_final_grad
#include "common.cuh"
#include "kernels.h"

namespace hs {

extern void count_launch();

namespace {

inline int launch_status() { return cudaGetLastError() == cudaSuccess ? HS_OK : HS_ERR_CUDA; }

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

template <typename T>
struct V4;
template <>
struct V4<float> {
  static __device__ __forceinline__ float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <>
struct V4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 load(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float4 v) {
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

constexpr int kVec = kHidden / 128;   // float4 per lane per 768-wide row

// ---------------------------------------------------------------------------------------------------------------------
// transpose (+ optional GELU, cast copy, column sums)
// ---------------------------------------------------------------------------------------------------------------------
template <typename TIn, typename TOut, int OP>
__global__ void __launch_bounds__(256) transpose_pad_kernel(const TIn* __restrict__ in, long ld_in, int M, int N,
                                                            TOut* __restrict__ out_t, int Mpad,
                                                            TOut* __restrict__ copy_out, long ld_copy,
                                                            float* __restrict__ colsum_partial) {
  __shared__ float tile[64][65];
  __shared__ float red[4][64];
  const int n0 = blockIdx.x * 64, m0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  float cs = 0.f;
#pragma unroll 4
  for (int r = ty; r < 64; r += 4) {
    const int m = m0 + r;
    float v = 0.f;
    if (m < M) {
      v = to_f<TIn>(in[static_cast<long>(m) * ld_in + n0 + tx]);
      if (OP == 1) v = gelu_erf(v);
      if (copy_out != nullptr) copy_out[static_cast<long>(m) * ld_copy + n0 + tx] = from_f<TOut>(v);
      // the column sum is taken over the values the tensor cores will see (after rounding to the activation dtype)
      v = to_f<TOut>(from_f<TOut>(v));
    }
    tile[r][tx] = v;
    cs += v;
  }
  if (colsum_partial != nullptr) red[ty][tx] = cs;
  __syncthreads();
  if (colsum_partial != nullptr && ty == 0)
    colsum_partial[static_cast<long>(blockIdx.y) * N + n0 + tx] = (red[0][tx] + red[1][tx]) + (red[2][tx] + red[3][tx]);
  if (out_t != nullptr) {
#pragma unroll 4
    for (int r = ty; r < 64; r += 4)       // output row n0 + r, 64 consecutive m
      out_t[static_cast<long>(n0 + r) * Mpad + m0 + tx] = from_f<TOut>(tile[tx][r]);
  }
}

// out[n] = sum_t partial[t][n], T rows of N columns, fixed summation order (deterministic).
// Tall case (bias / LayerNorm partials: T ~ 257, N <= 3072): block = 32 columns x 32 row groups, each thread sums its
// contiguous slice of t (loads unrolled so they overlap), the 32 slice sums are added in index order.  (A first version
// walked all T rows in one thread: 29 us of dependent DRAM latency per call, 8 % of a training step.)
__global__ void __launch_bounds__(1024) colsum_finish_tall_kernel(const float* __restrict__ partial, int T, int N,
                                                                  float* __restrict__ out) {
  __shared__ float red[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int per = (T + 31) / 32;
  const int t0 = ty * per, t1 = min(T, t0 + per);
  float s = 0.f;
  if (n < N) {
#pragma unroll 4
    for (int t = t0; t < t1; ++t) s += __ldg(partial + static_cast<long>(t) * N + n);
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float a = 0.f;
#pragma unroll
    for (int g = 0; g < 32; ++g) a += red[g][tx];
    out[n] = a;
  }
}
// Flat case (split-K partials of a weight gradient: T <= 8 slices of up to 2.4 M elements): one float4 per thread
__global__ void __launch_bounds__(256) colsum_finish_flat_kernel(const float* __restrict__ partial, int T, long n4,
                                                                 float* __restrict__ out) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    float4 a = __ldg(reinterpret_cast<const float4*>(partial) + i);
    for (int t = 1; t < T; ++t) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(partial) + t * n4 + i);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = a;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// GELU (exact erf), elementwise
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void gelu_rows_kernel(const T* __restrict__ in, T* __restrict__ out, long n4) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    float4 v = V4<T>::load(in + i * 4);
    v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
    V4<T>::store(out + i * 4, v);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// LayerNorm backward.  One warp per row, kLnRowsPerWarp rows per warp, 8 warps per block.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kLnRowsPerWarp = 4;
constexpr int kLnRowsPerBlock = 8 * kLnRowsPerWarp;
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            long ldx, const float* __restrict__ gamma, long rows,
                                                            float* __restrict__ dx, long ld_dx, int accumulate,
                                                            float* __restrict__ dgamma_partial,
                                                            float* __restrict__ dbeta_partial) {
  __shared__ float red[8][kHidden];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 ag[kVec], ab[kVec];
#pragma unroll
  for (int i = 0; i < kVec; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 g4[kVec];
#pragma unroll
  for (int i = 0; i < kVec; ++i) g4[i] = __ldg(reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4));
  for (int rr = 0; rr < kLnRowsPerWarp; ++rr) {
    const long row = static_cast<long>(blockIdx.x) * kLnRowsPerBlock + warp * kLnRowsPerWarp + rr;
    if (row >= rows) break;
    float4 xv[kVec], gv[kVec];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      xv[i] = *reinterpret_cast<const float4*>(x + row * ldx + (i * 32 + lane) * 4);
      s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    }
    const float mean = warp_sum(s) * (1.0f / kHidden);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
      q += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / kHidden) + kLnEps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const float4 d = *reinterpret_cast<const float4*>(dy + row * kHidden + (i * 32 + lane) * 4);
      xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;      // xhat
      ag[i].x += d.x * xv[i].x; ag[i].y += d.y * xv[i].y; ag[i].z += d.z * xv[i].z; ag[i].w += d.w * xv[i].w;
      ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
      gv[i] = make_float4(d.x * g4[i].x, d.y * g4[i].y, d.z * g4[i].z, d.w * g4[i].w);
      sg += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
      sgx += (gv[i].x * xv[i].x + gv[i].y * xv[i].y) + (gv[i].z * xv[i].z + gv[i].w * xv[i].w);
    }
    const float mg = warp_sum(sg) * (1.0f / kHidden), mgx = warp_sum(sgx) * (1.0f / kHidden);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      float* o = dx + row * ld_dx + (i * 32 + lane) * 4;
      float4 r;
      r.x = rstd * (gv[i].x - mg - xv[i].x * mgx);
      r.y = rstd * (gv[i].y - mg - xv[i].y * mgx);
      r.z = rstd * (gv[i].z - mg - xv[i].z * mgx);
      r.w = rstd * (gv[i].w - mg - xv[i].w * mgx);
      if (accumulate) {
        const float4 p = *reinterpret_cast<const float4*>(o);
        r.x += p.x; r.y += p.y; r.z += p.z; r.w += p.w;
      }
      *reinterpret_cast<float4*>(o) = r;
    }
  }
  // block partials of dgamma, then dbeta (fixed order: warps 0..7)
  for (int pass = 0; pass < 2; ++pass) {
    float* dst = pass == 0 ? dgamma_partial : dbeta_partial;
    if (dst == nullptr) continue;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kVec; ++i)
      *reinterpret_cast<float4*>(&red[warp][(i * 32 + lane) * 4]) = pass == 0 ? ag[i] : ab[i];
    __syncthreads();
    for (int c = threadIdx.x; c < kHidden; c += 256) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][c];
      dst[static_cast<long>(blockIdx.x) * kHidden + c] = t;
    }
  }
}

// dy [B*seq,768] fp32 = d_tokens (+ d_patch shifted behind the cls row); either may be null
template <typename T>
__global__ void __launch_bounds__(256) combine_final_grad_kernel(const T* __restrict__ d_tokens,
                                                                 const T* __restrict__ d_patch, int seq, long rows,
                                                                 float* __restrict__ dy) {
  const long row = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const long b = row / seq;
  const int t = static_cast<int>(row - b * seq);
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int c = (i * 32 + lane) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (d_tokens != nullptr) v = V4<T>::load(d_tokens + row * kHidden + c);
    if (d_patch != nullptr && t >= 1) {
      const float4 p = V4<T>::load(d_patch + (b * (seq - 1) + (t - 1)) * kHidden + c);
      v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    *reinterpret_cast<float4*>(dy + row * kHidden + c) = v;
  }
}

// out[r, :] = sum_b in[b * batch_stride + r * 768 + :]   (positional-embedding / cls-token gradients)
__global__ void __launch_bounds__(256) sum_over_batch_kernel(const float* __restrict__ in, long batch_stride, int B,
                                                             int rows, float* __restrict__ out) {
  const long r = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int c = (i * 32 + lane) * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
      const float4 v = *reinterpret_cast<const float4*>(in + b * batch_stride + r * kHidden + c);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    *reinterpret_cast<float4*>(out + r * kHidden + c) = a;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// attention: D[b,h,q] = sum_d dO o O  (padded to S_pad with zeros)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) attention_rowdot_kernel(const T* __restrict__ out, const T* __restrict__ d_out,
                                                               float* __restrict__ dvec, int S, int sp) {
  const int b = blockIdx.z, h = blockIdx.y;
  const int q = blockIdx.x * 128 + threadIdx.x;
  float d = 0.f;
  if (q < S) {
    const long off = (static_cast<long>(b) * S + q) * kHidden + h * kHeadDim;
#pragma unroll
    for (int i = 0; i < kHeadDim / 4; ++i) {
      const float4 a = V4<T>::load(out + off + i * 4), g = V4<T>::load(d_out + off + i * 4);
      d += (a.x * g.x + a.y * g.y) + (a.z * g.z + a.w * g.w);
    }
  }
  dvec[(static_cast<long>(b) * kHeads + h) * sp + q] = d;
}

// ---------------------------------------------------------------------------------------------------------------------
// packer: backward of the per-window 1 x 16 attention and of the (1,4,4) average pooling
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int window_member_b(int n, int e) {     // == rowops.cu window_member (the bit-exact map)
  const int dz = n >> 4, wy = (n >> 2) & 3, hx = n & 3;
  const int sw = e >> 2, sh = e & 3;
  return dz * 256 + (4 * wy + sw) * 16 + (4 * hx + sh);
}

// dO [B*128,768] T, Q [B*128,768] fp32, KV [B*2048,1536] T  ->  dQ [B*128,768] fp32, dKV [B*2048,1536] T.  Warp per window.
template <typename T>
__global__ void __launch_bounds__(256) window_attn_bwd_kernel(const T* __restrict__ dO, const float* __restrict__ Q,
                                                              const T* __restrict__ KV, float* __restrict__ dQ,
                                                              T* __restrict__ dKV, long total, DropSpec dr) {
  const long w = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (w >= total) return;
  const int lane = threadIdx.x & 31;
  const long b = w >> 7;
  const int n = static_cast<int>(w & 127);
  float4 q[kVec], go[kVec];
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    q[i] = *reinterpret_cast<const float4*>(Q + w * kHidden + (i * 32 + lane) * 4);
    go[i] = V4<T>::load(dO + w * kHidden + (i * 32 + lane) * 4);
  }
  float sc[16], dp[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const T* kr = KV + (b * kNPatch + window_member_b(n, e)) * (2 * kHidden);
    float d = 0.f, g = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const float4 k = V4<T>::load(kr + (i * 32 + lane) * 4);
      const float4 v = V4<T>::load(kr + kHidden + (i * 32 + lane) * 4);
      d += (q[i].x * k.x + q[i].y * k.y) + (q[i].z * k.z + q[i].w * k.w);
      g += (go[i].x * v.x + go[i].y * v.y) + (go[i].z * v.z + go[i].w * v.w);
    }
    sc[e] = warp_sum(d) * 0.036084391824351615f;
    dp[e] = warp_sum(g);
  }
  float m = sc[0];
#pragma unroll
  for (int e = 1; e < 16; ++e) m = fmaxf(m, sc[e]);
  float den = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) { sc[e] = expf(sc[e] - m); den += sc[e]; }
  const float inv = 1.0f / den;
  // train-mode dropout: O = (p o M) V, so dP = M o (dO V^T) and dV uses the dropped probabilities; mk[e] = M_e (1 when off)
  float mk[16];
#pragma unroll
  for (int e = 0; e < 16; ++e)
    mk[e] = dr.on() ? drop_mask(dr.seed, dr.keep_thresh, dr.scale, static_cast<unsigned long long>(w) * 16 + e) : 1.0f;
  float pd = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) { sc[e] *= inv; dp[e] *= mk[e]; pd += sc[e] * dp[e]; }
  float4 dq[kVec];
#pragma unroll
  for (int i = 0; i < kVec; ++i) dq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const long tok = b * kNPatch + window_member_b(n, e);
    const T* kr = KV + tok * (2 * kHidden);
    T* dk = dKV + tok * (2 * kHidden);
    const float ds = sc[e] * (dp[e] - pd) * 0.036084391824351615f;
    const float p = sc[e] * mk[e];
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 k = V4<T>::load(kr + c);
      dq[i].x += ds * k.x; dq[i].y += ds * k.y; dq[i].z += ds * k.z; dq[i].w += ds * k.w;
      V4<T>::store(dk + c, make_float4(ds * q[i].x, ds * q[i].y, ds * q[i].z, ds * q[i].w));
      V4<T>::store(dk + kHidden + c, make_float4(p * go[i].x, p * go[i].y, p * go[i].z, p * go[i].w));
    }
  }
#pragma unroll
  for (int i = 0; i < kVec; ++i) *reinterpret_cast<float4*>(dQ + w * kHidden + (i * 32 + lane) * 4) = dq[i];
}

// dHR[b, token, :] (+)= dLR[b, window(token), :] / 16.  One warp per HR token.
template <typename TG>
__global__ void __launch_bounds__(256) pool_bwd_kernel(const TG* __restrict__ dLR, float* __restrict__ dHR, long total,
                                                       int accumulate) {
  const long r = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);   // b*2048 + token
  if (r >= total) return;
  const int lane = threadIdx.x & 31;
  const long b = r >> 11;
  const int t = static_cast<int>(r & 2047);
  const int dz = t >> 8, y = (t >> 4) & 15, x = t & 15;
  const int n = dz * 16 + (y >> 2) * 4 + (x >> 2);
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int c = (i * 32 + lane) * 4;
    float4 g = V4<TG>::load(dLR + (b * 128 + n) * kHidden + c);
    g.x *= 0.0625f; g.y *= 0.0625f; g.z *= 0.0625f; g.w *= 0.0625f;
    float* o = dHR + r * kHidden + c;
    if (accumulate) {
      const float4 p = *reinterpret_cast<const float4*>(o);
      g.x += p.x; g.y += p.y; g.z += p.z; g.w += p.w;
    }
    *reinterpret_cast<float4*>(o) = g;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// 2E3 slice-guided scoring, backward
// ---------------------------------------------------------------------------------------------------------------------
// Row part of the slice cross attention: recompute p = softmax(q K^T / sqrt(768)) against the 32 slice keys, then
//   dp_j = dO . V_j,  ds_j = p_j (dp_j - sum p dp) / sqrt(768),  dQ (+)= sum_j ds_j K_j;  P and dS go to [Mp,32] for the
// key/value reduction below.  One warp per query row; K and V of the volume are read through L1/L2 (196 KB per volume).
template <typename T>
__global__ void __launch_bounds__(256) slice_xattn_bwd_rows_kernel(const float* __restrict__ Q,
                                                                   const float* __restrict__ KV,
                                                                   const T* __restrict__ dO, float* __restrict__ dQ,
                                                                   int accumulate, float* __restrict__ P,
                                                                   float* __restrict__ dS, long total, DropSpec dr) {
  const long r = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= total) return;
  const int lane = threadIdx.x & 31;
  const long b = r >> 11;
  float4 q[kVec], go[kVec];
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    q[i] = *reinterpret_cast<const float4*>(Q + r * kHidden + (i * 32 + lane) * 4);
    go[i] = V4<T>::load(dO + r * kHidden + (i * 32 + lane) * 4);
  }
  float sc = 0.f, dp = 0.f;                   // lane j holds key j
  for (int j = 0; j < kNSlice; ++j) {
    const float* kr = KV + (b * kNSlice + j) * (2 * kHidden);
    float d = 0.f, g = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const float4 k = __ldg(reinterpret_cast<const float4*>(kr + (i * 32 + lane) * 4));
      const float4 v = __ldg(reinterpret_cast<const float4*>(kr + kHidden + (i * 32 + lane) * 4));
      d += (q[i].x * k.x + q[i].y * k.y) + (q[i].z * k.z + q[i].w * k.w);
      g += (go[i].x * v.x + go[i].y * v.y) + (go[i].z * v.z + go[i].w * v.w);
    }
    d = warp_sum(d);
    g = warp_sum(g);
    if (lane == j) { sc = d * 0.036084391824351615f; dp = g; }
  }
  const float m = warp_max(sc);
  const float e = expf(sc - m);
  const float p = e / warp_sum(e);
  // train-mode dropout: O = (p o M) V  ->  dP = M o (dO V^T); the key/value reduction below needs the DROPPED p for dV
  const float mk = dr.on() ? drop_mask(dr.seed, dr.keep_thresh, dr.scale, static_cast<unsigned long long>(r) * kNSlice + lane)
                           : 1.0f;
  dp *= mk;
  const float pd = warp_sum(p * dp);
  const float ds = p * (dp - pd) * 0.036084391824351615f;
  P[r * kNSlice + lane] = p * mk;
  dS[r * kNSlice + lane] = ds;
  float4 acc[kVec];
#pragma unroll
  for (int i = 0; i < kVec; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = 0; j < kNSlice; ++j) {
    const float dsj = __shfl_sync(0xffffffffu, ds, j);
    const float* kr = KV + (b * kNSlice + j) * (2 * kHidden);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const float4 k = __ldg(reinterpret_cast<const float4*>(kr + (i * 32 + lane) * 4));
      acc[i].x += dsj * k.x; acc[i].y += dsj * k.y; acc[i].z += dsj * k.z; acc[i].w += dsj * k.w;
    }
  }
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    float* o = dQ + r * kHidden + (i * 32 + lane) * 4;
    if (accumulate) {
      const float4 pv = *reinterpret_cast<const float4*>(o);
      acc[i].x += pv.x; acc[i].y += pv.y; acc[i].z += pv.z; acc[i].w += pv.w;
    }
    *reinterpret_cast<float4*>(o) = acc[i];
  }
}

// dK_j = sum_t dS[t,j] Q_t,  dV_j = sum_t P[t,j] dO_t  over the 2048 queries of a volume: one thread per output column,
// 32 accumulators in registers, queries in index order (deterministic).  grid (1536 / 256, B).
template <typename T>
__global__ void __launch_bounds__(256) slice_xattn_bwd_kv_kernel(const float* __restrict__ Q, const T* __restrict__ dO,
                                                                 const float* __restrict__ P,
                                                                 const float* __restrict__ dS,
                                                                 float* __restrict__ dKV) {
  __shared__ float sw[64][kNSlice];
  const int b = blockIdx.y;
  const int col = blockIdx.x * 256 + threadIdx.x;         // 0..1535: K columns then V columns
  const bool is_v = col >= kHidden;
  const int c = is_v ? col - kHidden : col;
  const float* W = is_v ? P : dS;
  float acc[kNSlice];
#pragma unroll
  for (int j = 0; j < kNSlice; ++j) acc[j] = 0.f;
  for (int t0 = 0; t0 < kNPatch; t0 += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * kNSlice; i += 256)
      sw[i / kNSlice][i % kNSlice] = W[(static_cast<long>(b) * kNPatch + t0) * kNSlice + i];
    __syncthreads();
    for (int tt = 0; tt < 64; ++tt) {
      const long row = static_cast<long>(b) * kNPatch + t0 + tt;
      const float x = is_v ? to_f<T>(dO[row * kHidden + c]) : Q[row * kHidden + c];
#pragma unroll
      for (int j = 0; j < kNSlice; ++j) acc[j] = fmaf(sw[tt][j], x, acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < kNSlice; ++j) dKV[(static_cast<long>(b) * kNSlice + j) * (2 * kHidden) + col] = acc[j];
}

// Backward of score_scale_kernel (rowops.cu): X[b,1+t] = XP[t] * sigmoid(LN(Z[t]) . w_s + b_s).
//   dXs = dX[b,1+t,:] (fp32);  dXP = dXs * score;  dlogit = (dXs . XP) score (1 - score);  dLN = dlogit w_s;
//   dZ = LayerNorm backward of dLN;  block partials of dgamma, dbeta, dw_s (= dlogit * LNout) and db_s.
__global__ void __launch_bounds__(256) score_scale_bwd_kernel(const float* __restrict__ dX, const float* __restrict__ XP,
                                                              const float* __restrict__ Z, const float* __restrict__ g,
                                                              const float* __restrict__ be,
                                                              const float* __restrict__ ws,
                                                              const float* __restrict__ scores,
                                                              float* __restrict__ dXP, float* __restrict__ dZ,
                                                              float* __restrict__ dg_partial,
                                                              float* __restrict__ db_partial,
                                                              float* __restrict__ dws_partial,
                                                              float* __restrict__ dbs_partial, long total) {
  __shared__ float red[8][kHidden];
  __shared__ float red1[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 ag[kVec], ab[kVec], aw[kVec];
#pragma unroll
  for (int i = 0; i < kVec; ++i) ag[i] = ab[i] = aw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float abs_ = 0.f;
  for (int rr = 0; rr < kLnRowsPerWarp; ++rr) {
    const long row = static_cast<long>(blockIdx.x) * kLnRowsPerBlock + warp * kLnRowsPerWarp + rr;
    if (row >= total) break;
    const long b = row >> 11, t = row & 2047;
    const float* dxs = dX + (b * kSeq + 1 + t) * kHidden;
    const float score = scores[row];
    float4 zv[kVec], gx[kVec];
    float s = 0.f, dot = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int c = (i * 32 + lane) * 4;
      zv[i] = *reinterpret_cast<const float4*>(Z + row * kHidden + c);
      s += (zv[i].x + zv[i].y) + (zv[i].z + zv[i].w);
      gx[i] = *reinterpret_cast<const float4*>(dxs + c);
      const float4 xp = *reinterpret_cast<const float4*>(XP + row * kHidden + c);
      dot += (gx[i].x * xp.x + gx[i].y * xp.y) + (gx[i].z * xp.z + gx[i].w * xp.w);
      *reinterpret_cast<float4*>(dXP + row * kHidden + c) =
          make_float4(gx[i].x * score, gx[i].y * score, gx[i].z * score, gx[i].w * score);
    }
    const float mean = warp_sum(s) * (1.0f / kHidden);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      zv[i].x -= mean; zv[i].y -= mean; zv[i].z -= mean; zv[i].w -= mean;
      q += (zv[i].x * zv[i].x + zv[i].y * zv[i].y) + (zv[i].z * zv[i].z + zv[i].w * zv[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / kHidden) + kLnEps);
    const float dlogit = warp_sum(dot) * score * (1.0f - score);
    if (lane == 0) abs_ += dlogit;
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(be + c));
      const float4 w = __ldg(reinterpret_cast<const float4*>(ws + c));
      zv[i].x *= rstd; zv[i].y *= rstd; zv[i].z *= rstd; zv[i].w *= rstd;       // xhat
      const float4 d = make_float4(dlogit * w.x, dlogit * w.y, dlogit * w.z, dlogit * w.w);   // dLN
      ag[i].x += d.x * zv[i].x; ag[i].y += d.y * zv[i].y; ag[i].z += d.z * zv[i].z; ag[i].w += d.w * zv[i].w;
      ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
      aw[i].x += dlogit * (zv[i].x * gg.x + bb.x); aw[i].y += dlogit * (zv[i].y * gg.y + bb.y);
      aw[i].z += dlogit * (zv[i].z * gg.z + bb.z); aw[i].w += dlogit * (zv[i].w * gg.w + bb.w);
      gx[i] = make_float4(d.x * gg.x, d.y * gg.y, d.z * gg.z, d.w * gg.w);
      sg += (gx[i].x + gx[i].y) + (gx[i].z + gx[i].w);
      sgx += (gx[i].x * zv[i].x + gx[i].y * zv[i].y) + (gx[i].z * zv[i].z + gx[i].w * zv[i].w);
    }
    const float mg = warp_sum(sg) * (1.0f / kHidden), mgx = warp_sum(sgx) * (1.0f / kHidden);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      float4 r;
      r.x = rstd * (gx[i].x - mg - zv[i].x * mgx);
      r.y = rstd * (gx[i].y - mg - zv[i].y * mgx);
      r.z = rstd * (gx[i].z - mg - zv[i].z * mgx);
      r.w = rstd * (gx[i].w - mg - zv[i].w * mgx);
      *reinterpret_cast<float4*>(dZ + row * kHidden + (i * 32 + lane) * 4) = r;
    }
  }
  for (int pass = 0; pass < 3; ++pass) {
    float* dst = pass == 0 ? dg_partial : (pass == 1 ? db_partial : dws_partial);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kVec; ++i)
      *reinterpret_cast<float4*>(&red[warp][(i * 32 + lane) * 4]) = pass == 0 ? ag[i] : (pass == 1 ? ab[i] : aw[i]);
    __syncthreads();
    for (int c = threadIdx.x; c < kHidden; c += 256) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][c];
      dst[static_cast<long>(blockIdx.x) * kHidden + c] = t;
    }
  }
  if (lane == 0) red1[warp] = abs_;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red1[w];
    dbs_partial[blockIdx.x] = t;
  }
}

// out (+)= a  over n4 float4
__global__ void add_rows_kernel(float* __restrict__ out, const float* __restrict__ a, long n4) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    float4 o = reinterpret_cast<float4*>(out)[i];
    const float4 v = reinterpret_cast<const float4*>(a)[i];
    o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

inline unsigned grid_1d(long n, int per_block) {
  long b = (n + per_block - 1) / per_block;
  const long cap = 148L * 16;
  return static_cast<unsigned>(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------------
int padded_rows(int M) { return (M + 127) / 128 * 128; }

template <typename TIn, typename TOut>
int transpose_pad(const TIn* in, long ld_in, int M, int N, TOut* out_t, TOut* copy_out, long ld_copy, int gelu,
                  float* colsum_partial, cudaStream_t st) {
  if (M <= 0 || N <= 0) return HS_OK;
  if (N % 64) return HS_ERR_SHAPE;
  const int Mpad = padded_rows(M);
  const dim3 grid(N / 64, Mpad / 64);
  if (gelu)
    transpose_pad_kernel<TIn, TOut, 1><<<grid, 256, 0, st>>>(in, ld_in, M, N, out_t, Mpad, copy_out, ld_copy, colsum_partial);
  else
    transpose_pad_kernel<TIn, TOut, 0><<<grid, 256, 0, st>>>(in, ld_in, M, N, out_t, Mpad, copy_out, ld_copy, colsum_partial);
  count_launch();
  return launch_status();
}
template int transpose_pad<float, float>(const float*, long, int, int, float*, float*, long, int, float*, cudaStream_t);
template int transpose_pad<float, __nv_bfloat16>(const float*, long, int, int, __nv_bfloat16*, __nv_bfloat16*, long, int,
                                                 float*, cudaStream_t);
template int transpose_pad<__nv_bfloat16, __nv_bfloat16>(const __nv_bfloat16*, long, int, int, __nv_bfloat16*,
                                                         __nv_bfloat16*, long, int, float*, cudaStream_t);

int colsum_finish(const float* partial, int T, int N, float* out, cudaStream_t st) {
  if (N <= 0) return HS_OK;
  if (T <= 8 && N % 4 == 0 && N >= 4096) {
    const long n4 = N / 4;
    long blocks = (n4 + 255) / 256;
    if (blocks > 148L * 16) blocks = 148L * 16;
    colsum_finish_flat_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(partial, T, n4, out);
  } else {
    colsum_finish_tall_kernel<<<(N + 31) / 32, 1024, 0, st>>>(partial, T, N, out);
  }
  count_launch();
  return launch_status();
}

template <typename T>
int gelu_rows(const T* in, T* out, long n, cudaStream_t st) {
  if (n <= 0) return HS_OK;
  if (n % 4) return HS_ERR_SHAPE;
  gelu_rows_kernel<T><<<grid_1d(n / 4, 256), 256, 0, st>>>(in, out, n / 4);
  count_launch();
  return launch_status();
}
template int gelu_rows<float>(const float*, float*, long, cudaStream_t);
template int gelu_rows<__nv_bfloat16>(const __nv_bfloat16*, __nv_bfloat16*, long, cudaStream_t);

int layernorm_bwd_blocks(long rows) { return static_cast<int>((rows + kLnRowsPerBlock - 1) / kLnRowsPerBlock); }

int layernorm_bwd(const float* dy, const float* x, long ldx, const float* gamma, long rows, float* dx, long ld_dx,
                  int accumulate, float* dgamma_partial, float* dbeta_partial, cudaStream_t st) {
  if (rows <= 0) return HS_OK;
  layernorm_bwd_kernel<<<layernorm_bwd_blocks(rows), 256, 0, st>>>(dy, x, ldx, gamma, rows, dx, ld_dx, accumulate,
                                                                   dgamma_partial, dbeta_partial);
  count_launch();
  return launch_status();
}

template <typename T>
int combine_final_grad(const T* d_tokens, const T* d_patch, int B, int seq, float* dy, cudaStream_t st) {
  const long rows = static_cast<long>(B) * seq;
  if (rows <= 0) return HS_OK;
  combine_final_grad_kernel<T><<<static_cast<unsigned>((rows + 7) / 8), 256, 0, st>>>(d_tokens, d_patch, seq, rows, dy);
  count_launch();
  return launch_status();
}
template int combine_final_grad<float>(const float*, const float*, int, int, float*, cudaStream_t);
template int combine_final_grad<__nv_bfloat16>(const __nv_bfloat16*, const __nv_bfloat16*, int, int, float*, cudaStream_t);

int sum_over_batch(const float* in, long batch_stride, int B, int rows, float* out, cudaStream_t st) {
  if (rows <= 0) return HS_OK;
  sum_over_batch_kernel<<<(rows + 7) / 8, 256, 0, st>>>(in, batch_stride, B, rows, out);
  count_launch();
  return launch_status();
}

int attention_rowdot_bf16(const __nv_bfloat16* out, const __nv_bfloat16* d_out, float* dvec, int B, int S,
                          cudaStream_t st) {
  const int sp = padded_rows(S);
  attention_rowdot_kernel<__nv_bfloat16><<<dim3(sp / 128, kHeads, B), 128, 0, st>>>(out, d_out, dvec, S, sp);
  count_launch();
  return launch_status();
}
int attention_rowdot_f32(const float* out, const float* d_out, float* dvec, int B, int S, cudaStream_t st) {
  const int sp = padded_rows(S);
  attention_rowdot_kernel<float><<<dim3(sp / 128, kHeads, B), 128, 0, st>>>(out, d_out, dvec, S, sp);
  count_launch();
  return launch_status();
}

template <typename T>
int window_attn_bwd(const T* dO, const float* Q, const T* KV, float* dQ, T* dKV, int B, cudaStream_t st, DropSpec dr) {
  const long total = static_cast<long>(B) * 128;
  if (total <= 0) return HS_OK;
  window_attn_bwd_kernel<T><<<static_cast<unsigned>((total + 7) / 8), 256, 0, st>>>(dO, Q, KV, dQ, dKV, total, dr);
  count_launch();
  return launch_status();
}
template int window_attn_bwd<float>(const float*, const float*, const float*, float*, float*, int, cudaStream_t, DropSpec);
template int window_attn_bwd<__nv_bfloat16>(const __nv_bfloat16*, const float*, const __nv_bfloat16*, float*,
                                            __nv_bfloat16*, int, cudaStream_t, DropSpec);

template <typename TG>
int pool_bwd(const TG* dLR, float* dHR, int B, int accumulate, cudaStream_t st) {
  const long total = static_cast<long>(B) * kNPatch;
  if (total <= 0) return HS_OK;
  pool_bwd_kernel<TG><<<static_cast<unsigned>((total + 7) / 8), 256, 0, st>>>(dLR, dHR, total, accumulate);
  count_launch();
  return launch_status();
}
template int pool_bwd<float>(const float*, float*, int, int, cudaStream_t);
template int pool_bwd<__nv_bfloat16>(const __nv_bfloat16*, float*, int, int, cudaStream_t);

template <typename T>
int slice_xattn_bwd(const float* Q, const float* KV, const T* dO, float* dQ, int accumulate, float* P, float* dS,
                    float* dKV, int B, cudaStream_t st, DropSpec dr) {
  const long total = static_cast<long>(B) * kNPatch;
  if (total <= 0) return HS_OK;
  slice_xattn_bwd_rows_kernel<T><<<static_cast<unsigned>((total + 7) / 8), 256, 0, st>>>(Q, KV, dO, dQ, accumulate, P, dS,
                                                                                       total, dr);
  count_launch();
  slice_xattn_bwd_kv_kernel<T><<<dim3(2 * kHidden / 256, B), 256, 0, st>>>(Q, dO, P, dS, dKV);
  count_launch();
  return launch_status();
}
template int slice_xattn_bwd<float>(const float*, const float*, const float*, float*, int, float*, float*, float*, int,
                                    cudaStream_t, DropSpec);
template int slice_xattn_bwd<__nv_bfloat16>(const float*, const float*, const __nv_bfloat16*, float*, int, float*, float*,
                                            float*, int, cudaStream_t, DropSpec);

int score_scale_bwd(const float* dX, const float* XP, const float* Z, const float* g, const float* be, const float* ws,
                    const float* scores, float* dXP, float* dZ, float* dg_partial, float* db_partial,
                    float* dws_partial, float* dbs_partial, int B, cudaStream_t st) {
  const long total = static_cast<long>(B) * kNPatch;
  if (total <= 0) return HS_OK;
  score_scale_bwd_kernel<<<layernorm_bwd_blocks(total), 256, 0, st>>>(dX, XP, Z, g, be, ws, scores, dXP, dZ, dg_partial,
                                                                      db_partial, dws_partial, dbs_partial, total);
  count_launch();
  return launch_status();
}

int add_rows(float* out, const float* a, long n, cudaStream_t st) {
  if (n <= 0) return HS_OK;
  if (n % 4) return HS_ERR_SHAPE;
  add_rows_kernel<<<grid_1d(n / 4, 256), 256, 0, st>>>(out, a, n / 4);
  count_launch();
  return launch_status();
}

}  // namespace hs
