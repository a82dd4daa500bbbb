// HBM-bound kernels of the HSENet visual path: LayerNorm, patch im2col, packer pooling / window attention,
// 2E3 slice-guided attention + score gating, CLIP-head normalisation, 2D-slice extraction, integer maps.
// All use 128-bit vectorised, warp-coalesced global accesses and warp-shuffle reductions (one warp per 768-wide row).
#include "common.cuh"
#include "kernels.h"

namespace hs {

extern void count_launch();

namespace {

// ---------------------------------------------------------------------------------------------------------------
// Integer maps.  These two functions are the ONLY place the kernels derive gather addresses from, and they are also
// what hsenet_patch_gather_map / hsenet_packer_window_map dump for the bit-exact parity tests.
// ---------------------------------------------------------------------------------------------------------------
// MONAI perceptron rearrange 'b c (h p1) (w p2) (d p3) -> b (h w d) (p1 p2 p3 c)', p=(4,16,16)   (vit.py:437)
__host__ __device__ __forceinline__ int patch_voxel(int t, int f) {
  const int dz = t >> 8, hy = (t >> 4) & 15, wx = t & 15;
  const int p1 = f >> 8, p2 = (f >> 4) & 15, p3 = f & 15;
  return ((dz * 4 + p1) << 16) + ((hy * 16 + p2) << 8) + (wx * 16 + p3);
}
// window n = dz*16 + wy*4 + hx, member e = sw*4 + sh  ->  HR token   (spatial_pooling_projector.py:70-71, kernel (1,4,4))
__host__ __device__ __forceinline__ int window_member(int n, int e) {
  const int dz = n >> 4, wy = (n >> 2) & 3, hx = n & 3;
  const int sw = e >> 2, sh = e & 3;
  return dz * 256 + (4 * wy + sw) * 16 + (4 * hx + sh);
}

template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  static __device__ __forceinline__ float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <>
struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 load(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float4 v) {
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
  }
};
template <>
struct Vec4<__half> {
  static __device__ __forceinline__ float4 load(const __half* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
};

constexpr int kVecPerLane = kHidden / 128;   // 6 float4 per lane per 768-wide row

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm(768), eps 1e-5, fp32 statistics (two-pass in registers).  One warp per row.
// ---------------------------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long ldx,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, long rows,
                                                        OutT* __restrict__ out, long ldo,
                                                        OutT* __restrict__ out2, int seq, float eps) {
  pdl_prologue_done();
  const long row = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * ldx;
  float4 v[kVecPerLane];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    v[i] = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / kHidden);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / kHidden) + eps);
  OutT* o2 = nullptr;
  if (out2 != nullptr) {
    const long b = row / seq;
    const int t = static_cast<int>(row - b * seq);
    if (t >= 1) o2 = out2 + (b * (seq - 1) + (t - 1)) * kHidden;
  }
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + c));
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + bb.x;
    y.y = (v[i].y - mean) * rstd * g.y + bb.y;
    y.z = (v[i].z - mean) * rstd * g.z + bb.z;
    y.w = (v[i].w - mean) * rstd * g.w + bb.w;
    if (out != nullptr) Vec4<OutT>::store(out + row * ldo + c, y);
    if (o2 != nullptr) Vec4<OutT>::store(o2 + c, y);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Patch im2col: [B,1,32,256,256] fp32 -> [B*2048,1024].  One CTA per (b, dz, hy): reads 64 full 1 KB input rows
// (coalesced), scatters 32/64-byte pieces into the 16 token rows of that (dz,hy).
// ---------------------------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ vol, OutT* __restrict__ out) {
  const int blk = blockIdx.x;              // b*128 + dz*16 + hy
  const int b = blk >> 7;
  const int t0 = (blk & 127) << 4;         // first token of this (dz,hy) row of the grid: wx = 0
  const float* vb = vol + static_cast<long>(b) * (32 * 256 * 256);
  OutT* ob = out + (static_cast<long>(b) * kNPatch + t0) * kPatchDim;
  const int q = threadIdx.x & 63;          // float4 index inside the 256-float input row
  const int wx = q >> 2;
  const int p3 = (q & 3) << 2;
#pragma unroll 4
  for (int pr = threadIdx.x >> 6; pr < 64; pr += 4) {   // pr = p1*16 + p2
    const int f = (pr << 4) + p3;
    const float4 v = __ldg(reinterpret_cast<const float4*>(vb + patch_voxel(t0 + wx, f)));
    Vec4<OutT>::store(ob + static_cast<long>(wx) * kPatchDim + f, v);
  }
}

template <typename OutT>
__global__ void cast_kernel(const float* __restrict__ in, OutT* __restrict__ out, long n4) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    Vec4<OutT>::store(out + i * 4, *reinterpret_cast<const float4*>(in + i * 4));
  }
}

__global__ void cls_rows_kernel(float* __restrict__ X, const float* __restrict__ cls, int seq) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < kHidden / 4; c += blockDim.x)
    reinterpret_cast<float4*>(X + static_cast<long>(b) * seq * kHidden)[c] =
        __ldg(reinterpret_cast<const float4*>(cls) + c);
}

template <typename InT, typename OutT>
__global__ void __launch_bounds__(256) gather_rows_kernel(const InT* __restrict__ in, long batch_stride,
                                                          long row_stride, int rows, long total,
                                                          OutT* __restrict__ out) {
  const long r = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= total) return;
  const int lane = threadIdx.x & 31;
  const long b = r / rows;
  const long t = r - b * rows;
  const InT* src = in + b * batch_stride + t * row_stride;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const int c = (i * 32 + lane) * 4;
    Vec4<OutT>::store(out + r * kHidden + c, Vec4<InT>::load(src + c));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Packer: avg_pool3d (1,4,4) == mean over the 16 members of each window.  One warp per (b, window).
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) packer_pool_kernel(const T* __restrict__ hr, T* __restrict__ lr, long total) {
  const long w = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);   // b*128 + n
  if (w >= total) return;
  const int lane = threadIdx.x & 31;
  const long b = w >> 7;
  const int n = static_cast<int>(w & 127);
  float4 acc[kVecPerLane];
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int e = 0; e < 16; ++e) {
    const T* src = hr + (b * kNPatch + window_member(n, e)) * kHidden;
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
      const float4 v = Vec4<T>::load(src + (i * 32 + lane) * 4);
      acc[i].x += v.x; acc[i].y += v.y; acc[i].z += v.z; acc[i].w += v.w;
    }
  }
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const float4 y = make_float4(acc[i].x * 0.0625f, acc[i].y * 0.0625f, acc[i].z * 0.0625f, acc[i].w * 0.0625f);
    Vec4<T>::store(lr + w * kHidden + (i * 32 + lane) * 4, y);
  }
}

// Per-window single-head attention: 1 query x 16 keys, d_k = 768 (spatial_pooling_projector.py:8-16, 76).
template <typename T>
__global__ void __launch_bounds__(256) packer_window_attn_kernel(const float* __restrict__ Q,
                                                                 const T* __restrict__ KV, T* __restrict__ O,
                                                                 long total, DropSpec dr) {
  const long w = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (w >= total) return;
  const int lane = threadIdx.x & 31;
  const long b = w >> 7;
  const int n = static_cast<int>(w & 127);
  float4 q[kVecPerLane];
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) q[i] = *reinterpret_cast<const float4*>(Q + w * kHidden + (i * 32 + lane) * 4);
  float sc[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const T* kr = KV + (b * kNPatch + window_member(n, e)) * (2 * kHidden);
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
      const float4 k = Vec4<T>::load(kr + (i * 32 + lane) * 4);
      d += (q[i].x * k.x + q[i].y * k.y) + (q[i].z * k.z + q[i].w * k.w);
    }
    sc[e] = warp_sum(d) * 0.036084391824351615f;   // 1/sqrt(768)
  }
  float m = sc[0];
#pragma unroll
  for (int e = 1; e < 16; ++e) m = fmaxf(m, sc[e]);
  float den = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    sc[e] = expf(sc[e] - m);
    den += sc[e];
  }
  const float inv = 1.0f / den;
  float4 acc[kVecPerLane];
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const T* vr = KV + (b * kNPatch + window_member(n, e)) * (2 * kHidden) + kHidden;
    float p = sc[e] * inv;
    if (dr.on()) p *= drop_mask(dr.seed, dr.keep_thresh, dr.scale, static_cast<unsigned long long>(w) * 16 + e);
#pragma unroll
    for (int i = 0; i < kVecPerLane; ++i) {
      const float4 v = Vec4<T>::load(vr + (i * 32 + lane) * 4);
      acc[i].x += p * v.x; acc[i].y += p * v.y; acc[i].z += p * v.z; acc[i].w += p * v.w;
    }
  }
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) Vec4<T>::store(O + w * kHidden + (i * 32 + lane) * 4, acc[i]);
}

// ---------------------------------------------------------------------------------------------------------------
// 2E3 slice-guided cross attention core (vit.py:25-33 called at :59): 2048 queries x 32 slice keys per volume,
// single head, d_k = 768.  K and V of one volume are staged in shared memory (fp32, 192 KB); each warp processes
// 4 query rows at a time so every shared-memory operand is reused 4x.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kXaRowsPerCta = 128;
constexpr int kXaR = 4;
template <typename OutT>
__global__ void __launch_bounds__(256, 1) slice_xattn_kernel(const float* __restrict__ Q, const float* __restrict__ KV,
                                                             OutT* __restrict__ O, float* __restrict__ attn, DropSpec dr) {
  extern __shared__ float4 xa_smem[];
  float* sK = reinterpret_cast<float*>(xa_smem);       // [32][768]
  float* sV = sK + kNSlice * kHidden;                  // [32][768]
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kNSlice * kHidden / 4; i += blockDim.x) {
    const int j = i / (kHidden / 4), c = (i % (kHidden / 4)) * 4;
    const float* src = KV + (static_cast<long>(b) * kNSlice + j) * (2 * kHidden) + c;
    *reinterpret_cast<float4*>(sK + j * kHidden + c) = __ldg(reinterpret_cast<const float4*>(src));
    *reinterpret_cast<float4*>(sV + j * kHidden + c) = __ldg(reinterpret_cast<const float4*>(src + kHidden));
  }
  __syncthreads();
  const long row_base = static_cast<long>(b) * kNPatch + blockIdx.x * kXaRowsPerCta;
  for (int g = warp; g < kXaRowsPerCta / kXaR; g += 8) {
    const long r0 = row_base + g * kXaR;
    float4 q[kXaR][kVecPerLane];
#pragma unroll
    for (int r = 0; r < kXaR; ++r)
#pragma unroll
      for (int i = 0; i < kVecPerLane; ++i)
        q[r][i] = *reinterpret_cast<const float4*>(Q + (r0 + r) * kHidden + (i * 32 + lane) * 4);
    // lane j ends up holding the score of key j for each of the kXaR rows
    float sc[kXaR];
#pragma unroll
    for (int r = 0; r < kXaR; ++r) sc[r] = 0.f;
    for (int j = 0; j < kNSlice; ++j) {
      float d[kXaR];
#pragma unroll
      for (int r = 0; r < kXaR; ++r) d[r] = 0.f;
#pragma unroll
      for (int i = 0; i < kVecPerLane; ++i) {
        const float4 k = *reinterpret_cast<const float4*>(sK + j * kHidden + (i * 32 + lane) * 4);
#pragma unroll
        for (int r = 0; r < kXaR; ++r)
          d[r] += (q[r][i].x * k.x + q[r][i].y * k.y) + (q[r][i].z * k.z + q[r][i].w * k.w);
      }
#pragma unroll
      for (int r = 0; r < kXaR; ++r) {
        const float t = warp_sum(d[r]);
        if (lane == j) sc[r] = t * 0.036084391824351615f;   // 1/sqrt(768)
      }
    }
    float p[kXaR];
#pragma unroll
    for (int r = 0; r < kXaR; ++r) {
      const float m = warp_max(sc[r]);
      const float e = expf(sc[r] - m);
      p[r] = e / warp_sum(e);
      if (dr.on())      // train mode: p_attn = dropout(p_attn) (vit.py:31-32); the returned attention is the dropped one
        p[r] *= drop_mask(dr.seed, dr.keep_thresh, dr.scale, static_cast<unsigned long long>(r0 + r) * kNSlice + lane);
      if (attn != nullptr) attn[(r0 + r) * kNSlice + lane] = p[r];
    }
    float4 acc[kXaR][kVecPerLane];
#pragma unroll
    for (int r = 0; r < kXaR; ++r)
#pragma unroll
      for (int i = 0; i < kVecPerLane; ++i) acc[r][i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < kNSlice; ++j) {
      float pj[kXaR];
#pragma unroll
      for (int r = 0; r < kXaR; ++r) pj[r] = __shfl_sync(0xffffffffu, p[r], j);
#pragma unroll
      for (int i = 0; i < kVecPerLane; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(sV + j * kHidden + (i * 32 + lane) * 4);
#pragma unroll
        for (int r = 0; r < kXaR; ++r) {
          acc[r][i].x += pj[r] * v.x; acc[r][i].y += pj[r] * v.y;
          acc[r][i].z += pj[r] * v.z; acc[r][i].w += pj[r] * v.w;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kXaR; ++r)
#pragma unroll
      for (int i = 0; i < kVecPerLane; ++i)
        Vec4<OutT>::store(O + (r0 + r) * kHidden + (i * 32 + lane) * 4, acc[r][i]);
  }
}

// LN(Z) . w_s + b_s -> sigmoid -> X[b, 1+t, :] = XP[b, t, :] * score     (vit.py:62, 338-345).  One warp per row.
__global__ void __launch_bounds__(256) score_scale_kernel(const float* __restrict__ Z, const float* __restrict__ g,
                                                          const float* __restrict__ be,
                                                          const float* __restrict__ ws, const float* __restrict__ bs,
                                                          const float* __restrict__ XP, float* __restrict__ X,
                                                          float* __restrict__ scores, long total) {
  const long row = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);   // b*2048 + t
  if (row >= total) return;
  const int lane = threadIdx.x & 31;
  float4 v[kVecPerLane];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    v[i] = *reinterpret_cast<const float4*>(Z + row * kHidden + (i * 32 + lane) * 4);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / kHidden);
  float qv = 0.f;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    qv += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(qv) * (1.0f / kHidden) + kLnEps);
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(be + c));
    const float4 w = __ldg(reinterpret_cast<const float4*>(ws + c));
    dot += ((v[i].x - mean) * rstd * gg.x + bb.x) * w.x + ((v[i].y - mean) * rstd * gg.y + bb.y) * w.y +
           ((v[i].z - mean) * rstd * gg.z + bb.z) * w.z + ((v[i].w - mean) * rstd * gg.w + bb.w) * w.w;
  }
  dot = warp_sum(dot) + __ldg(bs);
  const float score = 1.0f / (1.0f + expf(-dot));
  if (scores != nullptr && lane == 0) scores[row] = score;
  const long b = row >> 11;
  const long t = row & 2047;
  float* xo = X + (b * kSeq + 1 + t) * kHidden;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 x = *reinterpret_cast<const float4*>(XP + row * kHidden + c);
    *reinterpret_cast<float4*>(xo + c) = make_float4(x.x * score, x.y * score, x.z * score, x.w * score);
  }
}

// F.normalize(dim=-1): x / max(||x||_2, 1e-12).  One warp per row, dim multiple of 128.
__global__ void __launch_bounds__(256) l2norm_kernel(const float* __restrict__ in, float* __restrict__ out, int rows,
                                                     int dim) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int c = lane * 4; c < dim; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(in + static_cast<long>(row) * dim + c);
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  const float inv = 1.0f / fmaxf(sqrtf(warp_sum(s)), 1e-12f);
  for (int c = lane * 4; c < dim; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(in + static_cast<long>(row) * dim + c);
    *reinterpret_cast<float4*>(out + static_cast<long>(row) * dim + c) =
        make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 2D-slice extraction (vit.py:529-531): trilinear (32,256,256) -> (32,oh,ow), align_corners=False.  Depth is
// 32 -> 32 so the depth weight is exactly (1,0) and the op is a per-slice bilinear resize; the result is written
// to the 3 broadcast channels of [B*32,3,oh,ow].  Each thread produces 4 consecutive x for one output row.
// ---------------------------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void __launch_bounds__(256) slice_extract_kernel(const float* __restrict__ vol, OutT* __restrict__ out,
                                                            int oh, int ow, float sy, float sx) {
  const int slice = blockIdx.y;                        // b*32 + z
  const int y = blockIdx.x * 4 + (threadIdx.x >> 6);
  const int x0 = (threadIdx.x & 63) * 4;
  if (y >= oh || x0 >= ow) return;
  const float* src = vol + static_cast<long>(slice) * (256 * 256);
  float fy = sy * (y + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy;
  const int y0 = static_cast<int>(fy);
  const int y1 = y0 + (y0 < 255 ? 1 : 0);
  const float ly = fy - y0, hy = 1.f - ly;
  float r[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = x0 + k;
    float fx = sx * (x + 0.5f) - 0.5f;
    fx = fx < 0.f ? 0.f : fx;
    const int xa = static_cast<int>(fx);
    const int xb = xa + (xa < 255 ? 1 : 0);
    const float lx = fx - xa, hx = 1.f - lx;
    r[k] = hy * (hx * __ldg(src + y0 * 256 + xa) + lx * __ldg(src + y0 * 256 + xb)) +
           ly * (hx * __ldg(src + y1 * 256 + xa) + lx * __ldg(src + y1 * 256 + xb));
  }
  const long plane = static_cast<long>(oh) * ow;
  OutT* o = out + static_cast<long>(slice) * 3 * plane + static_cast<long>(y) * ow + x0;
  if (x0 + 3 < ow) {
    const float4 v = make_float4(r[0], r[1], r[2], r[3]);
    Vec4<OutT>::store(o, v);
    Vec4<OutT>::store(o + plane, v);
    Vec4<OutT>::store(o + 2 * plane, v);
  } else {
    for (int k = 0; k < 4 && x0 + k < ow; ++k) {
      o[k] = static_cast<OutT>(r[k]);
      o[plane + k] = static_cast<OutT>(r[k]);
      o[2 * plane + k] = static_cast<OutT>(r[k]);
    }
  }
}

// 256 -> 224 (the reference's size, vit.py:529) is exactly 8 : 7.  With src = (8/7)(x + 0.5) - 0.5 the outputs 7k .. 7k+6 of a
// row interpolate between inputs 8k+j and 8k+j+1 (the fractional parts 0.07 .. 0.93 never come near an integer, so float
// rounding cannot move an index), i.e. each group of 7 outputs reads ONE aligned group of 8 inputs per source row:
// a lane loads 2 x 2 float4 (a warp = the two full 1 KB source rows, perfectly coalesced) and produces 7 outputs with
// compile-time register indices; the output row is staged in shared memory and stored as 128-bit words to the three
// channel planes.  The generic kernel above issued 16 scalar loads per 4 outputs and was issue-bound (ncu: 67 % issue
// slots, 44 % of the DRAM rate).  Weights are computed per x with the same fp32 expression, so results are bit-identical.
template <typename OutT>
__global__ void __launch_bounds__(256) slice_extract_224_kernel(const float* __restrict__ vol, OutT* __restrict__ out) {
  __shared__ __align__(16) OutT stage[8][224];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.y;
  const int y = blockIdx.x * 8 + warp;                 // 28 blocks x 8 rows = 224
  const float* src = vol + static_cast<long>(slice) * (256 * 256);
  const float sc = 256.0f / 224.0f;
  float fy = sc * (y + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy;
  const int y0 = static_cast<int>(fy);
  const int y1 = y0 + (y0 < 255 ? 1 : 0);
  const float ly = fy - y0, hy = 1.f - ly;
  float a[8], b[8];
  {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(src + y0 * 256 + lane * 8));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(src + y0 * 256 + lane * 8 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(src + y1 * 256 + lane * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(src + y1 * 256 + lane * 8 + 4));
    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
  }
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const int x = lane * 7 + j;
    float fx = sc * (x + 0.5f) - 0.5f;
    fx = fx < 0.f ? 0.f : fx;
    const float lx = fx - static_cast<float>(lane * 8 + j), hx = 1.f - lx;      // xa == 8 * lane + j
    const float r = hy * (hx * a[j] + lx * a[j + 1]) + ly * (hx * b[j] + lx * b[j + 1]);
    stage[warp][x] = static_cast<OutT>(r);
  }
  __syncwarp();
  constexpr int kVecs = 224 * static_cast<int>(sizeof(OutT)) / 16;      // 28 (bf16) or 56 (fp32) 16-byte words per row
  const long plane = 224L * 224;
  OutT* o = out + static_cast<long>(slice) * 3 * plane + static_cast<long>(y) * 224;
  for (int v = lane; v < kVecs; v += 32) {
    const uint4 w = reinterpret_cast<const uint4*>(&stage[warp][0])[v];
    reinterpret_cast<uint4*>(o)[v] = w;
    reinterpret_cast<uint4*>(o + plane)[v] = w;
    reinterpret_cast<uint4*>(o + 2 * plane)[v] = w;
  }
}

// Slice branch (row f-2): bilinear (32,256,256) -> (32,224,224) resize of the volume (the depth weight is exactly (1,0),
// vit.py:805) written directly as the 16x16 patch matrix of a ViT-B/16 stem: out[(slice*196 + py*14 + px), iy*16 + ix].
// The reference expands the slice to 3 identical channels before the stem conv; the channel sum is folded into the
// weight instead, so one channel is enough.  One thread per 4 consecutive x.
template <typename OutT>
__global__ void __launch_bounds__(256) slice_patches_kernel(const float* __restrict__ vol, OutT* __restrict__ out) {
  const int slice = blockIdx.y;
  const int y = blockIdx.x * 4 + (threadIdx.x >> 6);       // 56 blocks x 4 rows
  const int x0 = (threadIdx.x & 63) * 4;
  if (x0 >= 224) return;
  const float* src = vol + static_cast<long>(slice) * (256 * 256);
  const float sc = 256.0f / 224.0f;
  float fy = sc * (y + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy;
  const int y0 = static_cast<int>(fy);
  const int y1 = y0 + (y0 < 255 ? 1 : 0);
  const float ly = fy - y0, hy = 1.f - ly;
  float r[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float fx = sc * (x0 + k + 0.5f) - 0.5f;
    fx = fx < 0.f ? 0.f : fx;
    const int xa = static_cast<int>(fx);
    const int xb = xa + (xa < 255 ? 1 : 0);
    const float lx = fx - xa, hx = 1.f - lx;
    r[k] = hy * (hx * __ldg(src + y0 * 256 + xa) + lx * __ldg(src + y0 * 256 + xb)) +
           ly * (hx * __ldg(src + y1 * 256 + xa) + lx * __ldg(src + y1 * 256 + xb));
  }
  const int py = y >> 4, iy = y & 15, px = x0 >> 4, ix = x0 & 15;
  OutT* o = out + (static_cast<long>(slice) * 196 + py * 14 + px) * 256 + iy * 16 + ix;
  Vec4<OutT>::store(o, make_float4(r[0], r[1], r[2], r[3]));
}

__global__ void patch_map_kernel(int32_t* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kNPatch * kPatchDim) out[i] = patch_voxel(i / kPatchDim, i % kPatchDim);
}
__global__ void window_map_kernel(int32_t* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 128 * 16) out[i] = window_member(i / 16, i % 16);
}

inline int launch_status() { return cudaGetLastError() == cudaSuccess ? HS_OK : HS_ERR_CUDA; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------
template <typename OutT>
int layernorm_rows(const float* x, long ldx, const float* gamma, const float* beta, long rows, OutT* out, long ldo,
                   OutT* out2, int seq, cudaStream_t stream, float eps) {
  if (rows <= 0) return HS_OK;
  if (!aligned16(x) || (ldx % 4) || (ldo % 4)) return HS_ERR_ALIGN;
  ProfScope prof(PROF_LAYERNORM, 0.0, double(rows) * kHidden * (4.0 + sizeof(OutT) * ((out != nullptr) + (out2 != nullptr))), stream);
  launch_pdl(layernorm_kernel<OutT>, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, stream, x, ldx, gamma,
             beta, rows, out, ldo, out2, seq, eps);
  count_launch();
  return launch_status();
}
template int layernorm_rows<float>(const float*, long, const float*, const float*, long, float*, long, float*, int,
                                   cudaStream_t, float);
template int layernorm_rows<__nv_bfloat16>(const float*, long, const float*, const float*, long, __nv_bfloat16*, long,
                                           __nv_bfloat16*, int, cudaStream_t, float);

template <typename OutT>
int im2col_patches(const float* vol, int B, OutT* out, cudaStream_t stream) {
  if (B <= 0) return HS_OK;
  if (!aligned16(vol) || !aligned16(out)) return HS_ERR_ALIGN;
  ProfScope prof(PROF_IM2COL, 0.0, double(B) * kNPatch * kPatchDim * (4.0 + sizeof(OutT)), stream);
  im2col_kernel<OutT><<<B * 128, 256, 0, stream>>>(vol, out);
  count_launch();
  return launch_status();
}
template int im2col_patches<float>(const float*, int, float*, cudaStream_t);
template int im2col_patches<__nv_bfloat16>(const float*, int, __nv_bfloat16*, cudaStream_t);

template <typename OutT>
int cast_rows(const float* in, OutT* out, long n, cudaStream_t stream) {
  if (n <= 0) return HS_OK;
  if ((n % 4) || !aligned16(in) || (reinterpret_cast<uintptr_t>(out) & 7)) return HS_ERR_ALIGN;
  const long n4 = n / 4;
  long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cast_kernel<OutT><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(in, out, n4);
  count_launch();
  return launch_status();
}
template int cast_rows<float>(const float*, float*, long, cudaStream_t);
template int cast_rows<__nv_bfloat16>(const float*, __nv_bfloat16*, long, cudaStream_t);

// LayerNorm folded into the next linear layer (gemm_epilogue.cuh): one block per output feature n
//   w_folded[n,k] = bf16(gamma[k] * w[n,k]);  colsum[n] = sum_k float(w_folded[n,k]);  bias_folded[n] = bias[n] + w[n,:] . beta
namespace {
__global__ void fold_layernorm_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const float* __restrict__ bias,
                                      __nv_bfloat16* __restrict__ wf, float* __restrict__ colsum,
                                      float* __restrict__ bias_f, int K) {
  __shared__ float red[2][8];
  const long n = blockIdx.x;
  float s = 0.f, bb = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float wv = w[n * K + k];
    const __nv_bfloat16 f = __float2bfloat16(wv * gamma[k]);
    wf[n * K + k] = f;
    s += __bfloat162float(f);          // the sum of what the tensor cores will actually multiply by
    bb = fmaf(wv, beta[k], bb);
  }
  s = warp_sum(s);
  bb = warp_sum(bb);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = bb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tb = 0.f;
    for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) { ts += red[0][i]; tb += red[1][i]; }
    colsum[n] = ts;
    bias_f[n] = tb + (bias != nullptr ? bias[n] : 0.f);
  }
}
}  // namespace

int fold_layernorm(const float* w, const float* gamma, const float* beta, const float* bias, int N, int K,
                   __nv_bfloat16* w_folded, float* colsum, float* bias_folded, cudaStream_t stream) {
  if (N <= 0 || K <= 0) return HS_OK;
  fold_layernorm_kernel<<<static_cast<unsigned>(N), 256, 0, stream>>>(w, gamma, beta, bias, w_folded, colsum,
                                                                      bias_folded, K);
  count_launch();
  return launch_status();
}

int write_cls_rows(float* X, const float* cls, int B, int seq, cudaStream_t stream) {
  if (B <= 0) return HS_OK;
  cls_rows_kernel<<<B, 192, 0, stream>>>(X, cls, seq);
  count_launch();
  return launch_status();
}

template <typename OutT>
int gather_rows(const void* in, int in_dtype, long batch_stride, long row_stride, int B, int rows, OutT* out,
                cudaStream_t stream) {
  const long total = static_cast<long>(B) * rows;
  if (total <= 0) return HS_OK;
  if ((batch_stride % 4) || (row_stride % 4) || (reinterpret_cast<uintptr_t>(in) & 7)) return HS_ERR_ALIGN;
  if (in_dtype == HSENET_DTYPE_F32 && !aligned16(in)) return HS_ERR_ALIGN;
  const unsigned grid = static_cast<unsigned>((total + 7) / 8);
  switch (in_dtype) {
    case HSENET_DTYPE_F32:
      gather_rows_kernel<float, OutT><<<grid, 256, 0, stream>>>(static_cast<const float*>(in), batch_stride,
                                                                row_stride, rows, total, out);
      break;
    case HSENET_DTYPE_BF16:
      gather_rows_kernel<__nv_bfloat16, OutT><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in),
                                                                        batch_stride, row_stride, rows, total, out);
      break;
    case HSENET_DTYPE_F16:
      gather_rows_kernel<__half, OutT><<<grid, 256, 0, stream>>>(static_cast<const __half*>(in), batch_stride,
                                                                 row_stride, rows, total, out);
      break;
    default:
      return HS_ERR_ARG;
  }
  count_launch();
  return launch_status();
}
template int gather_rows<float>(const void*, int, long, long, int, int, float*, cudaStream_t);
template int gather_rows<__nv_bfloat16>(const void*, int, long, long, int, int, __nv_bfloat16*, cudaStream_t);

template <typename OutT>
int slice_cross_attention(const float* Q, const float* KV, OutT* O, float* attn, int B, cudaStream_t stream, DropSpec dr) {
  if (B <= 0) return HS_OK;
  constexpr int smem = 2 * kNSlice * kHidden * 4;   // 196,608 B
  static unsigned char attr_set[kMaxDevices] = {0};
  if (first_use_on_device(attr_set)) {
    if (cudaFuncSetAttribute(slice_xattn_kernel<OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
        cudaSuccess)
      return HS_ERR_CUDA;
  }
  ProfScope prof(PROF_SLICE_XATTN, 4.0 * B * kNPatch * double(kNSlice) * kHidden,
                 double(B) * (kNPatch * kHidden * (4.0 + sizeof(OutT)) + kNSlice * 2.0 * kHidden * 4.0), stream);
  slice_xattn_kernel<OutT><<<dim3(kNPatch / kXaRowsPerCta, B), 256, smem, stream>>>(Q, KV, O, attn, dr);
  count_launch();
  return launch_status();
}
template int slice_cross_attention<float>(const float*, const float*, float*, float*, int, cudaStream_t, DropSpec);
template int slice_cross_attention<__nv_bfloat16>(const float*, const float*, __nv_bfloat16*, float*, int,
                                                  cudaStream_t, DropSpec);

// ---- train-mode dropout helpers (kernels.h DropSpec) ----------------------------------------------------------------
DropSpec make_dropspec(float p, unsigned long long seed) {
  DropSpec d;
  if (p > 0.f) {
    const double keep = 1.0 - static_cast<double>(p);
    d.seed = seed;
    d.keep_thresh = keep <= 0.0 ? 0u : static_cast<unsigned int>(keep * 4294967296.0 > 4294967295.0 ? 4294967295.0 : keep * 4294967296.0);
    d.scale = keep <= 0.0 ? 1.0f : static_cast<float>(1.0 / keep);      // p = 1 drops everything (mask 0 either way)
  }
  return d;
}
namespace {
template <int MODE>   // 0: z = resid + m z   1: out = m in   2: out = m
__global__ void dropout_kernel(const float* __restrict__ in, const float* __restrict__ resid, float* __restrict__ out, long n,
                               DropSpec dr) {
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float m = drop_mask(dr.seed, dr.keep_thresh, dr.scale, static_cast<unsigned long long>(i));
    if (MODE == 0) out[i] = resid[i] + m * in[i];
    else if (MODE == 1) out[i] = m * in[i];
    else out[i] = m;
  }
}
unsigned drop_grid(long n) {
  const long blocks = (n + 255) / 256;
  return static_cast<unsigned>(blocks < 148L * 16 ? (blocks > 0 ? blocks : 1) : 148L * 16);
}
}  // namespace
int dropout_residual(float* z, const float* resid, long n, DropSpec dr, cudaStream_t st) {
  if (n <= 0) return HS_OK;
  dropout_kernel<0><<<drop_grid(n), 256, 0, st>>>(z, resid, z, n, dr);
  count_launch();
  return launch_status();
}
int dropout_scale(const float* in, float* out, long n, DropSpec dr, cudaStream_t st) {
  if (n <= 0) return HS_OK;
  dropout_kernel<1><<<drop_grid(n), 256, 0, st>>>(in, nullptr, out, n, dr);
  count_launch();
  return launch_status();
}
int dropout_mask(float* out, long n, DropSpec dr, cudaStream_t st) {
  if (n <= 0) return HS_OK;
  dropout_kernel<2><<<drop_grid(n), 256, 0, st>>>(nullptr, nullptr, out, n, dr);
  count_launch();
  return launch_status();
}

int score_and_scale(const float* Z, const float* ln_g, const float* ln_b, const float* w_s, const float* b_s,
                    const float* XP, float* X, float* scores, int B, cudaStream_t stream) {
  const long total = static_cast<long>(B) * kNPatch;
  if (total <= 0) return HS_OK;
  ProfScope prof(PROF_SCORE_SCALE, 0.0, double(total) * kHidden * 12.0, stream);
  score_scale_kernel<<<static_cast<unsigned>((total + 7) / 8), 256, 0, stream>>>(Z, ln_g, ln_b, w_s, b_s, XP, X,
                                                                                 scores, total);
  count_launch();
  return launch_status();
}

template <typename T>
int packer_pool(const T* HR, T* LR, int B, cudaStream_t stream) {
  const long total = static_cast<long>(B) * 128;
  if (total <= 0) return HS_OK;
  ProfScope prof(PROF_PACKER_POOL, 0.0, double(B) * (kNPatch + 128.0) * kHidden * sizeof(T), stream);
  packer_pool_kernel<T><<<static_cast<unsigned>((total + 7) / 8), 256, 0, stream>>>(HR, LR, total);
  count_launch();
  return launch_status();
}
template int packer_pool<float>(const float*, float*, int, cudaStream_t);
template int packer_pool<__nv_bfloat16>(const __nv_bfloat16*, __nv_bfloat16*, int, cudaStream_t);

template <typename T>
int packer_window_attention(const float* Q, const T* KV, T* O, int B, cudaStream_t stream, DropSpec dr) {
  const long total = static_cast<long>(B) * 128;
  if (total <= 0) return HS_OK;
  ProfScope prof(PROF_PACKER_WATTN, 0.0,
                 double(B) * (128.0 * kHidden * (4.0 + sizeof(T)) + kNPatch * 2.0 * kHidden * sizeof(T)), stream);
  packer_window_attn_kernel<T><<<static_cast<unsigned>((total + 7) / 8), 256, 0, stream>>>(Q, KV, O, total, dr);
  count_launch();
  return launch_status();
}
template int packer_window_attention<float>(const float*, const float*, float*, int, cudaStream_t, DropSpec);
template int packer_window_attention<__nv_bfloat16>(const float*, const __nv_bfloat16*, __nv_bfloat16*, int,
                                                    cudaStream_t, DropSpec);

int l2_normalize_rows(const float* in, float* out, int rows, int dim, cudaStream_t stream) {
  if (rows <= 0) return HS_OK;
  if (dim % 128) return HS_ERR_SHAPE;
  l2norm_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(in, out, rows, dim);
  count_launch();
  return launch_status();
}

template <typename OutT>
int slice_extract(const float* vol, OutT* out, int B, int oh, int ow, cudaStream_t stream) {
  if (B <= 0) return HS_OK;
  if (oh <= 0 || ow <= 0 || ow > 256 || oh > 65535 || (ow % 4)) return HS_ERR_SHAPE;
  const float sy = 256.0f / static_cast<float>(oh);
  const float sx = 256.0f / static_cast<float>(ow);
  ProfScope prof(PROF_SLICE_EXTRACT, 0.0, double(B) * 32.0 * (256.0 * 256.0 * 4.0 + 3.0 * oh * ow * sizeof(OutT)), stream);
  if (oh == 224 && ow == 224 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && aligned16(vol)) {
    slice_extract_224_kernel<OutT><<<dim3(28, B * 32), 256, 0, stream>>>(vol, out);
    count_launch();
    return launch_status();
  }
  slice_extract_kernel<OutT><<<dim3((oh + 3) / 4, B * 32), 256, 0, stream>>>(vol, out, oh, ow, sy, sx);
  count_launch();
  return launch_status();
}
template int slice_extract<float>(const float*, float*, int, int, int, cudaStream_t);
template int slice_extract<__nv_bfloat16>(const float*, __nv_bfloat16*, int, int, int, cudaStream_t);

template <typename OutT>
int slice_patches(const float* vol, OutT* out, int B, cudaStream_t stream) {
  if (B <= 0) return HS_OK;
  if (!aligned16(vol) || (reinterpret_cast<uintptr_t>(out) & 7)) return HS_ERR_ALIGN;
  slice_patches_kernel<OutT><<<dim3(56, B * 32), 256, 0, stream>>>(vol, out);
  count_launch();
  return launch_status();
}
template int slice_patches<float>(const float*, float*, int, cudaStream_t);
template int slice_patches<__nv_bfloat16>(const float*, __nv_bfloat16*, int, cudaStream_t);

int patch_gather_map(int32_t* out, cudaStream_t stream) {
  patch_map_kernel<<<kNPatch * kPatchDim / 256, 256, 0, stream>>>(out);
  count_launch();
  return launch_status();
}
int packer_window_map(int32_t* out, cudaStream_t stream) {
  window_map_kernel<<<8, 256, 0, stream>>>(out);
  count_launch();
  return launch_status();
}

}  // namespace hs
