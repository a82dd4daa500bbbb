// Volume ingest: the arithmetic of the reference's offline preprocessing
// Data/data_processing/CT-RATE/CT-RATE_nii_to_3D_volume_npy_file.py:25-110 as four HBM-bound kernels with no host
// synchronisation between them (min / max and the foreground box stay on the device):
//   hu_resample            slope*x + intercept, clamp to the HU window, (2,0,1) transpose and trilinear resample to the
//                          target spacing (F.interpolate(mode='trilinear', align_corners=False), lines 25-38, 73-91)
//   minmax                 exact global min / max (lines 103-104)
//   foreground_bbox        MONAI CropForeground with the default select_fn (x > 0) on the min-max normalised volume,
//                          i.e. x > min; box = [first, last + 1) foreground index per axis (line 116)
//   crop_normalize_resize  (x - min) / max(max - min, 1e-8) on the cropped box, trilinear resize to [32,256,256]
//                          (MONAI Resize(mode='bilinear') on a 3-D volume, align_corners=False; lines 105-106, 117)
// Interpolation follows PyTorch's area_pixel_compute_source_index / upsample_trilinear3d exactly: src = scale*(dst+0.5)
// - 0.5 clamped at 0, scale = in/out in fp32, i1 = i0 + (i0 < in-1), nested weights t(h(w)).
#include "common.cuh"
#include "kernels.h"

#include <cfloat>

namespace hs {

extern void count_launch();

namespace {

inline int launch_ok() { return cudaGetLastError() == cudaSuccess ? HS_OK : HS_ERR_CUDA; }

struct Axis {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Axis src_index(int dst, int in_size, float scale) {
  float s = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  Axis a;
  a.i0 = static_cast<int>(s);
  if (a.i0 > in_size - 1) a.i0 = in_size - 1;
  a.i1 = a.i0 + (a.i0 < in_size - 1 ? 1 : 0);
  a.l1 = s - static_cast<float>(a.i0);
  a.l0 = 1.f - a.l1;
  return a;
}
template <typename F>
__device__ __forceinline__ float trilinear(const Axis& t, const Axis& h, const Axis& w, F&& at) {
  return t.l0 * (h.l0 * (w.l0 * at(t.i0, h.i0, w.i0) + w.l1 * at(t.i0, h.i0, w.i1)) +
                 h.l1 * (w.l0 * at(t.i0, h.i1, w.i0) + w.l1 * at(t.i0, h.i1, w.i1))) +
         t.l1 * (h.l0 * (w.l0 * at(t.i1, h.i0, w.i0) + w.l1 * at(t.i1, h.i0, w.i1)) +
                 h.l1 * (w.l0 * at(t.i1, h.i1, w.i0) + w.l1 * at(t.i1, h.i1, w.i1)));
}

// raw is in NIfTI array order [n0][n1][n2] with the slice index n2 contiguous; the resampled volume is [z][a][b] with b
// contiguous.  A block produces an output tile of TA x TB x TZ voxels: it first stages the raw sub-block those voxels
// interpolate from in shared memory (windowed on the way in) with z-contiguous, coalesced loads, then writes the outputs
// b-contiguously.  (The direct gather version ran at 0.49 TB/s: adjacent threads read the raw volume n2*4 bytes apart;
// this tiling measures 0.79 TB/s at 512x512x303 -- a 2x32x32 tile with warp-per-row staging was slower, 0.52 TB/s.)
constexpr int TA = 4, TB = 32, TZ = 16;
constexpr int kRawTileFloats = 12288;        // 48 KB: upper bound of the staged sub-block, checked on the host

__global__ void __launch_bounds__(256) hu_resample_kernel(const float* __restrict__ raw, int n0, int n1, int n2,
                                                          float slope, float intercept, float lo, float hi,
                                                          float* __restrict__ out, int o0, int o1, int o2, float s0,
                                                          float s1, float s2) {
  extern __shared__ float tile[];
  const int b0 = blockIdx.x * TB, a0 = blockIdx.y * TA, z0 = blockIdx.z * TZ;
  const int b_end = min(b0 + TB, o2) - 1, a_end = min(a0 + TA, o1) - 1, z_end = min(z0 + TZ, o0) - 1;
  // raw index ranges touched by the tile (source indices are monotonic in the destination index)
  const int rz0 = src_index(z0, n2, s0).i0, rz1 = src_index(z_end, n2, s0).i1;
  const int ra0 = src_index(a0, n0, s1).i0, ra1 = src_index(a_end, n0, s1).i1;
  const int rb0 = src_index(b0, n1, s2).i0, rb1 = src_index(b_end, n1, s2).i1;
  const int nz = rz1 - rz0 + 1, na = ra1 - ra0 + 1, nb = rb1 - rb0 + 1;
  const int count = na * nb * nz;
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    const int zz = i % nz, bb = (i / nz) % nb, aa = i / (nz * nb);
    const float v = slope * __ldg(raw + (static_cast<long>(ra0 + aa) * n1 + (rb0 + bb)) * n2 + (rz0 + zz)) + intercept;
    tile[i] = fminf(fmaxf(v, lo), hi);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TA * TB * TZ; i += blockDim.x) {
    const int b = b0 + (i % TB), z = z0 + (i / TB) % TZ, a = a0 + i / (TB * TZ);
    if (b > b_end || z > z_end || a > a_end) continue;
    const Axis t = src_index(z, n2, s0), h = src_index(a, n0, s1), w = src_index(b, n1, s2);
    out[(static_cast<long>(z) * o1 + a) * o2 + b] = trilinear(t, h, w, [&](int zz, int aa, int bb) {
      return tile[((aa - ra0) * nb + (bb - rb0)) * nz + (zz - rz0)];
    });
  }
}

// Two-pass variant (used when the caller provides scratch of the raw size): windowed transpose [n0*n1][n2] -> [n2][n0*n1],
// then a resample whose eight taps are b-contiguous across adjacent threads.
// Transpose: a block takes 32 FULL raw rows (n2 contiguous floats each, read by one warp per row: the 32-column tiles of the
// first version started every 128-byte segment at an arbitrary 4-byte offset of the 1212-byte rows and pulled 1.78x the
// input from DRAM -- ncu dram__bytes_read 566 MB for 318 MB) through a shared tile whose row pitch is 1 mod 32 banks.
__global__ void __launch_bounds__(256) hu_transpose_rows_kernel(const float* __restrict__ raw, long rows, int cols,
                                                                int pitch, float slope, float intercept, float lo,
                                                                float hi, float* __restrict__ out) {
  extern __shared__ float trow[];                       // [32][pitch]
  const long r0 = static_cast<long>(blockIdx.x) * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int rr = warp + 8 * k;
    const long r = r0 + rr;
    if (r < rows) {
      const float* src = raw + r * cols;
      for (int c = lane; c < cols; c += 32) trow[rr * pitch + c] = fminf(fmaxf(slope * __ldg(src + c) + intercept, lo), hi);
    }
  }
  __syncthreads();
  const long r = r0 + lane;
  if (r < rows)
    for (int c = warp; c < cols; c += 8) out[static_cast<long>(c) * rows + r] = trow[lane * pitch + c];
}
// fallback for very long rows (tile would not fit in shared memory)
__global__ void __launch_bounds__(256) hu_transpose_kernel(const float* __restrict__ raw, long rows, int cols,
                                                           float slope, float intercept, float lo, float hi,
                                                           float* __restrict__ out) {
  __shared__ float t[32][33];
  const long r0 = static_cast<long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const long r = r0 + ty + k;
    const int c = c0 + tx;
    if (r < rows && c < cols) t[ty + k][tx] = fminf(fmaxf(slope * __ldg(raw + r * cols + c) + intercept, lo), hi);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int c = c0 + ty + k;
    const long r = r0 + tx;
    if (r < rows && c < cols) out[static_cast<long>(c) * rows + r] = t[tx][ty + k];
  }
}
// Resample of the transposed volume x [d0][d1][d2] -> out [o0][o1][o2]: one block per output row segment (z, a, 128 b's), so
// the z / a source indices and weights are block-uniform and the only per-thread index math is one axis.  (The first
// version decomposed a flat 64-bit index with three 64-bit divisions per voxel: ncu showed it issue-bound, 72 % issue
// slots at 19 % of the DRAM rate.)
constexpr int kResZ = 8;      // output slices per block: amortises block launch and the per-row index math
__global__ void __launch_bounds__(128) resample_kernel(const float* __restrict__ x, int d0, int d1, int d2,
                                                       float* __restrict__ out, int o0, int o1, int o2, float s0,
                                                       float s1, float s2) {
  const int a = blockIdx.y;
  const int b = blockIdx.x * 128 + threadIdx.x;
  if (b >= o2) return;
  const Axis h = src_index(a, d1, s1), w = src_index(b, d2, s2);
  const int z_end = min((blockIdx.z + 1) * kResZ, o0);
  for (int z = blockIdx.z * kResZ; z < z_end; ++z) {
    const Axis t = src_index(z, d0, s0);
    const float* p00 = x + (static_cast<long>(t.i0) * d1 + h.i0) * d2;
    const float* p01 = x + (static_cast<long>(t.i0) * d1 + h.i1) * d2;
    const float* p10 = x + (static_cast<long>(t.i1) * d1 + h.i0) * d2;
    const float* p11 = x + (static_cast<long>(t.i1) * d1 + h.i1) * d2;
    // same nesting as trilinear(): t(h(w))
    const float v = t.l0 * (h.l0 * (w.l0 * __ldg(p00 + w.i0) + w.l1 * __ldg(p00 + w.i1)) +
                            h.l1 * (w.l0 * __ldg(p01 + w.i0) + w.l1 * __ldg(p01 + w.i1))) +
                    t.l1 * (h.l0 * (w.l0 * __ldg(p10 + w.i0) + w.l1 * __ldg(p10 + w.i1)) +
                            h.l1 * (w.l0 * __ldg(p11 + w.i0) + w.l1 * __ldg(p11 + w.i1)));
    out[(static_cast<long>(z) * o1 + a) * o2 + b] = v;
  }
}

// order-preserving float <-> int so that integer atomics give an exact float min / max
__device__ __forceinline__ int f2ord(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void minmax_init_kernel(int* ord2) {
  ord2[0] = f2ord(FLT_MAX);
  ord2[1] = f2ord(-FLT_MAX);
}
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ x, long n, int* __restrict__ ord2) {
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float v = __ldg(x + i);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  mn = -warp_max(-mn);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&ord2[0], f2ord(mn));
    atomicMax(&ord2[1], f2ord(mx));
  }
}
__global__ void minmax_finish_kernel(const int* ord2, float* minmax2) {
  minmax2[0] = ord2f(ord2[0]);
  minmax2[1] = ord2f(ord2[1]);
}

__global__ void bbox_init_kernel(int* bbox6, int d0, int d1, int d2) {
  bbox6[0] = d0; bbox6[1] = d1; bbox6[2] = d2;
  bbox6[3] = 0; bbox6[4] = 0; bbox6[5] = 0;
}
// one warp per row of d2 contiguous voxels: the two outer indices are warp-uniform, divisions happen once per row
__global__ void __launch_bounds__(256) bbox_kernel(const float* __restrict__ x, int d0, int d1, int d2,
                                                   const float* __restrict__ minmax2, int* __restrict__ bbox6) {
  const float mn = minmax2[0];
  const int lane = threadIdx.x & 31;
  const long rows = static_cast<long>(d0) * d1;
  const long warp0 = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  int lo0 = d0, lo1 = d1, lo2 = d2, hi0 = 0, hi1 = 0, hi2 = 0;
  for (long r = warp0; r < rows; r += nwarps) {
    const float* row = x + r * d2;
    int rl = d2, rh = 0;
    for (int c = lane; c < d2; c += 32) {
      if (__ldg(row + c) > mn) {         // (x - min) / max(max - min, 1e-8) > 0
        rl = min(rl, c);
        rh = max(rh, c + 1);
      }
    }
    if (__any_sync(0xffffffffu, rh > 0)) {
      const int a = static_cast<int>(r / d1), b = static_cast<int>(r - static_cast<long>(a) * d1);
      lo0 = min(lo0, a); hi0 = max(hi0, a + 1);
      lo1 = min(lo1, b); hi1 = max(hi1, b + 1);
    }
    lo2 = min(lo2, rl); hi2 = max(hi2, rh);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo2 = min(lo2, __shfl_xor_sync(0xffffffffu, lo2, o));
    hi2 = max(hi2, __shfl_xor_sync(0xffffffffu, hi2, o));
  }
  if (lane == 0 && hi0 > 0) {
    atomicMin(&bbox6[0], lo0); atomicMin(&bbox6[1], lo1); atomicMin(&bbox6[2], lo2);
    atomicMax(&bbox6[3], hi0); atomicMax(&bbox6[4], hi1); atomicMax(&bbox6[5], hi2);
  }
}
// no foreground at all (constant volume): MONAI keeps the volume -> use the full extent
__global__ void bbox_finish_kernel(int* bbox6, int d0, int d1, int d2) {
  if (bbox6[3] <= bbox6[0] || bbox6[4] <= bbox6[1] || bbox6[5] <= bbox6[2]) {
    bbox6[0] = 0; bbox6[1] = 0; bbox6[2] = 0;
    bbox6[3] = d0; bbox6[4] = d1; bbox6[5] = d2;
  }
}

__global__ void __launch_bounds__(256) crop_normalize_resize_kernel(const float* __restrict__ x, int d1, int d2,
                                                                    const float* __restrict__ minmax2,
                                                                    const int* __restrict__ bbox6,
                                                                    float* __restrict__ out, int o0, int o1, int o2) {
  const float mn = minmax2[0];
  const float range = fmaxf(minmax2[1] - mn, 1e-8f);
  const int l0 = bbox6[0], l1 = bbox6[1], l2 = bbox6[2];
  const int c0 = bbox6[3] - l0, c1 = bbox6[4] - l1, c2 = bbox6[5] - l2;       // cropped extent
  const float s0 = static_cast<float>(c0) / o0, s1 = static_cast<float>(c1) / o1, s2 = static_cast<float>(c2) / o2;
  const long total = static_cast<long>(o0) * o1 * o2;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % o2);
    const int b = static_cast<int>((i / o2) % o1);
    const int a = static_cast<int>(i / (static_cast<long>(o2) * o1));
    const Axis t = src_index(a, c0, s0), h = src_index(b, c1, s1), w = src_index(c, c2, s2);
    out[i] = trilinear(t, h, w, [&](int aa, int bb, int cc) {
      const float v = __ldg(x + (static_cast<long>(l0 + aa) * d1 + (l1 + bb)) * d2 + (l2 + cc));
      return (v - mn) / range;
    });
  }
}

inline unsigned grid_for(long n) {
  long b = (n + 255) / 256;
  const long cap = 148L * 16;
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int hu_resample(const float* raw, int n0, int n1, int n2, float slope, float intercept, float hu_min, float hu_max,
                float* out, int o0, int o1, int o2, float* scratch, cudaStream_t st) {
  if (n0 <= 0 || n1 <= 0 || n2 <= 0 || o0 <= 0 || o1 <= 0 || o2 <= 0) return HS_ERR_SHAPE;
  const float s0 = static_cast<float>(n2) / o0, s1 = static_cast<float>(n0) / o1, s2 = static_cast<float>(n1) / o2;
  if (scratch != nullptr) {
    const long rows = static_cast<long>(n0) * n1;
    const int pitch = (n2 + 31) / 32 * 32 + 1;
    const size_t tsmem = static_cast<size_t>(32) * pitch * sizeof(float);
    if (tsmem <= 200 * 1024) {
      static unsigned char tattr[kMaxDevices] = {0};
      if (first_use_on_device(tattr) &&
          cudaFuncSetAttribute(hu_transpose_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
              cudaSuccess)
        return HS_ERR_CUDA;
      hu_transpose_rows_kernel<<<static_cast<unsigned>((rows + 31) / 32), 256, tsmem, st>>>(
          raw, rows, n2, pitch, slope, intercept, hu_min, hu_max, scratch);
    } else {
      const dim3 tg(static_cast<unsigned>((rows + 31) / 32), static_cast<unsigned>((n2 + 31) / 32));
      hu_transpose_kernel<<<tg, 256, 0, st>>>(raw, rows, n2, slope, intercept, hu_min, hu_max, scratch);
    }
    if (o1 > 65535) return HS_ERR_SHAPE;
    resample_kernel<<<dim3((o2 + 127) / 128, o1, (o0 + kResZ - 1) / kResZ), 128, 0, st>>>(scratch, n2, n0, n1, out, o0, o1, o2, s0, s1, s2);
    count_launch();
    count_launch();
    return launch_ok();
  }
  // staged sub-block per tile: (T*scale + 3) source indices per axis at most
  const long need = static_cast<long>(TA * s1 + 3) * static_cast<long>(TB * s2 + 3) * static_cast<long>(TZ * s0 + 3);
  if (need > kRawTileFloats) return HS_ERR_SHAPE;          // down-sampling by more than ~4x per axis: not an ingest case
  static unsigned char attr_set[kMaxDevices] = {0};
  if (first_use_on_device(attr_set)) {
    if (cudaFuncSetAttribute(hu_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kRawTileFloats * static_cast<int>(sizeof(float))) != cudaSuccess)
      return HS_ERR_CUDA;
  }
  const dim3 grid((o2 + TB - 1) / TB, (o1 + TA - 1) / TA, (o0 + TZ - 1) / TZ);
  hu_resample_kernel<<<grid, 256, kRawTileFloats * sizeof(float), st>>>(raw, n0, n1, n2, slope, intercept, hu_min,
                                                                       hu_max, out, o0, o1, o2, s0, s1, s2);
  count_launch();
  return launch_ok();
}

int minmax(const float* x, long n, float* minmax2, int* scratch2, cudaStream_t st) {
  if (n <= 0) return HS_ERR_SHAPE;
  minmax_init_kernel<<<1, 1, 0, st>>>(scratch2);
  minmax_kernel<<<grid_for(n), 256, 0, st>>>(x, n, scratch2);
  minmax_finish_kernel<<<1, 1, 0, st>>>(scratch2, minmax2);
  count_launch();
  return launch_ok();
}

int foreground_bbox(const float* x, int d0, int d1, int d2, const float* minmax2, int* bbox6, cudaStream_t st) {
  if (d0 <= 0 || d1 <= 0 || d2 <= 0) return HS_ERR_SHAPE;
  bbox_init_kernel<<<1, 1, 0, st>>>(bbox6, d0, d1, d2);
  bbox_kernel<<<grid_for(static_cast<long>(d0) * d1 * d2), 256, 0, st>>>(x, d0, d1, d2, minmax2, bbox6);
  bbox_finish_kernel<<<1, 1, 0, st>>>(bbox6, d0, d1, d2);
  count_launch();
  return launch_ok();
}

int crop_normalize_resize(const float* x, int d0, int d1, int d2, const float* minmax2, const int* bbox6, float* out,
                          int o0, int o1, int o2, cudaStream_t st) {
  if (d0 <= 0 || d1 <= 0 || d2 <= 0 || o0 <= 0 || o1 <= 0 || o2 <= 0) return HS_ERR_SHAPE;
  crop_normalize_resize_kernel<<<grid_for(static_cast<long>(o0) * o1 * o2), 256, 0, st>>>(x, d1, d2, minmax2, bbox6,
                                                                                         out, o0, o1, o2);
  count_launch();
  return launch_ok();
}

}  // namespace hs
