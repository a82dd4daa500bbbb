// fp32 verification mode (north-star "fp32-accumulate verification mode", tolerance 1e-4 vs the oracle).
// Same pipeline and epilogue semantics as the tensor-core path but with fp32 operands and fp32 FMA accumulation on
// CUDA cores, so differences against the CPU oracle are pure summation-order noise (~1e-6).  Not a fallback: it is
// only selected explicitly with precision = HSENET_PREC_FP32_VERIFY and is still sm_100a device code.
#include "common.cuh"
#include "kernels.h"

namespace hs {

extern void count_launch();

namespace {

constexpr int TM = 128, TN = 128, TK = 16;

// C[M,N] = A[M,K] * W[N,K]^T ; 256 threads, 8x8 outputs per thread, smem tiles stored k-major.
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W,
                                                    int ldw, int M, int N, int K, const GemmEpilogue ep) {
  __shared__ float As[2][TK][TM + 4];
  __shared__ float Ws[2][TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int ty = tid >> 4, tx = tid & 15;          // 16 x 16 thread grid
  // global->smem mapping: 128 rows x 16 k = 512 float4, two per thread
  const int lr = tid >> 2;                         // 0..63  (+64)
  const int lk = (tid & 3) << 2;                   // 0,4,8,12
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  auto load_tile = [&](int kb, float4 (&ra)[2], float4 (&rw)[2]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lr + h * 64;
      const int gm = m0 + r;
      ra[h] = gm < M ? __ldg(reinterpret_cast<const float4*>(A + static_cast<long>(gm) * lda + kb * TK + lk))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
      rw[h] = __ldg(reinterpret_cast<const float4*>(W + static_cast<long>(n0 + r) * ldw + kb * TK + lk));
    }
  };
  auto store_tile = [&](int buf, const float4 (&ra)[2], const float4 (&rw)[2]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lr + h * 64;
      As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y; As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
      Ws[buf][lk + 0][r] = rw[h].x; Ws[buf][lk + 1][r] = rw[h].y; Ws[buf][lk + 2][r] = rw[h].z; Ws[buf][lk + 3][r] = rw[h].w;
    }
  };

  const int kblocks = K / TK;
  float4 ra[2], rw[2];
  load_tile(0, ra, rw);
  store_tile(0, ra, rw);
  __syncthreads();
  for (int kb = 0; kb < kblocks; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < kblocks) load_tile(kb + 1, ra, rw);
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[8], w[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[buf][k][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (kb + 1 < kblocks) {
      store_tile(buf ^ 1, ra, rw);
      __syncthreads();
    }
  }

  float* out_act = reinterpret_cast<float*>(ep.out_bf16);   // fp32 in verification mode
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= M) continue;
    long orow = row;
    int rr = 0;
    if (ep.rows_per_group > 0) {
      const int g = row / ep.rows_per_group;
      rr = row - g * ep.rows_per_group;
      orow = static_cast<long>(g) * ep.group_stride + ep.group_offset + rr;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = n0 + h * 64 + tx * 4;
      float4 x = make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
      if (ep.bias != nullptr) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
        x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
      }
      if (ep.rows_per_group > 0 && ep.row_add != nullptr) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(ep.row_add + static_cast<long>(rr) * N + col));
        x.x += a.x; x.y += a.y; x.z += a.z; x.w += a.w;
      }
      if (ep.resid != nullptr) {
        const float4 a = *reinterpret_cast<const float4*>(ep.resid + orow * ep.ld_resid + col);
        x.x += a.x; x.y += a.y; x.z += a.z; x.w += a.w;
      }
      if (ep.gelu) {
        x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w);
      }
      if (ep.dgelu_src != nullptr) {
        const float4 a = *reinterpret_cast<const float4*>(static_cast<const float*>(ep.dgelu_src) + orow * ep.ld_dgelu + col);
        x.x *= gelu_erf_grad(a.x); x.y *= gelu_erf_grad(a.y); x.z *= gelu_erf_grad(a.z); x.w *= gelu_erf_grad(a.w);
      }
      if (ep.out_f32 != nullptr) *reinterpret_cast<float4*>(ep.out_f32 + orow * ep.ld_f32 + col) = x;
      if (out_act != nullptr) *reinterpret_cast<float4*>(out_act + orow * ep.ld_bf16 + col) = x;
    }
  }
}

// fp32 flash attention on CUDA cores: one thread per query row, 32-key shared-memory tiles.
constexpr int AQ = 128, AK = 32;
__global__ void __launch_bounds__(AQ) attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                           float* __restrict__ lse_out, int S) {
  __shared__ float4 sK[AK][kHeadDim / 4];
  __shared__ float4 sV[AK][kHeadDim / 4];
  const int b = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * AQ + threadIdx.x;
  const int qrow = qi < S ? qi : S - 1;
  const long base = static_cast<long>(b) * S;
  const int ld = 3 * kHidden;
  float4 q[kHeadDim / 4], o[kHeadDim / 4];
  {
    const float4* qp = reinterpret_cast<const float4*>(qkv + (base + qrow) * ld + h * kHeadDim);
#pragma unroll
    for (int i = 0; i < kHeadDim / 4; ++i) {
      q[i] = __ldg(qp + i);
      o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < S; k0 += AK) {
    __syncthreads();
    for (int i = threadIdx.x; i < AK * (kHeadDim / 4); i += AQ) {
      const int j = i / (kHeadDim / 4), c = i % (kHeadDim / 4);
      const int kr = k0 + j < S ? k0 + j : S - 1;
      const float* src = qkv + (base + kr) * ld + h * kHeadDim;
      sK[j][c] = __ldg(reinterpret_cast<const float4*>(src + kHidden) + c);
      sV[j][c] = __ldg(reinterpret_cast<const float4*>(src + 2 * kHidden) + c);
    }
    __syncthreads();
    float s[AK];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < AK; ++j) {
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < kHeadDim / 4; ++i) {
        const float4 k = sK[j][i];
        d = fmaf(q[i].x, k.x, d); d = fmaf(q[i].y, k.y, d); d = fmaf(q[i].z, k.z, d); d = fmaf(q[i].w, k.w, d);
      }
      s[j] = (k0 + j < S) ? d * 0.125f : -INFINITY;
      tmax = fmaxf(tmax, s[j]);
    }
    const float mn = fmaxf(m, tmax);
    const float alpha = expf(m - mn);
    l *= alpha;
#pragma unroll
    for (int i = 0; i < kHeadDim / 4; ++i) {
      o[i].x *= alpha; o[i].y *= alpha; o[i].z *= alpha; o[i].w *= alpha;
    }
#pragma unroll
    for (int j = 0; j < AK; ++j) {
      const float p = expf(s[j] - mn);
      l += p;
#pragma unroll
      for (int i = 0; i < kHeadDim / 4; ++i) {
        const float4 v = sV[j][i];
        o[i].x = fmaf(p, v.x, o[i].x); o[i].y = fmaf(p, v.y, o[i].y);
        o[i].z = fmaf(p, v.z, o[i].z); o[i].w = fmaf(p, v.w, o[i].w);
      }
    }
    m = mn;
  }
  if (lse_out != nullptr) {      // log2-domain log-sum-exp of the scaled scores (same convention as the tcgen05 kernel)
    const int sp = ((S + 127) / 128) * 128;
    if (qi < sp) lse_out[(static_cast<long>(b) * kHeads + h) * sp + qi] = qi < S ? (m + logf(l)) * 1.4426950408889634f : INFINITY;
  }
  if (qi < S) {
    const float inv = 1.0f / l;
    float4* op = reinterpret_cast<float4*>(out + (base + qi) * kHidden + h * kHeadDim);
#pragma unroll
    for (int i = 0; i < kHeadDim / 4; ++i) op[i] = make_float4(o[i].x * inv, o[i].y * inv, o[i].z * inv, o[i].w * inv);
  }
}

// ---- fp32 attention backward on CUDA cores (verification mode): P is rebuilt from qkv and the saved log-sum-exp ----
//   P = 2^(s c - lse),  dP = dO V^T,  dS = P o (dP - D) / 8,  dQ = dS K,  dK = dS^T Q,  dV = P^T dO,  D = rowsum(dO o O)
constexpr float kAttC = 0.125f * 1.4426950408889634f;
// dQ: one thread per query row (q, dO row and the dQ accumulator in registers), 32-key shared-memory tiles
__global__ void __launch_bounds__(AQ) attention_bwd_dq_f32_kernel(const float* __restrict__ qkv,
                                                                  const float* __restrict__ d_out,
                                                                  const float* __restrict__ lse,
                                                                  const float* __restrict__ dvec,
                                                                  float* __restrict__ d_qkv, int S) {
  __shared__ float4 sK[AK][kHeadDim / 4];
  __shared__ float4 sV[AK][kHeadDim / 4];
  const int b = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * AQ + threadIdx.x;
  const int qrow = qi < S ? qi : S - 1;
  const long base = static_cast<long>(b) * S;
  const int ld = 3 * kHidden;
  const int sp = ((S + 127) / 128) * 128;
  float4 q[kHeadDim / 4], go[kHeadDim / 4], dq[kHeadDim / 4];
  {
    const float4* qp = reinterpret_cast<const float4*>(qkv + (base + qrow) * ld + h * kHeadDim);
    const float4* gp = reinterpret_cast<const float4*>(d_out + (base + qrow) * kHidden + h * kHeadDim);
#pragma unroll
    for (int i = 0; i < kHeadDim / 4; ++i) {
      q[i] = __ldg(qp + i);
      go[i] = __ldg(gp + i);
      dq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float l2 = lse[(static_cast<long>(b) * kHeads + h) * sp + qrow];
  const float dv = dvec[(static_cast<long>(b) * kHeads + h) * sp + qrow];
  for (int k0 = 0; k0 < S; k0 += AK) {
    __syncthreads();
    for (int i = threadIdx.x; i < AK * (kHeadDim / 4); i += AQ) {
      const int j = i / (kHeadDim / 4), c = i % (kHeadDim / 4);
      const int kr = k0 + j < S ? k0 + j : S - 1;
      const float* src = qkv + (base + kr) * ld + h * kHeadDim;
      sK[j][c] = __ldg(reinterpret_cast<const float4*>(src + kHidden) + c);
      sV[j][c] = __ldg(reinterpret_cast<const float4*>(src + 2 * kHidden) + c);
    }
    __syncthreads();
    for (int j = 0; j < AK && k0 + j < S; ++j) {
      float sd = 0.f, dp = 0.f;
#pragma unroll
      for (int i = 0; i < kHeadDim / 4; ++i) {
        const float4 k = sK[j][i], v = sV[j][i];
        sd = fmaf(q[i].x, k.x, sd); sd = fmaf(q[i].y, k.y, sd); sd = fmaf(q[i].z, k.z, sd); sd = fmaf(q[i].w, k.w, sd);
        dp = fmaf(go[i].x, v.x, dp); dp = fmaf(go[i].y, v.y, dp); dp = fmaf(go[i].z, v.z, dp); dp = fmaf(go[i].w, v.w, dp);
      }
      const float p = exp2f(sd * kAttC - l2);
      const float ds = p * (dp - dv) * 0.125f;
#pragma unroll
      for (int i = 0; i < kHeadDim / 4; ++i) {
        const float4 k = sK[j][i];
        dq[i].x = fmaf(ds, k.x, dq[i].x); dq[i].y = fmaf(ds, k.y, dq[i].y);
        dq[i].z = fmaf(ds, k.z, dq[i].z); dq[i].w = fmaf(ds, k.w, dq[i].w);
      }
    }
  }
  if (qi < S) {
    float4* op = reinterpret_cast<float4*>(d_qkv + (base + qi) * ld + h * kHeadDim);
#pragma unroll
    for (int i = 0; i < kHeadDim / 4; ++i) op[i] = dq[i];
  }
}

// dK (WHICH = 0) or dV (WHICH = 1): one thread per key row, 32-query shared-memory tiles of Q and dO
template <int WHICH>
__global__ void __launch_bounds__(AQ) attention_bwd_dkv_f32_kernel(const float* __restrict__ qkv,
                                                                   const float* __restrict__ d_out,
                                                                   const float* __restrict__ lse,
                                                                   const float* __restrict__ dvec,
                                                                   float* __restrict__ d_qkv, int S) {
  __shared__ float4 sQ[AK][kHeadDim / 4];
  __shared__ float4 sG[AK][kHeadDim / 4];
  __shared__ float sL[AK], sD[AK];
  const int b = blockIdx.z, h = blockIdx.y;
  const int ki = blockIdx.x * AQ + threadIdx.x;
  const int krow = ki < S ? ki : S - 1;
  const long base = static_cast<long>(b) * S;
  const int ld = 3 * kHidden;
  const int sp = ((S + 127) / 128) * 128;
  float4 k[kHeadDim / 4], v[kHeadDim / 4], acc[kHeadDim / 4];
  {
    const float* src = qkv + (base + krow) * ld + h * kHeadDim;
#pragma unroll
    for (int i = 0; i < kHeadDim / 4; ++i) {
      k[i] = __ldg(reinterpret_cast<const float4*>(src + kHidden) + i);
      v[i] = __ldg(reinterpret_cast<const float4*>(src + 2 * kHidden) + i);
      acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (int q0 = 0; q0 < S; q0 += AK) {
    __syncthreads();
    for (int i = threadIdx.x; i < AK * (kHeadDim / 4); i += AQ) {
      const int j = i / (kHeadDim / 4), c = i % (kHeadDim / 4);
      const int qr = q0 + j < S ? q0 + j : S - 1;
      sQ[j][c] = __ldg(reinterpret_cast<const float4*>(qkv + (base + qr) * ld + h * kHeadDim) + c);
      sG[j][c] = __ldg(reinterpret_cast<const float4*>(d_out + (base + qr) * kHidden + h * kHeadDim) + c);
    }
    if (threadIdx.x < AK) {
      const int qr = q0 + threadIdx.x < S ? q0 + threadIdx.x : S - 1;
      sL[threadIdx.x] = lse[(static_cast<long>(b) * kHeads + h) * sp + qr];
      sD[threadIdx.x] = dvec[(static_cast<long>(b) * kHeads + h) * sp + qr];
    }
    __syncthreads();
    for (int j = 0; j < AK && q0 + j < S; ++j) {
      float sd = 0.f, dp = 0.f;
#pragma unroll
      for (int i = 0; i < kHeadDim / 4; ++i) {
        const float4 q = sQ[j][i];
        sd = fmaf(q.x, k[i].x, sd); sd = fmaf(q.y, k[i].y, sd); sd = fmaf(q.z, k[i].z, sd); sd = fmaf(q.w, k[i].w, sd);
        if (WHICH == 0) {
          const float4 g = sG[j][i];
          dp = fmaf(g.x, v[i].x, dp); dp = fmaf(g.y, v[i].y, dp); dp = fmaf(g.z, v[i].z, dp); dp = fmaf(g.w, v[i].w, dp);
        }
      }
      const float p = exp2f(sd * kAttC - sL[j]);
      const float w = WHICH == 0 ? p * (dp - sD[j]) * 0.125f : p;
#pragma unroll
      for (int i = 0; i < kHeadDim / 4; ++i) {
        const float4 x = WHICH == 0 ? sQ[j][i] : sG[j][i];
        acc[i].x = fmaf(w, x.x, acc[i].x); acc[i].y = fmaf(w, x.y, acc[i].y);
        acc[i].z = fmaf(w, x.z, acc[i].z); acc[i].w = fmaf(w, x.w, acc[i].w);
      }
    }
  }
  if (ki < S) {
    float4* op = reinterpret_cast<float4*>(d_qkv + (base + ki) * ld + (WHICH == 0 ? kHidden : 2 * kHidden) + h * kHeadDim);
#pragma unroll
    for (int i = 0; i < kHeadDim / 4; ++i) op[i] = acc[i];
  }
}

}  // namespace

int attention_rowdot_f32(const float* out, const float* d_out, float* dvec, int B, int S, cudaStream_t stream);

int attention_bwd_f32(const float* qkv, const float* out, const float* d_out, const float* lse, float* dvec, float* d_qkv,
                      int B, int S, cudaStream_t stream) {
  if (B <= 0 || S <= 0) return HS_OK;
  const int rc = attention_rowdot_f32(out, d_out, dvec, B, S, stream);
  if (rc != HS_OK) return rc;
  const dim3 grid((S + AQ - 1) / AQ, kHeads, B);
  attention_bwd_dq_f32_kernel<<<grid, AQ, 0, stream>>>(qkv, d_out, lse, dvec, d_qkv, S);
  attention_bwd_dkv_f32_kernel<0><<<grid, AQ, 0, stream>>>(qkv, d_out, lse, dvec, d_qkv, S);
  attention_bwd_dkv_f32_kernel<1><<<grid, AQ, 0, stream>>>(qkv, d_out, lse, dvec, d_qkv, S);
  count_launch(); count_launch(); count_launch();
  return cudaGetLastError() == cudaSuccess ? HS_OK : HS_ERR_CUDA;
}

int gemm_f32(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
             cudaStream_t stream) {
  if (M <= 0) return HS_OK;
  if (N % TN != 0 || K % TK != 0) return HS_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15) || (lda % 4) || (ldw % 4))
    return HS_ERR_ALIGN;
  sgemm_kernel<<<dim3(N / TN, (M + TM - 1) / TM), 256, 0, stream>>>(A, lda, W, ldw, M, N, K, ep);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? HS_OK : HS_ERR_CUDA;
}

int attention_f32(const float* qkv, float* out, float* lse, int B, int S, cudaStream_t stream) {
  if (B <= 0 || S <= 0) return HS_OK;
  attention_f32_kernel<<<dim3((S + AQ - 1) / AQ, kHeads, B), AQ, 0, stream>>>(qkv, out, lse, S);
  count_launch();
  return cudaGetLastError() == cudaSuccess ? HS_OK : HS_ERR_CUDA;
}

}  // namespace hs
