// Helpers shared by the composite forward (api.cu) and training (train.cu) passes: workspace bump allocator,
// precision dispatch (bf16 tcgen05 kernels / fp32 CUDA-core verification kernels), status propagation.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace hs {
namespace {

constexpr size_t kAlign = 1024;
inline size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

struct Bump {
  uint8_t* base;
  size_t off = 0;
  explicit Bump(void* p) : base(static_cast<uint8_t*>(p)) {}
  template <typename T>
  T* take(size_t count) {
    T* p = reinterpret_cast<T*>(base + off);
    off += align_up(count * sizeof(T));
    return p;
  }
};

template <typename T>
struct Prec;
template <>
struct Prec<__nv_bfloat16> {
  static int gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
                  cudaStream_t s) {
    return gemm_bf16(A, lda, W, ldw, M, N, K, ep, s);
  }
  static int attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse, float* kmax, int B, int S,
                       cudaStream_t s) {
    return attention_bf16(qkv, out, lse, kmax, B, S, s);
  }
  static int attention_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* out, const __nv_bfloat16* d_out,
                           const float* lse, float* dvec, __nv_bfloat16* d_qkv, int B, int S, cudaStream_t s) {
    return attention_bwd_bf16(qkv, out, d_out, lse, dvec, d_qkv, B, S, s);
  }
};
template <>
struct Prec<float> {
  static int gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
                  cudaStream_t s) {
    return gemm_f32(static_cast<const float*>(A), lda, static_cast<const float*>(W), ldw, M, N, K, ep, s);
  }
  static int attention(const float* qkv, float* out, float* lse, float* /*kmax*/, int B, int S, cudaStream_t s) {
    return attention_f32(qkv, out, lse, B, S, s);
  }
  static int attention_bwd(const float* qkv, const float* out, const float* d_out, const float* lse, float* dvec,
                           float* d_qkv, int B, int S, cudaStream_t s) {
    return attention_bwd_f32(qkv, out, d_out, lse, dvec, d_qkv, B, S, s);
  }
};

template <typename T>
inline void set_act_out(GemmEpilogue& ep, T* p, int ld) {
  ep.out_bf16 = reinterpret_cast<__nv_bfloat16*>(p);
  ep.ld_bf16 = ld;
}

#define HS_TRY(expr)                 \
  do {                               \
    const int _rc = (expr);          \
    if (_rc != HS_OK) return _rc;    \
  } while (0)


}  // namespace
}  // namespace hs
