// hsenet_b200 -- internal launcher prototypes shared by the .cu files (the public C ABI is include/hsenet_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hsenet_b200.h"

namespace hs {

enum : int {
  HS_OK = HSENET_OK,
  HS_ERR_SHAPE = HSENET_ERR_SHAPE,
  HS_ERR_ALIGN = HSENET_ERR_ALIGN,
  HS_ERR_CUDA = HSENET_ERR_CUDA,
  HS_ERR_ARG = HSENET_ERR_ARG,
  HS_ERR_DRIVER = HSENET_ERR_DRIVER,
};

// Fused GEMM epilogue:   v = acc + bias[col] (+ row_add[row % rows_per_group, col]) (+ resid[orow, col]);
//                        v = gelu(v) if gelu;  out_f32[orow, col] = v;  out_act[orow, col] = (Act)v
// orow = row, or (row / rows_per_group) * group_stride + group_offset + row % rows_per_group when rows_per_group>0.
// `out_act` is bf16 for the tcgen05 GEMM and fp32 for the fp32 verification GEMM.
struct GemmEpilogue {
  const float* bias = nullptr;
  const float* row_add = nullptr;   // [rows_per_group, N] fp32 (positional embedding)
  const float* resid = nullptr;     // fp32, indexed by orow; may alias out_f32
  int ld_resid = 0;
  float* out_f32 = nullptr;
  int ld_f32 = 0;
  __nv_bfloat16* out_bf16 = nullptr;  // "out_act"
  int ld_bf16 = 0;
  int gelu = 0;
  int rows_per_group = 0;
  int group_stride = 0;
  int group_offset = 0;
  // backward of GELU fused into the GEMM that produces the gradient of the hidden activation (generic epilogue only):
  //   v *= gelu'(dgelu_src[orow, col]),  dgelu_src = the saved PRE-activation in the activation dtype (bf16 / fp32)
  const void* dgelu_src = nullptr;
  int ld_dgelu = 0;
  // LayerNorm folded into the GEMMs on either side of it (gemm_epilogue.cuh):
  float2* stats_out = nullptr;        // producer (EPI_F32_RESID): (sum, sum of squares) of every output row over each 128-
                                      //   column slab, stored at [(col / 128) * ld_stats + row]  (N <= 1024)
  const float2* stats_in = nullptr;   // consumer (EPI_BF16[_GELU]): those partials of the un-normalised A operand
  long ld_stats = 0;                  // rows between two slab slots
  int stats_slots = 0;                // consumer: slots to add per row (= producer N / 128, <= 8)
  const float* colsum = nullptr;      // consumer: sum_k W'[n,k] of the gain-folded weight
  float ln_inv_dim = 0.f;             // 1 / (row length the statistics were taken over)
  float ln_eps = 0.f;
};

// RAII CUDA-event bracket around one launch, active only between hsenet_profile_start/stop.
enum : int {
  PROF_GEMM = 0, PROF_ATTENTION = 1, PROF_LAYERNORM = 2, PROF_OTHER = 3, PROF_PACKER_POOL = 4, PROF_PACKER_WATTN = 5,
  PROF_IM2COL = 6, PROF_SLICE_XATTN = 7, PROF_SCORE_SCALE = 8, PROF_SLICE_EXTRACT = 9,
};
struct ProfScope {
  ProfScope(int cls, double flops, double bytes, cudaStream_t st);
  ~ProfScope();
  cudaStream_t st_;
  long idx_;
};

// Launch with the programmatic-stream-serialization attribute (PDL) unless HSENET_PDL=0.  Only for kernels that call
// pdl_prologue_done() before touching global memory.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int num_sms();
// true exactly once per (call site slot, current device): kernel attributes are per device, and one process may
// drive several GPUs.  `slot` must point to a zero-initialised static array of kMaxDevices flags.
constexpr int kMaxDevices = 64;
bool first_use_on_device(unsigned char* slot);
int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                      uint32_t box_inner, uint32_t box_outer);

// ---- dense contractions -------------------------------------------------------------------------------------
int gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
              cudaStream_t stream);
// Split-K variant for the weight gradients: fp32 partials [slices][M][N]; *used = number of slices written.
int gemm_bf16_splitk(const void* A, int lda, const void* W, int ldw, int M, int N, int K, float* partials, int ksplit,
                     cudaStream_t stream, int* used);
// fp32 verification GEMM (CUDA cores, fp32 operands and accumulation); ep.out_bf16 is reinterpreted as float*.
int gemm_f32(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const GemmEpilogue& ep,
             cudaStream_t stream);

// Implicit-im2col patch embedding of one or both towers (patch_embed_tcgen05.cu): tf32 tcgen05 GEMM whose A tiles are
// 5-D TMA boxes of the fp32 volume; w_stack fp32 [n_towers*768,1024]; ep0 / ep1 = the towers' epilogues.
int patch_embed_tf32(const float* images, const float* w_stack, int B, int n_towers, const GemmEpilogue& ep0,
                     const GemmEpilogue& ep1, cudaStream_t stream);

// ---- self attention over the 3D token sequence (MONAI SABlock core) --------------------------------------------
// qkv [B*S, 2304] with feature order (qkv, head, d); out [B*S, 768] heads concatenated.
// lse (optional, training forward): [B, 12, S_pad] fp32 with S_pad = ceil(S/128)*128, log2-domain log-sum-exp of the scaled
// scores (m + log2 l); entries S <= q < S_pad are written as +inf.
// kmax_scratch (optional): fp32 [B*12]; when given, a pre-pass stores the largest squared key norm per (volume, head) there and
// the kernel runs its max-free softmax (attention_tcgen05.cu) where the Cauchy-Schwarz bound allows it.
int attention_bf16(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse, float* kmax_scratch, int B, int S,
                   cudaStream_t stream);
int attention_f32(const float* qkv, float* out, float* lse, int B, int S, cudaStream_t stream);
// Backward of the fused attention (recompute style: P is rebuilt from qkv and lse).  d_out [B*S,768] (gradient of the
// re-concatenated heads), out [B*S,768] (the forward result) -> d_qkv [B*S,2304].  dvec: scratch [B,12,S_pad] fp32.
int attention_bwd_bf16(const __nv_bfloat16* qkv, const __nv_bfloat16* out, const __nv_bfloat16* d_out, const float* lse,
                       float* dvec, __nv_bfloat16* d_qkv, int B, int S, cudaStream_t stream);
int attention_bwd_f32(const float* qkv, const float* out, const float* d_out, const float* lse, float* dvec, float* d_qkv,
                      int B, int S, cudaStream_t stream);

// ---- row kernels ------------------------------------------------------------------------------------------------
template <typename OutT>
int layernorm_rows(const float* x, long ldx, const float* gamma, const float* beta, long rows, OutT* out, long ldo,
                   OutT* out2_patch_only, int seq, cudaStream_t stream, float eps = 1e-5f);
template <typename OutT>
int im2col_patches(const float* vol, int B, OutT* out, cudaStream_t stream);
template <typename OutT>
int cast_rows(const float* in, OutT* out, long n, cudaStream_t stream);
// volume ingest (ingest.cu)
int hu_resample(const float* raw, int n0, int n1, int n2, float slope, float intercept, float hu_min, float hu_max,
                float* out, int o0, int o1, int o2, float* scratch, cudaStream_t st);
int minmax(const float* x, long n, float* minmax2, int* scratch2, cudaStream_t st);
int foreground_bbox(const float* x, int d0, int d1, int d2, const float* minmax2, int* bbox6, cudaStream_t st);
int crop_normalize_resize(const float* x, int d0, int d1, int d2, const float* minmax2, const int* bbox6, float* out,
                          int o0, int o1, int o2, cudaStream_t st);
int fold_layernorm(const float* w, const float* gamma, const float* beta, const float* bias, int N, int K,
                   __nv_bfloat16* w_folded, float* colsum, float* bias_folded, cudaStream_t stream);
int write_cls_rows(float* X, const float* cls, int B, int seq, cudaStream_t stream);
// gather arbitrary-strided [B, rows, 768] (fp32 / bf16 / fp16 tagged by dtype code) into contiguous Act rows
template <typename OutT>
int gather_rows(const void* in, int in_dtype, long batch_stride, long row_stride, int B, int rows, OutT* out,
                cudaStream_t stream);

// ---- 2E3 slice-guided scoring (vit.py:332-345) ------------------------------------------------------------------
// Q [B*2048,768] fp32 (projected query), KV [B*32,1536] fp32 (Wk|Wv of the slice features) -> O Act [B*2048,768]
// Train-mode dropout of the two small attentions (vit.py:31-32,62 / spatial_pooling_projector.py:14-15,78): element `idx` of a
// tensor is kept when the top 32 bits of splitmix64(seed + idx) are below keep_thresh and scaled by scale = 1 / (1 - p).
// A counter-based mask: the backward kernels regenerate it from (seed, idx) instead of storing it.  scale == 0: disabled.
struct DropSpec {
  unsigned long long seed = 0;
  unsigned int keep_thresh = 0;
  float scale = 0.f;
  __host__ __device__ bool on() const { return scale != 0.f; }
};
DropSpec make_dropspec(float p, unsigned long long seed);       // p <= 0: disabled
// z[i] = resid[i] + mask(i) * z[i]  (dropout_2 in front of the residual add);  out[i] = mask(i) * in[i];  out[i] = mask(i)
int dropout_residual(float* z, const float* resid, long n, DropSpec dr, cudaStream_t st);
int dropout_scale(const float* in, float* out, long n, DropSpec dr, cudaStream_t st);
int dropout_mask(float* out, long n, DropSpec dr, cudaStream_t st);
template <typename OutT>
int slice_cross_attention(const float* Q, const float* KV, OutT* O, float* attn_opt, int B, cudaStream_t stream,
                          DropSpec dr = DropSpec());
// Z = Wq(x)+out_proj(...) [B*2048,768] fp32 -> LN -> . w_s + b_s -> sigmoid -> X[b,1+t,:] = XP[b,t,:] * score
int score_and_scale(const float* Z, const float* ln_g, const float* ln_b, const float* w_s, const float* b_s,
                    const float* XP, float* X, float* scores_opt, int B, cudaStream_t stream);

// ---- spatial packer (spatial_pooling_projector.py:121-153) -----------------------------------------------------------
template <typename T>
int packer_pool(const T* HR, T* LR, int B, cudaStream_t stream);
// Q [B*128,768] fp32; KV [B*2048,1536] Act (Wk|Wv of the HR tokens, natural token order) -> O Act [B*128,768]
template <typename T>
int packer_window_attention(const float* Q, const T* KV, T* O, int B, cudaStream_t stream, DropSpec dr = DropSpec());

// ---- CLIP head (CLIP_stage1.py:100-101,117) ------------------------------------------------------------------------
int l2_normalize_rows(const float* in, float* out, int rows, int dim, cudaStream_t stream);

// ---- 2D slice extraction (vit.py:529-531) ---------------------------------------------------------------------------
template <typename OutT>
int slice_extract(const float* vol, OutT* out, int B, int out_h, int out_w, cudaStream_t stream);

// ---- training path: row kernels of the backward pass (backward.cu) -----------------------------------------------------
int padded_rows(int M);                       // ceil(M / 128) * 128
// in [M,N] (row stride ld_in) -> out_t [N, padded_rows(M)] (zero padded), optional cast copy [M,N] (ld_copy), optional exact GELU
// applied first, optional per-64-row-tile column sums colsum_partial [padded_rows(M)/64, N] (finish with colsum_finish)
template <typename TIn, typename TOut>
int transpose_pad(const TIn* in, long ld_in, int M, int N, TOut* out_t, TOut* copy_out, long ld_copy, int gelu,
                  float* colsum_partial, cudaStream_t st);
int colsum_finish(const float* partial, int T, int N, float* out, cudaStream_t st);
template <typename T>
int gelu_rows(const T* in, T* out, long n, cudaStream_t st);
int layernorm_bwd_blocks(long rows);          // number of partial rows layernorm_bwd / score_scale_bwd write
int layernorm_bwd(const float* dy, const float* x, long ldx, const float* gamma, long rows, float* dx, long ld_dx,
                  int accumulate, float* dgamma_partial, float* dbeta_partial, cudaStream_t st);
template <typename T>
int combine_final_grad(const T* d_tokens, const T* d_patch, int B, int seq, float* dy, cudaStream_t st);
int sum_over_batch(const float* in, long batch_stride, int B, int rows, float* out, cudaStream_t st);
template <typename T>
int window_attn_bwd(const T* dO, const float* Q, const T* KV, float* dQ, T* dKV, int B, cudaStream_t st,
                    DropSpec dr = DropSpec());
template <typename TG>
int pool_bwd(const TG* dLR, float* dHR, int B, int accumulate, cudaStream_t st);
template <typename T>
int slice_xattn_bwd(const float* Q, const float* KV, const T* dO, float* dQ, int accumulate, float* P, float* dS,
                    float* dKV, int B, cudaStream_t st, DropSpec dr = DropSpec());
int score_scale_bwd(const float* dX, const float* XP, const float* Z, const float* g, const float* be, const float* ws,
                    const float* scores, float* dXP, float* dZ, float* dg_partial, float* db_partial,
                    float* dws_partial, float* dbs_partial, int B, cudaStream_t st);
int add_rows(float* out, const float* a, long n, cudaStream_t st);

// slice branch (row f-2): resized slices of the volume as the [B*32*196, 256] patch matrix of a ViT-B/16 stem
template <typename OutT>
int slice_patches(const float* vol, OutT* out, int B, cudaStream_t stream);

// integer maps (device-computed, for bit-exact tests against the oracle's closed forms)
int patch_gather_map(int32_t* out, cudaStream_t stream);      // [2048,1024]
int packer_window_map(int32_t* out, cudaStream_t stream);     // [128,16]

}  // namespace hs
