/* hsenet_b200.h -- C ABI of libhsenet_sm100a.so: the B200-native HSENet visual-encoding hot path.
 *
 * The reference (YanzhaoShi/HSENet) has no FFI layer: its seam is Python nn.Module duck typing
 * (Preprint/LaMed/src/model/multimodal_encoder/builder.py:4-11, multimodal_projector/builder.py:81-105).  The
 * Python facades in hsenet_b200/ keep those module signatures and call the entry points below through ctypes;
 * any other host (C++, Go/cgo, Rust FFI ...) can bind the same symbols.  Each entry point names the reference
 * code it replaces (paths relative to Preprint/LaMed/src/).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - no allocation, no synchronisation, no global mutable state: workspaces are caller-allocated
 *     (size from the *_workspace_bytes functions), calls are asynchronous on `stream`;
 *   - return value: HSENET_OK or a negative HSENET_ERR_* code (bad shape / misaligned pointer / CUDA launch
 *     failure).  Nothing aborts; there is no CPU fallback.
 *   - geometry is the one the reference hard-codes: volume 1x32x256x256, patch 4x16x16 -> 8x16x16 = 2048 tokens
 *     (+1 cls), hidden 768, 12 heads x 64, MLP 3072 (model/multimodal_projector/spatial_pooling_projector.py:140,
 *     model/lamed_arch.py:125, model/multimodal_encoder/vit.py:437).
 *   - precision: HSENET_PREC_BF16 = bf16 tensor-core kernels (tcgen05) with fp32 accumulation, fp32 residual
 *     stream and fp32 LayerNorm/softmax statistics; weights are bf16 copies ([out,in] row-major, as nn.Linear
 *     stores them), biases / LayerNorm / positional parameters stay fp32.
 *     HSENET_PREC_FP32_VERIFY = the same pipeline with fp32 operands and fp32 accumulation on CUDA cores
 *     (the north-star "fp32-accumulate verification mode", tolerance 1e-4); weights are the fp32 parameters.
 *     "act" below means bf16 in BF16 mode and fp32 in FP32_VERIFY mode.
 */
#ifndef HSENET_B200_H
#define HSENET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSENET_OK 0
#define HSENET_ERR_SHAPE (-1)  /* unsupported dimension (e.g. N not a multiple of 256) */
#define HSENET_ERR_ALIGN (-2)  /* pointer or leading dimension not 16-byte aligned */
#define HSENET_ERR_CUDA (-3)   /* kernel launch failed (cudaGetLastError) */
#define HSENET_ERR_ARG (-4)    /* null pointer / bad enum / workspace too small */
#define HSENET_ERR_DRIVER (-5) /* cuTensorMapEncodeTiled unavailable or failed */

#define HSENET_PREC_BF16 0
#define HSENET_PREC_FP32_VERIFY 1

#define HSENET_DTYPE_F32 0
#define HSENET_DTYPE_BF16 1
#define HSENET_DTYPE_F16 2

typedef void* hsenet_stream_t;

/* One MONAI TransformerBlock (constructed at vit.py:438-443).  w_* are act-typed [out,in]; the rest fp32. */
typedef struct hsenet_block_weights {
  const void* w_qkv;   /* attn.qkv.weight      [2304,768], no bias (qkv_bias=False, vit.py:440) */
  const void* w_out;   /* attn.out_proj.weight [768,768]  */
  const float* b_out;  /* attn.out_proj.bias   [768]      */
  const void* w_fc1;   /* mlp.linear1.weight   [3072,768] */
  const float* b_fc1;
  const void* w_fc2;   /* mlp.linear2.weight   [768,3072] */
  const float* b_fc2;
  const float* ln1_g;  /* norm1 */
  const float* ln1_b;
  const float* ln2_g;  /* norm2 */
  const float* ln2_b;
  /* Optional (bf16 precision only; leave NULL to run every LayerNorm as its own kernel): norm1 / norm2 folded into the
   * linear layer that follows them, as produced by hsenet_fold_layernorm from the fp32 parameters.  With these set,
   * norm2 of every block and norm1 of blocks >= 1 cost no kernel and no extra pass over the residual stream: the GEMM
   * that writes the residual also writes its bf16 copy and per-row (sum, sum of squares), and the qkv / linear1 GEMM
   * applies LN(x) W^T = rstd (x W'^T) - rstd mean colsum(W') + (W beta + b) in its epilogue. */
  const void* w_qkv_ln;   /* bf16 [2304,768]  gamma1 (.) attn.qkv.weight      */
  const float* cs_qkv;    /* [2304]           row sums of w_qkv_ln            */
  const float* b_qkv_ln;  /* [2304]           attn.qkv.weight . beta1         */
  const void* w_fc1_ln;   /* bf16 [3072,768]  gamma2 (.) mlp.linear1.weight   */
  const float* cs_fc1;    /* [3072]                                            */
  const float* b_fc1_ln;  /* [3072]           mlp.linear1.bias + weight . beta2 */
  /* qkv bias [2304]: NULL for the MONAI blocks of the 3D towers (vit.py passes qkv_bias=False); set for the timm blocks of
   * the slice trunk (hsenet_slice_trunk_forward). */
  const float* b_qkv;
} hsenet_block_weights;

/* ViT_stage1 (vit.py:360-469) when stage == 1, ViT_stage2 (vit.py:222-357) when stage == 2. */
typedef struct hsenet_vit_weights {
  int32_t stage;       /* 1 or 2 */
  int32_t num_layers;  /* 12 in every shipped config */
  const float* cls_token;                  /* [768]       cls_token (vit.py:447)                         */
  const float* pos_embed;                  /* [2048,768]  patch_embedding.position_embeddings            */
  const void* w_patch;                     /* [768,1024]  patch_embedding.patch_embeddings.1.weight      */
  const float* b_patch;                    /* [768]                                                      */
  const hsenet_block_weights* blocks_host; /* HOST array of num_layers entries                           */
  const float* norm_g;                     /* final LayerNorm (vit.py:445)                               */
  const float* norm_b;
  /* stage 2 only: slice_guided_attention = regular_attention (vit.py:36-64) + patch_score_proj (vit.py:308) */
  const void* w_sq;    /* Wq.weight [768,768] */
  const float* b_sq;
  const void* w_skv;   /* rows 0..767 = Wk.weight, rows 768..1535 = Wv.weight  -> [1536,768] */
  const float* b_skv;  /* [1536] = Wk.bias | Wv.bias */
  const void* w_so;    /* output_linear.weight [768,768] */
  const float* b_so;
  const float* sn_g;   /* slice_guided_attention.norm */
  const float* sn_b;
  const float* w_score; /* patch_score_proj.weight [768] */
  const float* b_score; /* patch_score_proj.bias   [1]   */
  /* Optional: the fp32 patch projection [768,1024].  When set (bf16 precision) the patch embedding runs as an
   * implicit-im2col tf32 tcgen05 GEMM fed by 5-D TMA boxes of the volume (no patch matrix is materialised); when NULL
   * it runs as im2col + GEMM on w_patch. */
  const float* w_patch_f32;
} hsenet_vit_weights;

/* hsenet_vit_forward flags */
#define HSENET_VIT_PATCH_DONE 1 /* the patch embedding of this call was already written into the workspace by
                                   hsenet_patch_embed_dual (bf16 precision only) */

/* VisualPacker_3d_phi_v3 (spatial_pooling_projector.py:121-153). */
typedef struct hsenet_packer_weights {
  int32_t out_dim;     /* 3072 for Phi-4-mini; must be a multiple of 256 */
  const void* w_q;     /* resolution_attention.Wq.weight [768,768] */
  const float* b_q;
  const void* w_kv;    /* Wk.weight | Wv.weight stacked -> [1536,768] */
  const float* b_kv;
  const void* w_o;     /* resolution_attention.output_linear.weight */
  const float* b_o;
  const float* ln_g;   /* resolution_attention.norm */
  const float* ln_b;
  const void* w_p0;    /* proj_mpls.0.weight [out_dim,768] */
  const float* b_p0;
  const void* w_p2;    /* proj_mpls.2.weight [out_dim,out_dim] */
  const float* b_p2;
} hsenet_packer_weights;

const char* hsenet_version(void);
const char* hsenet_error_string(int code);
/* Number of kernels launched by this library in this process so far (bench.py's `gpu_launches`). */
uint64_t hsenet_launch_count(void);

/* Per-kernel-class CUDA-event timing (used by bench.py's roofline pass; off by default, zero cost when off).
 * Classes: 0 = tcgen05 GEMM, 1 = self attention, 2 = LayerNorm, 3 = other, 4 = packer pooling, 5 = packer window
 * attention, 6 = patch im2col, 7 = slice cross attention, 8 = score gating, 9 = 2D-slice extraction.  stop()
 * synchronises the device and fills, per class: summed milliseconds, algorithmic FLOPs, algorithmic bytes, launch
 * count (arrays of HSENET_PROFILE_CLASSES). */
#define HSENET_PROFILE_CLASSES 10
void hsenet_profile_start(void);
int hsenet_profile_stop(double* ms, double* flops, double* bytes, uint64_t* launches);

/* ---- composite entry points ----------------------------------------------------------------------------------- */

/* Replaces ViT_stage1.forward (vit.py:449-469) / ViT_stage2.forward (vit.py:315-357).
 *   images      fp32 [B,1,32,256,256] contiguous
 *   images_2d   fp32 [B,32,768] (stage 2 only; the pre-computed slice features, dataset/multi_dataset.py:357-362)
 *   out_tokens  act  [B,2049,768]  final LayerNorm output (cls row first)            (may be NULL)
 *   out_patch   act  [B,2048,768]  the same without the cls row, contiguous -- what
 *               ViT3DTower_dual_encoders hands to the packer with select_feature == "patch" (vit.py:934-936) (may be NULL)
 *   hidden_f32  fp32 [num_layers,B,2049,768] per-block outputs (the second return value, vit.py:463-466) (may be NULL)
 *   scores_f32  fp32 [B,2048] sigmoid patch scores of stage 2 (vit.py:339)           (may be NULL)
 */
size_t hsenet_vit_workspace_bytes(int B, int precision, int stage);
int hsenet_vit_forward(const hsenet_vit_weights* w, const float* images, const float* images_2d, int B,
                       int precision, void* out_tokens, void* out_patch, float* hidden_f32, float* scores_f32,
                       void* workspace, size_t workspace_bytes, int flags, hsenet_stream_t stream);

/* Replaces VisualPacker_3d_phi_v3.forward (spatial_pooling_projector.py:138-146) and the torch.cat of
 * LamedMetaForCausalLM.encode_images (lamed_arch.py:132): writes packed token n of batch b to
 *   out[(b * out_tokens_per_batch + token_offset + n) * out_dim + :],  n in [0,128).
 *   hr   act [B,2048,768] contiguous (tower features without cls);  out_dtype HSENET_DTYPE_* of `out`
 *   (BF16 mode: BF16 or F32; FP32_VERIFY: F32). */
/* Patch embedding of ViT_stage1 AND ViT_stage2 in one launch (both read the same volume; vit.py:928-929 runs them on the
 * same `images`): implicit-im2col tf32 GEMM over the stacked fp32 patch projections w_stack_f32 [1536,1024] (stage 1
 * rows first), results written where hsenet_vit_forward(..., flags = HSENET_VIT_PATCH_DONE) of each tower expects them
 * inside its own workspace.  bf16 precision only. */
int hsenet_patch_embed_dual(const hsenet_vit_weights* w1, const hsenet_vit_weights* w2, const float* w_stack_f32,
                            const float* images, int B, void* workspace1, size_t workspace1_bytes, void* workspace2,
                            size_t workspace2_bytes, hsenet_stream_t stream);

size_t hsenet_packer_workspace_bytes(int B, int precision, int out_dim);
int hsenet_packer_forward(const hsenet_packer_weights* w, const void* hr, int B, int precision, void* out,
                          int out_dtype, int out_tokens_per_batch, int token_offset, void* workspace,
                          size_t workspace_bytes, hsenet_stream_t stream);

/* Replaces the image half of M3DCLIP_stage1.encode_image + [:,0] (CLIP_stage1.py:100-101,117):
 * out[b,:] = normalize(W * tokens[b,0,:] + bias)  -- only the cls row is projected (row 0 is numerically the
 * same as projecting all 2049 rows and slicing).  tokens act [B,2049,768]; w_proj act [768,768]; out fp32 [B,768];
 * workspace >= B*768*4 bytes. */
int hsenet_clip_image_head(const void* tokens, const void* w_proj, const float* b_proj, int B, int precision,
                           float* out, void* workspace, size_t workspace_bytes, hsenet_stream_t stream);

/* ---- operator-level entry points (unit-tested individually; the composites are built from these) ------------- */

/* out = epilogue(A[M,K] * W[N,K]^T): bias, optional GELU (the exact-erf function; bf16-only outputs evaluate a tanh form
 * fitted to it, max |err| 2.5e-5, below bf16 rounding), optional fp32 residual (may alias out_f32), fp32 and/or act outputs.  nn.Linear semantics (MONAI SABlock.qkv/out_proj, MLPBlock.linear1/2, ...). */
int hsenet_linear(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                  const float* resid, int ld_resid, int gelu, float* out_f32, int ld_f32, void* out_act,
                  int ld_act, int precision, hsenet_stream_t stream);
/* MONAI SABlock core: softmax(q k^T / 8) v per head; qkv act [B*S,2304] in (qkv, head, d) order -> out act [B*S,768]. */
int hsenet_self_attention(const void* qkv, void* out, int B, int S, int precision, hsenet_stream_t stream);
/* nn.LayerNorm(768), eps 1e-5, fp32 statistics; x fp32 [rows,768] -> out (F32 or BF16). */
int hsenet_layernorm(const float* x, const float* gamma, const float* beta, long rows, void* out, int out_dtype,
                     hsenet_stream_t stream);
/* MONAI PatchEmbeddingBlock rearrange (vit.py:437): images fp32 [B,1,32,256,256] -> out [B*2048,1024] (F32/BF16). */
int hsenet_patch_im2col(const float* images, int B, void* out, int out_dtype, hsenet_stream_t stream);
/* avg_pool3d kernel (1,4,4) over the 8x16x16 grid (spatial_pooling_projector.py:141): hr [B,2048,768] -> [B,128,768]. */
int hsenet_packer_pool(const void* hr, void* lr, int B, int dtype, hsenet_stream_t stream);
/* per-window single-head attention (spatial_pooling_projector.py:76 via :8-16): q fp32 [B*128,768],
 * kv [B*2048,1536] (dtype) -> out [B*128,768] (dtype). */
int hsenet_packer_window_attention(const float* q, const void* kv, void* out, int B, int dtype,
                                   hsenet_stream_t stream);
/* regular_attention core (vit.py:25-33,59): q fp32 [B*2048,768], kv fp32 [B*32,1536] -> out [B*2048,768] (dtype),
 * optional attention map fp32 [B,2048,32]. */
int hsenet_slice_cross_attention(const float* q, const float* kv, void* out, float* attn, int B, int dtype,
                                 hsenet_stream_t stream);
/* trilinear (32,256,256)->(32,out_h,out_w) + 3-channel expand + permute (vit.py:529-531):
 * images fp32 [B,1,32,256,256] -> out [B*32,3,out_h,out_w] (F32/BF16). */
int hsenet_slice_extract(const float* images, void* out, int B, int out_h, int out_w, int out_dtype,
                         hsenet_stream_t stream);
/* strided [B,rows,768] (any HSENET_DTYPE_*) -> contiguous [B*rows,768] (F32/BF16); strides in elements. */
int hsenet_gather_rows(const void* in, int in_dtype, long batch_stride, long row_stride, int B, int rows,
                       void* out, int out_dtype, hsenet_stream_t stream);
/* integer maps computed on the device exactly as the kernels index memory (bit-exact parity tests). */
int hsenet_patch_gather_map(int32_t* out /*[2048,1024]*/, hsenet_stream_t stream);
int hsenet_packer_window_map(int32_t* out /*[128,16]*/, hsenet_stream_t stream);
/* fp32 -> bf16 (round-to-nearest-even) cast used to build the weight cache. */
int hsenet_cast_bf16(const float* in, void* out, long n, hsenet_stream_t stream);
/* ---- volume ingest: Data/data_processing/CT-RATE/CT-RATE_nii_to_3D_volume_npy_file.py:25-117.
 * All four calls are asynchronous on `stream`; min / max and the foreground box stay in device memory, so the chain
 * raw volume -> [32,256,256] network input needs no host synchronisation. */
/* raw [n0,n1,n2] (NIfTI array order, slice index contiguous) -> out [o0,o1,o2] = trilinear(align_corners=False) resample
 * of clamp(slope*raw + intercept, hu_min, hu_max) viewed as [n2,n0,n1]  (lines 25-38, 73-91).
 * scratch: n0*n1*n2 floats of device memory (windowed transpose, then a coalesced resample) or NULL (one gather pass). */
int hsenet_hu_resample(const float* raw, int n0, int n1, int n2, float slope, float intercept, float hu_min,
                       float hu_max, float* out, int o0, int o1, int o2, float* scratch, hsenet_stream_t stream);
/* exact global min / max -> minmax2[0..1] (device); scratch2 = 2 ints of device scratch  (lines 103-104) */
int hsenet_minmax(const float* x, long n, float* minmax2, int32_t* scratch2, hsenet_stream_t stream);
/* MONAI CropForeground (select_fn x > 0 on the min-max normalised volume == x > min): bbox6 = lo0,lo1,lo2,hi0,hi1,hi2
 * with hi exclusive (device); the full extent if nothing is foreground  (line 116) */
int hsenet_foreground_bbox(const float* x, int d0, int d1, int d2, const float* minmax2, int32_t* bbox6,
                           hsenet_stream_t stream);
/* (x - min) / max(max - min, 1e-8) on the box, trilinear(align_corners=False) resize to [o0,o1,o2]  (lines 105-106, 117) */
int hsenet_crop_normalize_resize(const float* x, int d0, int d1, int d2, const float* minmax2, const int32_t* bbox6,
                                 float* out, int o0, int o1, int o2, hsenet_stream_t stream);
/* LayerNorm(gamma, beta) folded into the nn.Linear(w [N,K], bias [N] or NULL) that consumes it (weight cache):
 * w_folded[n,k] = bf16(gamma[k] w[n,k]), colsum[n] = sum_k w_folded[n,k], bias_folded[n] = bias[n] + sum_k w[n,k] beta[k]. */
int hsenet_fold_layernorm(const float* w, const float* gamma, const float* beta, const float* bias, int N, int K,
                          void* w_folded_bf16, float* colsum, float* bias_folded, hsenet_stream_t stream);

/* ================================================================================================================
 * Training path.  The reference trains both ViTs in CLIP stages 1-2
 * (train/train_CLIP_stage1.py:231-257, train_CLIP_stage2.py:239-267) and the two packers in the VLM stage
 * (train/train_VLM.py:406-414) through torch.autograd; here the façades wrap these entry points in
 * torch.autograd.Function.  A training forward saves its activations on a caller-allocated TAPE; the backward
 * consumes the tape and writes one fp32 gradient per parameter (pointers may be NULL to skip a gradient).
 * Deterministic: every reduction has a fixed order, there are no atomics.
 * Weight operands of the input-gradient GEMMs are the TRANSPOSED matrices ([in,out] row-major, activation dtype), which
 * the façade derives once per weight update (hsenet_transpose_weight).
 * ================================================================================================================ */
typedef struct hsenet_block_weights_t {
  const void* w_qkv_t; /* [768,2304]  attn.qkv.weight^T      */
  const void* w_out_t; /* [768,768]   attn.out_proj.weight^T */
  const void* w_fc1_t; /* [768,3072]  mlp.linear1.weight^T   */
  const void* w_fc2_t; /* [3072,768]  mlp.linear2.weight^T   */
} hsenet_block_weights_t;

typedef struct hsenet_vit_weights_t {
  const hsenet_block_weights_t* blocks_host; /* HOST array of num_layers entries */
  const void* w_sq_t;                        /* stage 2: Wq.weight^T [768,768]            */
  const void* w_so_t;                        /* stage 2: output_linear.weight^T [768,768] */
} hsenet_vit_weights_t;

typedef struct hsenet_block_grads {  /* fp32, shapes of the parameters; qkv has no bias */
  float *w_qkv, *w_out, *b_out, *w_fc1, *b_fc1, *w_fc2, *b_fc2, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
} hsenet_block_grads;

typedef struct hsenet_vit_grads {
  const hsenet_block_grads* blocks_host; /* HOST array of num_layers entries */
  float *cls_token, *pos_embed, *w_patch, *b_patch, *norm_g, *norm_b;
  /* stage 2 */
  float *w_sq, *b_sq, *w_skv, *b_skv, *w_so, *b_so, *sn_g, *sn_b, *w_score, *b_score;
} hsenet_vit_grads;

typedef struct hsenet_packer_weights_t {
  const void* w_q_t;  /* [768,768]       Wq.weight^T            */
  const void* w_kv_t; /* [768,1536]      (Wk | Wv).weight^T     (only used when d_hr is requested) */
  const void* w_o_t;  /* [768,768]       output_linear.weight^T */
  const void* w_p0_t; /* [768,out_dim]   proj_mpls.0.weight^T   */
  const void* w_p2_t; /* [out_dim,out_dim] proj_mpls.2.weight^T */
} hsenet_packer_weights_t;

typedef struct hsenet_packer_grads {
  float *w_q, *b_q, *w_kv, *b_kv, *w_o, *b_o, *ln_g, *ln_b, *w_p0, *b_p0, *w_p2, *b_p2;
} hsenet_packer_grads;

/* Train-mode dropout of the two small attentions: regular_attention (vit.py:25-33, 46-47, 60-62) and resolution_attention_v3
 * (spatial_pooling_projector.py:8-16, 58-59, 76-78) apply nn.Dropout(p = 0.1) to the attention probabilities and to the
 * output_linear result in front of the residual add.  The kernels use a counter-based keep mask: element i of the tensor is
 * kept when the top 32 bits of splitmix64(seed + (i + 1) * 0x9E3779B97F4A7C15) are < (1 - p) * 2^32, and scaled by
 * 1 / (1 - p); i runs over the row-major [rows, 32] / [windows, 16] probabilities (seed_attn) and the [rows, 768]
 * projection (seed_out).  The backward call must receive the forward's struct (the mask is regenerated, not stored).
 * NULL or p_* = 0: no dropout (eval).  Different random stream from torch's: statistically, not bitwise, the reference's. */
typedef struct hsenet_dropout {
  float p_attn;                  /* nn.Dropout on the attention probabilities, 0 <= p < 1 (0: off) */
  float p_out;                   /* nn.Dropout (dropout_2) on the output projection                */
  unsigned long long seed_attn;
  unsigned long long seed_out;
} hsenet_dropout;
/* out[i] = keep-mask value (0 or 1 / (1 - p)) of element i, fp32: what the kernels multiply by (tests, debugging). */
int hsenet_dropout_mask(float p, unsigned long long seed, long long n, float* out, hsenet_stream_t stream);

/* out[cols,rows] = in[rows,cols]^T, fp32 in, activation dtype out (HSENET_DTYPE_BF16 / _F32). */
int hsenet_transpose_weight(const float* in, int rows, int cols, void* out, int out_dtype, hsenet_stream_t stream);

size_t hsenet_vit_tape_bytes(int B, int precision, int stage, int num_layers);
size_t hsenet_vit_train_workspace_bytes(int B, int precision, int stage);
/* Same results contract as hsenet_vit_forward (without hidden states); additionally fills `tape`.  GELU of the MLP is the
 * exact erf form applied to the bf16-rounded pre-activation (the pre-activation is what the tape keeps). */
int hsenet_vit_forward_train(const hsenet_vit_weights* w, const float* images, const float* images_2d, int B,
                             int precision, void* out_tokens, void* out_patch, float* scores_f32, void* tape,
                             size_t tape_bytes, void* workspace, size_t workspace_bytes, const hsenet_dropout* dropout,
                             hsenet_stream_t stream);
/* d_tokens [B,2049,768] / d_patch [B,2048,768] in the activation dtype: gradients of the two outputs (either may be
 * NULL).  images: the forward's input (the patch matrix is re-derived from it, not taped). */
int hsenet_vit_backward(const hsenet_vit_weights* w, const hsenet_vit_weights_t* wt, const float* images, int B,
                        int precision, const void* d_tokens, const void* d_patch, const void* tape, size_t tape_bytes,
                        const hsenet_vit_grads* grads, void* workspace, size_t workspace_bytes,
                        const hsenet_dropout* dropout, hsenet_stream_t stream);

size_t hsenet_packer_tape_bytes(int B, int precision, int out_dim);
size_t hsenet_packer_train_workspace_bytes(int B, int precision, int out_dim);
/* out: [B,128,out_dim] contiguous in the activation dtype. */
int hsenet_packer_forward_train(const hsenet_packer_weights* w, const void* hr, int B, int precision, void* out,
                                void* tape, size_t tape_bytes, void* workspace, size_t workspace_bytes,
                                const hsenet_dropout* dropout, hsenet_stream_t stream);
/* d_out [B,128,out_dim] activation dtype; hr: the forward's input.  d_hr (optional): fp32 [B,2048,768] gradient of the
 * tower features (needs wt->w_kv_t). */
int hsenet_packer_backward(const hsenet_packer_weights* w, const hsenet_packer_weights_t* wt, const void* hr, int B,
                           int precision, const void* d_out, const void* tape, size_t tape_bytes,
                           const hsenet_packer_grads* grads, float* d_hr, void* workspace, size_t workspace_bytes,
                           const hsenet_dropout* dropout, hsenet_stream_t stream);

/* ---- online 2D-slice branch -------------------------------------------------------------
 * Replaces the offline JPG -> BiomedCLIP -> npy pipeline (Data/data_processing/CT-RATE/CT-RATE_2D_to_npy_file.py:75-98) and
 * the online formulation of ViT4LLM_v3_med2e3.forward (vit.py:805-808): the 32 slices of a volume are resized to 224x224
 * (trilinear with unit depth weight == per-slice bilinear), expanded to three identical channels, and encoded by a
 * ViT-B/16 trunk (timm vit_base_patch16_224 as instantiated by open_clip for BiomedCLIP: 196 patches + cls, 12 pre-LN
 * blocks with qkv bias, LayerNorm eps 1e-6, exact GELU, token pooling, no head) -> [B,32,768] slice features, the
 * `images_2d` input of ViT_stage2.  Runs on the same tcgen05 GEMM / attention kernels as the 3D towers (S = 197).
 * The three identical input channels are folded into the stem: w_patch_sum[o, iy*16+ix] = sum_c conv.weight[o,c,iy,ix]. */
typedef struct hsenet_trunk_weights {
  int32_t num_layers;
  float ln_eps;                            /* 1e-6 for timm ViTs */
  const void* w_patch_sum;                 /* activation dtype [768,256] */
  const float* b_patch;                    /* [768] patch_embed.proj.bias */
  const float* pos_patch;                  /* [196,768] pos_embed[0, 1:] */
  const float* cls_pos0;                   /* [768] cls_token + pos_embed[0, 0] */
  const hsenet_block_weights* blocks_host; /* HOST array; b_qkv set; LayerNorm-fold fields unused */
  const float* norm_g;                     /* final norm */
  const float* norm_b;
} hsenet_trunk_weights;

size_t hsenet_slice_trunk_workspace_bytes(int B, int precision);
/* images fp32 [B,1,32,256,256] -> out fp32 [B*32,768] (pooled cls token after the final norm). */
int hsenet_slice_trunk_forward(const hsenet_trunk_weights* w, const float* images, int B, int precision, float* out,
                               void* workspace, size_t workspace_bytes, hsenet_stream_t stream);

/* Operator-level: backward of hsenet_self_attention.  lse / dvec: fp32 [B,12,ceil(S/128)*128]; lse comes from
 * hsenet_self_attention_train. */
/* hsenet_self_attention with optional outputs / scratch: lse (optional, fp32 [B,12,ceil(S/128)*128]) and scratch (optional,
 * fp32 [B*12]).  With scratch AND the environment switch HSENET_ATT_MAXFREE=1 the bf16 kernel runs its max-free softmax: a
 * pre-pass stores the largest key norm per (volume, head) there and exp2(s - c |q| max|k|) replaces the running-maximum
 * bookkeeping wherever that bound is <= 50 (log2 units); rows with a looser bound take the exact online softmax.  Results
 * differ from the exact path only by fp32 rounding.  (Opt-in: measured not faster on B200.) */
int hsenet_self_attention_ws(const void* qkv, void* out, float* lse, float* scratch, int B, int S, int precision,
                             hsenet_stream_t stream);
int hsenet_self_attention_train(const void* qkv, void* out, float* lse, int B, int S, int precision,
                                hsenet_stream_t stream);
int hsenet_self_attention_backward(const void* qkv, const void* out, const void* d_out, const float* lse, float* dvec,
                                   void* d_qkv, int B, int S, int precision, hsenet_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HSENET_B200_H */
