// Training path of the hot path: activation-taping forward and backward of ViT_stage1 /
// ViT_stage2 (vit.py:449-469, 315-357) and of VisualPacker_3d_phi_v3 (spatial_pooling_projector.py:121-153), built from
// the same tcgen05 GEMM / attention kernels as the inference path plus the row kernels of backward.cu.
//
// Gradient GEMMs.  For y = x W^T + b with x [M,K], W [N,K]:
//     dx = dy W      -> GEMM(A = dy [M,N],        "weight" = W^T [K,N])                (W^T is kept by the façade)
//     dW = dy^T x    -> GEMM(A = dy^T [N,Mpad],   "weight" = x^T [K,Mpad]) -> fp32 [N,K]  (transpose_pad, zero padded rows)
//     db = colsum dy -> per-tile partial sums emitted by the transpose of dy, finished in fixed order
// so every contraction of the backward runs on the forward's GEMM kernel (both operands K-major) without atomics.
#include "common.cuh"
#include "composite.cuh"
#include "kernels.h"

#include <type_traits>

namespace hs {
namespace {

constexpr size_t kDwPartialFloats = 6u << 20;     // split-K partials of one weight gradient (24 MB)

// ---- shared scratch of a training forward / backward ------------------------------------------------------------------
template <typename T>
struct TrainWs {
  float* dX;      // [M,768]   running gradient of the residual stream
  float* dXN;     // [M,768]   fp32 GEMM output feeding a LayerNorm backward (also: stage-2 dQ)
  T* dYb;         // [M,768]   activation-dtype copy of a residual-stream gradient
  T* dH;          // [M,3072]  gradient of the MLP hidden pre-activation / of qkv [M,2304]
  T* dATT;        // [M,768]
  T* TA;          // [3072,Mpad]  transposed gradient (A operand of a weight-gradient GEMM)
  T* TB;          // [3072,Mpad]  transposed activation (B operand)
  T* H2;          // [M,3072]  forward: GELU output;  also the im2col patch matrix [Mp,1024]
  float* DVEC;    // [B,12,S_pad]
  float* CS;      // [Mpad/64,3072]  column-sum partials
  float* LNP;     // [2][nblk,768]   LayerNorm gain / shift partials
  // stage 2
  float* dXP;     // [Mp,768]
  float* dZ;      // [Mp,768]
  float* Pm;      // [Mp,32]
  float* dSm;     // [Mp,32]
  float* dSKV;    // [B*32,1536]
  float* SP;      // [3][nblkp,768] + [nblkp]  score partials
  float* DWP;     // split-K partials of a weight gradient
  size_t total;
  int Mpad, nblk, nblkp;
  TrainWs(void* base, int B) {
    const size_t M = static_cast<size_t>(B) * kSeq, Mp = static_cast<size_t>(B) * kNPatch;
    Mpad = padded_rows(static_cast<int>(M));
    nblk = layernorm_bwd_blocks(static_cast<long>(M));
    nblkp = layernorm_bwd_blocks(static_cast<long>(Mp));
    Bump b(base);
    dX = b.take<float>(M * kHidden);
    dXN = b.take<float>(M * kHidden);
    dYb = b.take<T>(M * kHidden);
    dH = b.take<T>(M * kMlp);
    dATT = b.take<T>(M * kHidden);
    TA = b.take<T>(static_cast<size_t>(kMlp) * Mpad);
    TB = b.take<T>(static_cast<size_t>(kMlp) * Mpad);
    H2 = b.take<T>(M * kMlp);
    DVEC = b.take<float>(static_cast<size_t>(B) * kHeads * padded_rows(kSeq));
    CS = b.take<float>(static_cast<size_t>(Mpad / 64) * kMlp);
    LNP = b.take<float>(2 * static_cast<size_t>(nblk) * kHidden);
    dXP = b.take<float>(Mp * kHidden);
    dZ = b.take<float>(Mp * kHidden);
    Pm = b.take<float>(Mp * kNSlice);
    dSm = b.take<float>(Mp * kNSlice);
    dSKV = b.take<float>(static_cast<size_t>(B) * kNSlice * 2 * kHidden);
    SP = b.take<float>(3 * static_cast<size_t>(nblkp) * kHidden + nblkp);
    DWP = b.take<float>(kDwPartialFloats);
    total = b.off;
  }
};

// ---- activation tape of one ViT forward -------------------------------------------------------------------------------
constexpr int kMaxTapeLayers = 64;
template <typename T>
struct VitTape {
  float* X[2 * kMaxTapeLayers + 1];   // residual stream: X[2l] enters block l, X[2l+1] enters its MLP half, X[2L] the final norm
  T* XN1[kMaxTapeLayers];             // [M,768]   norm1 output
  T* QKV[kMaxTapeLayers];             // [M,2304]
  T* ATT[kMaxTapeLayers];             // [M,768]   attention output (heads concatenated)
  float* LSE[kMaxTapeLayers];         // [B,12,S_pad]
  T* XN2[kMaxTapeLayers];             // [M,768]   norm2 output
  T* H1[kMaxTapeLayers];              // [M,3072]  MLP pre-activation
  // stage 2 (slice-guided scoring, vit.py:332-345)
  float* XP;      // [Mp,768]  patch embedding
  T* XPa;         // [Mp,768]  ... in the activation dtype (A operand of Wq)
  T* S16;         // [B*32,768]
  float* SKV;     // [B*32,1536]
  float* Q;       // [Mp,768]  Wq(x)
  T* O;           // [Mp,768]  attention output
  float* Z;       // [Mp,768]  Wq(x) + output_linear(o): input of the LayerNorm
  float* scores;  // [Mp]
  size_t total;
  VitTape(void* base, int B, int stage, int L) {
    const size_t M = static_cast<size_t>(B) * kSeq, Mp = static_cast<size_t>(B) * kNPatch;
    Bump b(base);
    for (int i = 0; i <= 2 * L; ++i) X[i] = b.take<float>(M * kHidden);
    for (int l = 0; l < L; ++l) {
      XN1[l] = b.take<T>(M * kHidden);
      QKV[l] = b.take<T>(M * 3 * kHidden);
      ATT[l] = b.take<T>(M * kHidden);
      LSE[l] = b.take<float>(static_cast<size_t>(B) * kHeads * padded_rows(kSeq));
      XN2[l] = b.take<T>(M * kHidden);
      H1[l] = b.take<T>(M * kMlp);
    }
    XP = nullptr; XPa = nullptr; S16 = nullptr; SKV = nullptr; Q = nullptr; O = nullptr; Z = nullptr; scores = nullptr;
    if (stage == 2) {
      XP = b.take<float>(Mp * kHidden);
      XPa = b.take<T>(Mp * kHidden);
      S16 = b.take<T>(static_cast<size_t>(B) * kNSlice * kHidden);
      SKV = b.take<float>(static_cast<size_t>(B) * kNSlice * 2 * kHidden);
      Q = b.take<float>(Mp * kHidden);
      O = b.take<T>(Mp * kHidden);
      Z = b.take<float>(Mp * kHidden);
      scores = b.take<float>(Mp);
    }
    total = b.off;
  }
};

// dW [N,K] fp32 = dY^T X from the transposed operands dYt [N,Mpad], Xt [K,Mpad].  The output has few tiles (9 .. 36 of
// 256x256) and a long contraction (the token dimension): on the tensor-core path the contraction is split into slices that
// run as independent tiles and write fp32 partials, summed in fixed order (deterministic; one wave of 74 CTA pairs instead
// of 9 .. 36 busy pairs -- 122 us per weight gradient before).
template <typename T>
int grad_weight(const T* dYt, const T* Xt, int N, int K, int Mpad, float* dW, float* partials, cudaStream_t st) {
  if (dW == nullptr) return HS_OK;
  if (std::is_same<T, __nv_bfloat16>::value && partials != nullptr && Mpad % 128 == 0 && N > 128) {
    const int tiles = ((N + 255) / 256) * (K / 256);
    int split = tiles > 0 ? (num_sms() / 2) / tiles : 1;
    if (split > 8) split = 8;
    while (split > 1 && static_cast<size_t>(split) * N * K > kDwPartialFloats) --split;
    if (split > 1) {
      int used = 1;
      HS_TRY(gemm_bf16_splitk(dYt, Mpad, Xt, Mpad, N, K, Mpad, partials, split, st, &used));
      return colsum_finish(partials, used, N * K, dW, st);
    }
  }
  GemmEpilogue ep;
  ep.out_f32 = dW; ep.ld_f32 = K;
  return Prec<T>::gemm(dYt, Mpad, Xt, Mpad, N, K, Mpad, ep, st);
}

// ----------------------------------------------------------------------------------------------------------------------
// ViT
// ----------------------------------------------------------------------------------------------------------------------
template <typename T>
int vit_forward_train(const hsenet_vit_weights* w, const float* images, const float* images_2d, int B, T* out_tokens,
                      T* out_patch, float* scores_out, void* tape_mem, size_t tape_bytes, void* workspace,
                      size_t workspace_bytes, const hsenet_dropout* drop, cudaStream_t st) {
  const DropSpec drop_attn = drop ? make_dropspec(drop->p_attn, drop->seed_attn) : DropSpec();
  const DropSpec drop_out = drop ? make_dropspec(drop->p_out, drop->seed_out) : DropSpec();
  const int L = w->num_layers;
  if (L > kMaxTapeLayers) return HS_ERR_ARG;
  VitTape<T> tp(tape_mem, B, w->stage, L);
  TrainWs<T> ws(workspace, B);
  if (tape_bytes < tp.total || workspace_bytes < ws.total) return HS_ERR_ARG;
  const int M = B * kSeq, Mp = B * kNPatch;
  T* P = ws.H2;                                       // im2col patches [Mp,1024]
  HS_TRY(im2col_patches<T>(images, B, P, st));
  if (w->stage == 1) {
    GemmEpilogue ep;
    ep.bias = w->b_patch; ep.row_add = w->pos_embed;
    ep.rows_per_group = kNPatch; ep.group_stride = kSeq; ep.group_offset = 1;
    ep.out_f32 = tp.X[0]; ep.ld_f32 = kHidden;
    HS_TRY(Prec<T>::gemm(P, kPatchDim, w->w_patch, kPatchDim, Mp, kHidden, kPatchDim, ep, st));
  } else {
    if (images_2d == nullptr) return HS_ERR_ARG;
    {
      GemmEpilogue ep;
      ep.bias = w->b_patch; ep.row_add = w->pos_embed;
      ep.rows_per_group = kNPatch; ep.group_stride = kNPatch; ep.group_offset = 0;
      ep.out_f32 = tp.XP; ep.ld_f32 = kHidden;
      set_act_out(ep, tp.XPa, kHidden);
      HS_TRY(Prec<T>::gemm(P, kPatchDim, w->w_patch, kPatchDim, Mp, kHidden, kPatchDim, ep, st));
    }
    HS_TRY(cast_rows<T>(images_2d, tp.S16, static_cast<long>(B) * kNSlice * kHidden, st));
    {
      GemmEpilogue ep;
      ep.bias = w->b_skv; ep.out_f32 = tp.SKV; ep.ld_f32 = 2 * kHidden;
      HS_TRY(Prec<T>::gemm(tp.S16, kHidden, w->w_skv, kHidden, B * kNSlice, 2 * kHidden, kHidden, ep, st));
    }
    {
      GemmEpilogue ep;
      ep.bias = w->b_sq; ep.out_f32 = tp.Q; ep.ld_f32 = kHidden;
      HS_TRY(Prec<T>::gemm(tp.XPa, kHidden, w->w_sq, kHidden, Mp, kHidden, kHidden, ep, st));
    }
    HS_TRY(slice_cross_attention<T>(tp.Q, tp.SKV, tp.O, nullptr, B, st, drop_attn));
    {
      GemmEpilogue ep;   // Z = Wq(x) + dropout_2(output_linear(o))   (vit.py:61-62)
      ep.bias = w->b_so; ep.out_f32 = tp.Z; ep.ld_f32 = kHidden;
      if (!drop_out.on()) { ep.resid = tp.Q; ep.ld_resid = kHidden; }
      HS_TRY(Prec<T>::gemm(tp.O, kHidden, w->w_so, kHidden, Mp, kHidden, kHidden, ep, st));
      if (drop_out.on()) HS_TRY(dropout_residual(tp.Z, tp.Q, static_cast<long>(Mp) * kHidden, drop_out, st));
    }
    HS_TRY(score_and_scale(tp.Z, w->sn_g, w->sn_b, w->w_score, w->b_score, tp.XP, tp.X[0], tp.scores, B, st));
    if (scores_out != nullptr &&
        cudaMemcpyAsync(scores_out, tp.scores, static_cast<size_t>(Mp) * sizeof(float), cudaMemcpyDeviceToDevice, st) !=
            cudaSuccess)
      return HS_ERR_CUDA;
  }
  HS_TRY(write_cls_rows(tp.X[0], w->cls_token, B, kSeq, st));
  for (int l = 0; l < L; ++l) {
    const hsenet_block_weights& bw = w->blocks_host[l];
    HS_TRY(layernorm_rows<T>(tp.X[2 * l], kHidden, bw.ln1_g, bw.ln1_b, M, tp.XN1[l], kHidden, nullptr, kSeq, st));
    {
      GemmEpilogue ep;
      set_act_out(ep, tp.QKV[l], 3 * kHidden);
      HS_TRY(Prec<T>::gemm(tp.XN1[l], kHidden, bw.w_qkv, kHidden, M, 3 * kHidden, kHidden, ep, st));
    }
    HS_TRY(Prec<T>::attention(tp.QKV[l], tp.ATT[l], tp.LSE[l], ws.DVEC, B, kSeq, st));   // DVEC doubles as key-norm scratch
    {
      GemmEpilogue ep;
      ep.bias = bw.b_out; ep.resid = tp.X[2 * l]; ep.ld_resid = kHidden; ep.out_f32 = tp.X[2 * l + 1]; ep.ld_f32 = kHidden;
      HS_TRY(Prec<T>::gemm(tp.ATT[l], kHidden, bw.w_out, kHidden, M, kHidden, kHidden, ep, st));
    }
    HS_TRY(layernorm_rows<T>(tp.X[2 * l + 1], kHidden, bw.ln2_g, bw.ln2_b, M, tp.XN2[l], kHidden, nullptr, kSeq, st));
    {
      GemmEpilogue ep;   // pre-activation only: the tape keeps it, GELU is a separate pass here
      ep.bias = bw.b_fc1;
      set_act_out(ep, tp.H1[l], kMlp);
      HS_TRY(Prec<T>::gemm(tp.XN2[l], kHidden, bw.w_fc1, kHidden, M, kMlp, kHidden, ep, st));
    }
    HS_TRY(gelu_rows<T>(tp.H1[l], ws.H2, static_cast<long>(M) * kMlp, st));
    {
      GemmEpilogue ep;
      ep.bias = bw.b_fc2; ep.resid = tp.X[2 * l + 1]; ep.ld_resid = kHidden; ep.out_f32 = tp.X[2 * l + 2]; ep.ld_f32 = kHidden;
      HS_TRY(Prec<T>::gemm(ws.H2, kMlp, bw.w_fc2, kMlp, M, kHidden, kMlp, ep, st));
    }
  }
  if (out_tokens != nullptr || out_patch != nullptr)
    HS_TRY(layernorm_rows<T>(tp.X[2 * L], kHidden, w->norm_g, w->norm_b, M, out_tokens, kHidden, out_patch, kSeq, st));
  return HS_OK;
}

template <typename T>
int vit_backward(const hsenet_vit_weights* w, const hsenet_vit_weights_t* wt, const float* images, int B,
                 const T* d_tokens, const T* d_patch, void* tape_mem, size_t tape_bytes, const hsenet_vit_grads* g,
                 void* workspace, size_t workspace_bytes, const hsenet_dropout* drop, cudaStream_t st) {
  const DropSpec drop_attn = drop ? make_dropspec(drop->p_attn, drop->seed_attn) : DropSpec();
  const DropSpec drop_out = drop ? make_dropspec(drop->p_out, drop->seed_out) : DropSpec();
  const int L = w->num_layers;
  if (L > kMaxTapeLayers) return HS_ERR_ARG;
  VitTape<T> tp(tape_mem, B, w->stage, L);
  TrainWs<T> ws(workspace, B);
  if (tape_bytes < tp.total || workspace_bytes < ws.total) return HS_ERR_ARG;
  const int M = B * kSeq, Mp = B * kNPatch, Mpad = ws.Mpad, Mppad = padded_rows(Mp);
  float* lnp_g = ws.LNP;
  float* lnp_b = ws.LNP + static_cast<size_t>(ws.nblk) * kHidden;

  // final LayerNorm (vit.py:467): dy = d_tokens (+ d_patch behind the cls row)
  HS_TRY(combine_final_grad<T>(d_tokens, d_patch, B, kSeq, ws.dXN, st));
  HS_TRY(layernorm_bwd(ws.dXN, tp.X[2 * L], kHidden, w->norm_g, M, ws.dX, kHidden, 0, g->norm_g ? lnp_g : nullptr,
                       g->norm_b ? lnp_b : nullptr, st));
  if (g->norm_g) HS_TRY(colsum_finish(lnp_g, ws.nblk, kHidden, g->norm_g, st));
  if (g->norm_b) HS_TRY(colsum_finish(lnp_b, ws.nblk, kHidden, g->norm_b, st));

  for (int l = L - 1; l >= 0; --l) {
    const hsenet_block_weights& bw = w->blocks_host[l];
    const hsenet_block_weights_t& bt = wt->blocks_host[l];
    const hsenet_block_grads& bg = g->blocks_host[l];
    // ---- MLP half: x_out = x_mid + linear2(gelu(linear1(norm2(x_mid)))) ----
    // dy2 = dX: activation-dtype copy, transpose, bias gradient
    HS_TRY((transpose_pad<float, T>(ws.dX, kHidden, M, kHidden, ws.TA, ws.dYb, kHidden, 0, bg.b_fc2 ? ws.CS : nullptr, st)));
    if (bg.b_fc2) HS_TRY(colsum_finish(ws.CS, Mpad / 64, kHidden, bg.b_fc2, st));
    // dW2 = dy2^T gelu(h1): the hidden activation is recomputed from the taped pre-activation while it is transposed
    HS_TRY((transpose_pad<T, T>(tp.H1[l], kMlp, M, kMlp, ws.TB, nullptr, 0, 1, nullptr, st)));
    HS_TRY(grad_weight<T>(ws.TA, ws.TB, kHidden, kMlp, Mpad, bg.w_fc2, ws.DWP, st));
    {
      GemmEpilogue ep;   // dh1 = (dy2 W2) o gelu'(h1)
      set_act_out(ep, ws.dH, kMlp);
      ep.dgelu_src = tp.H1[l]; ep.ld_dgelu = kMlp;
      HS_TRY(Prec<T>::gemm(ws.dYb, kHidden, bt.w_fc2_t, kHidden, M, kMlp, kHidden, ep, st));
    }
    HS_TRY((transpose_pad<T, T>(ws.dH, kMlp, M, kMlp, ws.TA, nullptr, 0, 0, bg.b_fc1 ? ws.CS : nullptr, st)));
    if (bg.b_fc1) HS_TRY(colsum_finish(ws.CS, Mpad / 64, kMlp, bg.b_fc1, st));
    HS_TRY((transpose_pad<T, T>(tp.XN2[l], kHidden, M, kHidden, ws.TB, nullptr, 0, 0, nullptr, st)));
    HS_TRY(grad_weight<T>(ws.TA, ws.TB, kMlp, kHidden, Mpad, bg.w_fc1, ws.DWP, st));
    {
      GemmEpilogue ep;   // d norm2 output = dh1 W1
      ep.out_f32 = ws.dXN; ep.ld_f32 = kHidden;
      HS_TRY(Prec<T>::gemm(ws.dH, kMlp, bt.w_fc1_t, kMlp, M, kHidden, kMlp, ep, st));
    }
    HS_TRY(layernorm_bwd(ws.dXN, tp.X[2 * l + 1], kHidden, bw.ln2_g, M, ws.dX, kHidden, 1, bg.ln2_g ? lnp_g : nullptr,
                         bg.ln2_b ? lnp_b : nullptr, st));
    if (bg.ln2_g) HS_TRY(colsum_finish(lnp_g, ws.nblk, kHidden, bg.ln2_g, st));
    if (bg.ln2_b) HS_TRY(colsum_finish(lnp_b, ws.nblk, kHidden, bg.ln2_b, st));

    // ---- attention half: x_mid = x_in + out_proj(attention(qkv(norm1(x_in)))) ----
    HS_TRY((transpose_pad<float, T>(ws.dX, kHidden, M, kHidden, ws.TA, ws.dYb, kHidden, 0, bg.b_out ? ws.CS : nullptr, st)));
    if (bg.b_out) HS_TRY(colsum_finish(ws.CS, Mpad / 64, kHidden, bg.b_out, st));
    HS_TRY((transpose_pad<T, T>(tp.ATT[l], kHidden, M, kHidden, ws.TB, nullptr, 0, 0, nullptr, st)));
    HS_TRY(grad_weight<T>(ws.TA, ws.TB, kHidden, kHidden, Mpad, bg.w_out, ws.DWP, st));
    {
      GemmEpilogue ep;   // d attention output = dy W_out
      set_act_out(ep, ws.dATT, kHidden);
      HS_TRY(Prec<T>::gemm(ws.dYb, kHidden, bt.w_out_t, kHidden, M, kHidden, kHidden, ep, st));
    }
    T* dQKV = ws.dH;     // [M,2304]
    HS_TRY(Prec<T>::attention_bwd(tp.QKV[l], tp.ATT[l], ws.dATT, tp.LSE[l], ws.DVEC, dQKV, B, kSeq, st));
    HS_TRY((transpose_pad<T, T>(dQKV, 3 * kHidden, M, 3 * kHidden, ws.TA, nullptr, 0, 0, nullptr, st)));
    HS_TRY((transpose_pad<T, T>(tp.XN1[l], kHidden, M, kHidden, ws.TB, nullptr, 0, 0, nullptr, st)));
    HS_TRY(grad_weight<T>(ws.TA, ws.TB, 3 * kHidden, kHidden, Mpad, bg.w_qkv, ws.DWP, st));
    {
      GemmEpilogue ep;   // d norm1 output = dqkv W_qkv
      ep.out_f32 = ws.dXN; ep.ld_f32 = kHidden;
      HS_TRY(Prec<T>::gemm(dQKV, 3 * kHidden, bt.w_qkv_t, 3 * kHidden, M, kHidden, 3 * kHidden, ep, st));
    }
    HS_TRY(layernorm_bwd(ws.dXN, tp.X[2 * l], kHidden, bw.ln1_g, M, ws.dX, kHidden, 1, bg.ln1_g ? lnp_g : nullptr,
                         bg.ln1_b ? lnp_b : nullptr, st));
    if (bg.ln1_g) HS_TRY(colsum_finish(lnp_g, ws.nblk, kHidden, bg.ln1_g, st));
    if (bg.ln1_b) HS_TRY(colsum_finish(lnp_b, ws.nblk, kHidden, bg.ln1_b, st));
  }

  // ---- embedding: cls token, (stage 2: score gating), positional embedding, patch projection ----
  if (g->cls_token) HS_TRY(sum_over_batch(ws.dX, static_cast<long>(kSeq) * kHidden, B, 1, g->cls_token, st));
  const float* dXPtot = nullptr;     // gradient of the patch embedding output, [B,2048,768] with batch stride `dxp_bs`
  long dxp_bs = 0;
  if (w->stage == 1) {
    dXPtot = ws.dX + kHidden;        // rows 1.. of every volume
    dxp_bs = static_cast<long>(kSeq) * kHidden;
  } else {
    float* sp_g = ws.SP;
    float* sp_b = ws.SP + static_cast<size_t>(ws.nblkp) * kHidden;
    float* sp_w = ws.SP + 2 * static_cast<size_t>(ws.nblkp) * kHidden;
    float* sp_bs = ws.SP + 3 * static_cast<size_t>(ws.nblkp) * kHidden;
    HS_TRY(score_scale_bwd(ws.dX, tp.XP, tp.Z, w->sn_g, w->sn_b, w->w_score, tp.scores, ws.dXP, ws.dZ, sp_g, sp_b, sp_w,
                           sp_bs, B, st));
    if (g->sn_g) HS_TRY(colsum_finish(sp_g, ws.nblkp, kHidden, g->sn_g, st));
    if (g->sn_b) HS_TRY(colsum_finish(sp_b, ws.nblkp, kHidden, g->sn_b, st));
    if (g->w_score) HS_TRY(colsum_finish(sp_w, ws.nblkp, kHidden, g->w_score, st));
    if (g->b_score) HS_TRY(colsum_finish(sp_bs, ws.nblkp, 1, g->b_score, st));
    // Z = Q + dropout_2(output_linear(O)):  dZ -> output_linear (weight, bias, dO) and the residual branch into Q
    const float* dlin = ws.dZ;
    if (drop_out.on()) {             // gradient entering output_linear = mask o dZ (the residual branch keeps dZ)
      HS_TRY(dropout_scale(ws.dZ, ws.dXN, static_cast<long>(Mp) * kHidden, drop_out, st));
      dlin = ws.dXN;
    }
    HS_TRY((transpose_pad<float, T>(dlin, kHidden, Mp, kHidden, ws.TA, ws.dYb, kHidden, 0, g->b_so ? ws.CS : nullptr, st)));
    if (g->b_so) HS_TRY(colsum_finish(ws.CS, Mppad / 64, kHidden, g->b_so, st));
    HS_TRY((transpose_pad<T, T>(tp.O, kHidden, Mp, kHidden, ws.TB, nullptr, 0, 0, nullptr, st)));
    HS_TRY(grad_weight<T>(ws.TA, ws.TB, kHidden, kHidden, Mppad, g->w_so, ws.DWP, st));
    {
      GemmEpilogue ep;   // dO = dZ W_so
      set_act_out(ep, ws.dATT, kHidden);
      HS_TRY(Prec<T>::gemm(ws.dYb, kHidden, wt->w_so_t, kHidden, Mp, kHidden, kHidden, ep, st));
    }
    // attention core: dQ = dZ (residual) + dQ_attention;  dK | dV of the 32 slice keys
    float* dQ = ws.dZ;               // in place: the kernel adds its contribution to dZ
    HS_TRY(slice_xattn_bwd<T>(tp.Q, tp.SKV, ws.dATT, dQ, 1, ws.Pm, ws.dSm, ws.dSKV, B, st, drop_attn));
    // Wk | Wv of the slice features: dW = dSKV^T S16, db = colsum(dSKV)
    if (g->w_skv != nullptr || g->b_skv != nullptr) {
      const int R = B * kNSlice, Rpad = padded_rows(R);
      HS_TRY((transpose_pad<float, T>(ws.dSKV, 2 * kHidden, R, 2 * kHidden, ws.TA, nullptr, 0, 0,
                                      g->b_skv ? ws.CS : nullptr, st)));
      if (g->b_skv) HS_TRY(colsum_finish(ws.CS, Rpad / 64, 2 * kHidden, g->b_skv, st));
      HS_TRY((transpose_pad<T, T>(tp.S16, kHidden, R, kHidden, ws.TB, nullptr, 0, 0, nullptr, st)));
      HS_TRY(grad_weight<T>(ws.TA, ws.TB, 2 * kHidden, kHidden, Rpad, g->w_skv, ws.DWP, st));
    }
    // Wq: dW = dQ^T XPa, db = colsum(dQ), dXP += dQ Wq
    HS_TRY((transpose_pad<float, T>(dQ, kHidden, Mp, kHidden, ws.TA, ws.dYb, kHidden, 0, g->b_sq ? ws.CS : nullptr, st)));
    if (g->b_sq) HS_TRY(colsum_finish(ws.CS, Mppad / 64, kHidden, g->b_sq, st));
    HS_TRY((transpose_pad<T, T>(tp.XPa, kHidden, Mp, kHidden, ws.TB, nullptr, 0, 0, nullptr, st)));
    HS_TRY(grad_weight<T>(ws.TA, ws.TB, kHidden, kHidden, Mppad, g->w_sq, ws.DWP, st));
    {
      GemmEpilogue ep;   // dXP (score path) = dQ Wq, added to the gating path already in dXP
      ep.resid = ws.dXP; ep.ld_resid = kHidden; ep.out_f32 = ws.dXP; ep.ld_f32 = kHidden;
      HS_TRY(Prec<T>::gemm(ws.dYb, kHidden, wt->w_sq_t, kHidden, Mp, kHidden, kHidden, ep, st));
    }
    dXPtot = ws.dXP;
    dxp_bs = static_cast<long>(kNPatch) * kHidden;
  }
  if (g->pos_embed) HS_TRY(sum_over_batch(dXPtot, dxp_bs, B, kNPatch, g->pos_embed, st));
  if (g->w_patch != nullptr || g->b_patch != nullptr) {
    // contiguous [Mp,768] view of dXP for the transposes (stage 1: rows are strided by the cls row)
    const float* src = dXPtot;
    if (w->stage == 1) {
      for (int b = 0; b < B; ++b)
        if (cudaMemcpyAsync(ws.dXN + static_cast<size_t>(b) * kNPatch * kHidden, dXPtot + b * dxp_bs,
                            static_cast<size_t>(kNPatch) * kHidden * sizeof(float), cudaMemcpyDeviceToDevice, st) !=
            cudaSuccess)
          return HS_ERR_CUDA;
      src = ws.dXN;
    }
    HS_TRY((transpose_pad<float, T>(src, kHidden, Mp, kHidden, ws.TA, nullptr, 0, 0, g->b_patch ? ws.CS : nullptr, st)));
    if (g->b_patch) HS_TRY(colsum_finish(ws.CS, Mppad / 64, kHidden, g->b_patch, st));
    if (g->w_patch) {
      T* P = ws.H2;
      HS_TRY(im2col_patches<T>(images, B, P, st));
      HS_TRY((transpose_pad<T, T>(P, kPatchDim, Mp, kPatchDim, ws.TB, nullptr, 0, 0, nullptr, st)));
      HS_TRY(grad_weight<T>(ws.TA, ws.TB, kHidden, kPatchDim, Mppad, g->w_patch, ws.DWP, st));
    }
  }
  return HS_OK;
}

// ----------------------------------------------------------------------------------------------------------------------
// packer
// ----------------------------------------------------------------------------------------------------------------------
template <typename T>
struct PackerTape {
  T* LR;        // [B*128,768]   pooled tokens
  T* KV;        // [B*2048,1536] Wk | Wv of the HR tokens
  float* Q;     // [B*128,768]   Wq(LR)
  T* O;         // [B*128,768]   window attention output
  float* Z;     // [B*128,768]   Wq(LR) + output_linear(o): LayerNorm input
  T* A;         // [B*128,768]   LayerNorm output
  T* H1;        // [B*128,D]     proj_mpls.0 pre-activation
  size_t total;
  PackerTape(void* base, int B, int D) {
    const size_t n = static_cast<size_t>(B) * 128;
    Bump b(base);
    LR = b.take<T>(n * kHidden);
    KV = b.take<T>(static_cast<size_t>(B) * kNPatch * 2 * kHidden);
    Q = b.take<float>(n * kHidden);
    O = b.take<T>(n * kHidden);
    Z = b.take<float>(n * kHidden);
    A = b.take<T>(n * kHidden);
    H1 = b.take<T>(n * D);
    total = b.off;
  }
};

template <typename T>
struct PackerTrainWs {
  T* H2;        // [Mw,D]       forward: gelu output
  T* dYb;       // [Mw,D]       activation-dtype gradient rows
  T* dH;        // [Mw,D]
  float* dF;    // [Mw,768]     fp32 GEMM outputs (dA, dZ, dQ)
  float* dQ;    // [Mw,768]
  T* dO;        // [Mw,768]
  T* dKV;       // [Mp,1536]
  T* TA;        // [max(D,1536), max(Mwpad, Mppad)]
  T* TB;        // [max(D,768), max(Mwpad, Mppad)]
  float* CS;    // [Mppad/64, max(D,1536)]
  float* LNP;   // [2][nblk,768]
  float* DWP;   // split-K partials of a weight gradient
  size_t total;
  int Mwpad, Mppad, nblk;
  PackerTrainWs(void* base, int B, int D) {
    const size_t Mw = static_cast<size_t>(B) * 128, Mp = static_cast<size_t>(B) * kNPatch;
    Mwpad = padded_rows(static_cast<int>(Mw));
    Mppad = padded_rows(static_cast<int>(Mp));
    nblk = layernorm_bwd_blocks(static_cast<long>(Mw));
    const size_t wide = static_cast<size_t>(D > 2 * kHidden ? D : 2 * kHidden);
    Bump b(base);
    H2 = b.take<T>(Mw * D);
    dYb = b.take<T>(Mw * D);
    dH = b.take<T>(Mw * D);
    dF = b.take<float>(Mw * kHidden);
    dQ = b.take<float>(Mw * kHidden);
    dO = b.take<T>(Mw * kHidden);
    dKV = b.take<T>(Mp * 2 * kHidden);
    TA = b.take<T>(wide * Mppad);
    TB = b.take<T>(wide * Mppad);
    CS = b.take<float>(static_cast<size_t>(Mppad / 64) * wide);
    LNP = b.take<float>(2 * static_cast<size_t>(nblk) * kHidden);
    DWP = b.take<float>(kDwPartialFloats);
    total = b.off;
  }
};

template <typename T>
int packer_forward_train(const hsenet_packer_weights* w, const T* hr, int B, T* out, void* tape_mem, size_t tape_bytes,
                         void* workspace, size_t workspace_bytes, const hsenet_dropout* drop, cudaStream_t st) {
  const DropSpec drop_attn = drop ? make_dropspec(drop->p_attn, drop->seed_attn) : DropSpec();
  const DropSpec drop_out = drop ? make_dropspec(drop->p_out, drop->seed_out) : DropSpec();
  const int D = w->out_dim;
  if (D <= 0 || D % 256 != 0) return HS_ERR_SHAPE;
  PackerTape<T> tp(tape_mem, B, D);
  PackerTrainWs<T> ws(workspace, B, D);
  if (tape_bytes < tp.total || workspace_bytes < ws.total) return HS_ERR_ARG;
  const int Mp = B * kNPatch, Mw = B * 128;
  HS_TRY(packer_pool<T>(hr, tp.LR, B, st));
  {
    GemmEpilogue ep;
    ep.bias = w->b_kv; set_act_out(ep, tp.KV, 2 * kHidden);
    HS_TRY(Prec<T>::gemm(hr, kHidden, w->w_kv, kHidden, Mp, 2 * kHidden, kHidden, ep, st));
  }
  {
    GemmEpilogue ep;
    ep.bias = w->b_q; ep.out_f32 = tp.Q; ep.ld_f32 = kHidden;
    HS_TRY(Prec<T>::gemm(tp.LR, kHidden, w->w_q, kHidden, Mw, kHidden, kHidden, ep, st));
  }
  HS_TRY(packer_window_attention<T>(tp.Q, tp.KV, tp.O, B, st, drop_attn));
  {
    GemmEpilogue ep;   // Z = Wq(LR) + dropout_2(output_linear(o))   (spatial_pooling_projector.py:77-78)
    ep.bias = w->b_o; ep.out_f32 = tp.Z; ep.ld_f32 = kHidden;
    if (!drop_out.on()) { ep.resid = tp.Q; ep.ld_resid = kHidden; }
    HS_TRY(Prec<T>::gemm(tp.O, kHidden, w->w_o, kHidden, Mw, kHidden, kHidden, ep, st));
    if (drop_out.on()) HS_TRY(dropout_residual(tp.Z, tp.Q, static_cast<long>(Mw) * kHidden, drop_out, st));
  }
  HS_TRY(layernorm_rows<T>(tp.Z, kHidden, w->ln_g, w->ln_b, Mw, tp.A, kHidden, nullptr, 128, st));
  {
    GemmEpilogue ep;
    ep.bias = w->b_p0; set_act_out(ep, tp.H1, D);
    HS_TRY(Prec<T>::gemm(tp.A, kHidden, w->w_p0, kHidden, Mw, D, kHidden, ep, st));
  }
  HS_TRY(gelu_rows<T>(tp.H1, ws.H2, static_cast<long>(Mw) * D, st));
  {
    GemmEpilogue ep;
    ep.bias = w->b_p2; set_act_out(ep, out, D);
    HS_TRY(Prec<T>::gemm(ws.H2, D, w->w_p2, D, Mw, D, D, ep, st));
  }
  return HS_OK;
}

template <typename T>
int packer_backward(const hsenet_packer_weights* w, const hsenet_packer_weights_t* wt, const T* hr, int B, const T* d_out,
                    void* tape_mem, size_t tape_bytes, const hsenet_packer_grads* g, float* d_hr, void* workspace,
                    size_t workspace_bytes, const hsenet_dropout* drop, cudaStream_t st) {
  const DropSpec drop_attn = drop ? make_dropspec(drop->p_attn, drop->seed_attn) : DropSpec();
  const DropSpec drop_out = drop ? make_dropspec(drop->p_out, drop->seed_out) : DropSpec();
  const int D = w->out_dim;
  if (D <= 0 || D % 256 != 0) return HS_ERR_SHAPE;
  PackerTape<T> tp(tape_mem, B, D);
  PackerTrainWs<T> ws(workspace, B, D);
  if (tape_bytes < tp.total || workspace_bytes < ws.total) return HS_ERR_ARG;
  const int Mp = B * kNPatch, Mw = B * 128, Mwpad = ws.Mwpad, Mppad = ws.Mppad;
  float* lnp_g = ws.LNP;
  float* lnp_b = ws.LNP + static_cast<size_t>(ws.nblk) * kHidden;
  // proj_mpls.2: dW = d_out^T gelu(h1), db, dh2 = d_out W2
  HS_TRY((transpose_pad<T, T>(d_out, D, Mw, D, ws.TA, nullptr, 0, 0, g->b_p2 ? ws.CS : nullptr, st)));
  if (g->b_p2) HS_TRY(colsum_finish(ws.CS, Mwpad / 64, D, g->b_p2, st));
  HS_TRY((transpose_pad<T, T>(tp.H1, D, Mw, D, ws.TB, nullptr, 0, 1, nullptr, st)));
  HS_TRY(grad_weight<T>(ws.TA, ws.TB, D, D, Mwpad, g->w_p2, ws.DWP, st));
  {
    GemmEpilogue ep;   // dh1 = (d_out W2) o gelu'(h1)
    set_act_out(ep, ws.dH, D);
    ep.dgelu_src = tp.H1; ep.ld_dgelu = D;
    HS_TRY(Prec<T>::gemm(d_out, D, wt->w_p2_t, D, Mw, D, D, ep, st));
  }
  // proj_mpls.0
  HS_TRY((transpose_pad<T, T>(ws.dH, D, Mw, D, ws.TA, nullptr, 0, 0, g->b_p0 ? ws.CS : nullptr, st)));
  if (g->b_p0) HS_TRY(colsum_finish(ws.CS, Mwpad / 64, D, g->b_p0, st));
  HS_TRY((transpose_pad<T, T>(tp.A, kHidden, Mw, kHidden, ws.TB, nullptr, 0, 0, nullptr, st)));
  HS_TRY(grad_weight<T>(ws.TA, ws.TB, D, kHidden, Mwpad, g->w_p0, ws.DWP, st));
  {
    GemmEpilogue ep;   // dA = dh1 W_p0
    ep.out_f32 = ws.dF; ep.ld_f32 = kHidden;
    HS_TRY(Prec<T>::gemm(ws.dH, D, wt->w_p0_t, D, Mw, kHidden, D, ep, st));
  }
  // LayerNorm(Z): dZ
  HS_TRY(layernorm_bwd(ws.dF, tp.Z, kHidden, w->ln_g, Mw, ws.dQ, kHidden, 0, g->ln_g ? lnp_g : nullptr,
                       g->ln_b ? lnp_b : nullptr, st));
  if (g->ln_g) HS_TRY(colsum_finish(lnp_g, ws.nblk, kHidden, g->ln_g, st));
  if (g->ln_b) HS_TRY(colsum_finish(lnp_b, ws.nblk, kHidden, g->ln_b, st));
  // Z = Q + dropout_2(output_linear(O)): ws.dQ holds dZ = the residual part of dQ; output_linear sees mask o dZ
  const float* dlin = ws.dQ;
  if (drop_out.on()) {
    HS_TRY(dropout_scale(ws.dQ, ws.dF, static_cast<long>(Mw) * kHidden, drop_out, st));
    dlin = ws.dF;
  }
  HS_TRY((transpose_pad<float, T>(dlin, kHidden, Mw, kHidden, ws.TA, ws.dYb, kHidden, 0, g->b_o ? ws.CS : nullptr, st)));
  if (g->b_o) HS_TRY(colsum_finish(ws.CS, Mwpad / 64, kHidden, g->b_o, st));
  HS_TRY((transpose_pad<T, T>(tp.O, kHidden, Mw, kHidden, ws.TB, nullptr, 0, 0, nullptr, st)));
  HS_TRY(grad_weight<T>(ws.TA, ws.TB, kHidden, kHidden, Mwpad, g->w_o, ws.DWP, st));
  {
    GemmEpilogue ep;   // dO = dZ W_o
    set_act_out(ep, ws.dO, kHidden);
    HS_TRY(Prec<T>::gemm(ws.dYb, kHidden, wt->w_o_t, kHidden, Mw, kHidden, kHidden, ep, st));
  }
  // per-window attention: dQ_attn (into dF), dKV
  HS_TRY(window_attn_bwd<T>(ws.dO, tp.Q, tp.KV, ws.dF, ws.dKV, B, st, drop_attn));
  HS_TRY(add_rows(ws.dQ, ws.dF, static_cast<long>(Mw) * kHidden, st));       // dQ = dZ + dQ_attn
  // Wq(LR)
  HS_TRY((transpose_pad<float, T>(ws.dQ, kHidden, Mw, kHidden, ws.TA, ws.dYb, kHidden, 0, g->b_q ? ws.CS : nullptr, st)));
  if (g->b_q) HS_TRY(colsum_finish(ws.CS, Mwpad / 64, kHidden, g->b_q, st));
  HS_TRY((transpose_pad<T, T>(tp.LR, kHidden, Mw, kHidden, ws.TB, nullptr, 0, 0, nullptr, st)));
  HS_TRY(grad_weight<T>(ws.TA, ws.TB, kHidden, kHidden, Mwpad, g->w_q, ws.DWP, st));
  // Wk | Wv over the HR tokens
  HS_TRY((transpose_pad<T, T>(ws.dKV, 2 * kHidden, Mp, 2 * kHidden, ws.TA, nullptr, 0, 0, g->b_kv ? ws.CS : nullptr, st)));
  if (g->b_kv) HS_TRY(colsum_finish(ws.CS, Mppad / 64, 2 * kHidden, g->b_kv, st));
  if (g->w_kv) {
    HS_TRY((transpose_pad<T, T>(hr, kHidden, Mp, kHidden, ws.TB, nullptr, 0, 0, nullptr, st)));
    HS_TRY(grad_weight<T>(ws.TA, ws.TB, 2 * kHidden, kHidden, Mppad, g->w_kv, ws.DWP, st));
  }
  if (d_hr != nullptr) {
    if (wt->w_kv_t == nullptr || wt->w_q_t == nullptr) return HS_ERR_ARG;
    {
      GemmEpilogue ep;   // through Wk | Wv
      ep.out_f32 = d_hr; ep.ld_f32 = kHidden;
      HS_TRY(Prec<T>::gemm(ws.dKV, 2 * kHidden, wt->w_kv_t, 2 * kHidden, Mp, kHidden, 2 * kHidden, ep, st));
    }
    {
      GemmEpilogue ep;   // through Wq and the (1,4,4) average pooling
      ep.out_f32 = ws.dF; ep.ld_f32 = kHidden;
      HS_TRY(Prec<T>::gemm(ws.dYb, kHidden, wt->w_q_t, kHidden, Mw, kHidden, kHidden, ep, st));
    }
    HS_TRY(pool_bwd<float>(ws.dF, d_hr, B, 1, st));
  }
  return HS_OK;
}

template <typename TOut>
__global__ void transpose_weight_kernel(const float* __restrict__ in, int rows, int cols, TOut* __restrict__ out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[static_cast<long>(r) * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) {
      if constexpr (sizeof(TOut) == 2) out[static_cast<long>(c) * rows + r] = __float2bfloat16(tile[threadIdx.x][i]);
      else out[static_cast<long>(c) * rows + r] = tile[threadIdx.x][i];
    }
  }
}

}  // namespace
extern void count_launch();
}  // namespace hs

using namespace hs;

extern "C" {

int hsenet_transpose_weight(const float* in, int rows, int cols, void* out, int out_dtype, hsenet_stream_t stream) {
  if (in == nullptr || out == nullptr || rows <= 0 || cols <= 0) return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  if (out_dtype == HSENET_DTYPE_BF16)
    transpose_weight_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(in, rows, cols, static_cast<__nv_bfloat16*>(out));
  else if (out_dtype == HSENET_DTYPE_F32)
    transpose_weight_kernel<float><<<grid, block, 0, st>>>(in, rows, cols, static_cast<float*>(out));
  else
    return HSENET_ERR_ARG;
  count_launch();
  return cudaGetLastError() == cudaSuccess ? HSENET_OK : HSENET_ERR_CUDA;
}

size_t hsenet_vit_tape_bytes(int B, int precision, int stage, int num_layers) {
  if (B <= 0 || num_layers < 0 || num_layers > kMaxTapeLayers) return 0;
  if (precision == HSENET_PREC_BF16) return VitTape<__nv_bfloat16>(nullptr, B, stage, num_layers).total;
  if (precision == HSENET_PREC_FP32_VERIFY) return VitTape<float>(nullptr, B, stage, num_layers).total;
  return 0;
}

size_t hsenet_vit_train_workspace_bytes(int B, int precision, int stage) {
  (void)stage;
  if (B <= 0) return 0;
  if (precision == HSENET_PREC_BF16) return TrainWs<__nv_bfloat16>(nullptr, B).total;
  if (precision == HSENET_PREC_FP32_VERIFY) return TrainWs<float>(nullptr, B).total;
  return 0;
}

int hsenet_vit_forward_train(const hsenet_vit_weights* w, const float* images, const float* images_2d, int B,
                             int precision, void* out_tokens, void* out_patch, float* scores_f32, void* tape,
                             size_t tape_bytes, void* workspace, size_t workspace_bytes, const hsenet_dropout* dropout,
                             hsenet_stream_t stream) {
  if (w == nullptr || images == nullptr || tape == nullptr || workspace == nullptr || w->blocks_host == nullptr)
    return HSENET_ERR_ARG;
  if (dropout != nullptr && !(dropout->p_attn >= 0.f && dropout->p_attn < 1.f && dropout->p_out >= 0.f && dropout->p_out < 1.f))
    return HSENET_ERR_ARG;
  if (B <= 0 || w->num_layers < 0 || (w->stage != 1 && w->stage != 2)) return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision == HSENET_PREC_BF16)
    return vit_forward_train<__nv_bfloat16>(w, images, images_2d, B, static_cast<__nv_bfloat16*>(out_tokens),
                                            static_cast<__nv_bfloat16*>(out_patch), scores_f32, tape, tape_bytes,
                                            workspace, workspace_bytes, dropout, st);
  if (precision == HSENET_PREC_FP32_VERIFY)
    return vit_forward_train<float>(w, images, images_2d, B, static_cast<float*>(out_tokens),
                                    static_cast<float*>(out_patch), scores_f32, tape, tape_bytes, workspace,
                                    workspace_bytes, dropout, st);
  return HSENET_ERR_ARG;
}

int hsenet_vit_backward(const hsenet_vit_weights* w, const hsenet_vit_weights_t* wt, const float* images, int B,
                        int precision, const void* d_tokens, const void* d_patch, const void* tape, size_t tape_bytes,
                        const hsenet_vit_grads* grads, void* workspace, size_t workspace_bytes,
                        const hsenet_dropout* dropout, hsenet_stream_t stream) {
  if (w == nullptr || wt == nullptr || images == nullptr || tape == nullptr || grads == nullptr || workspace == nullptr ||
      w->blocks_host == nullptr || wt->blocks_host == nullptr || grads->blocks_host == nullptr)
    return HSENET_ERR_ARG;
  if (B <= 0 || (w->stage != 1 && w->stage != 2) || (d_tokens == nullptr && d_patch == nullptr)) return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* tp = const_cast<void*>(tape);
  if (precision == HSENET_PREC_BF16)
    return vit_backward<__nv_bfloat16>(w, wt, images, B, static_cast<const __nv_bfloat16*>(d_tokens),
                                       static_cast<const __nv_bfloat16*>(d_patch), tp, tape_bytes, grads, workspace,
                                       workspace_bytes, dropout, st);
  if (precision == HSENET_PREC_FP32_VERIFY)
    return vit_backward<float>(w, wt, images, B, static_cast<const float*>(d_tokens), static_cast<const float*>(d_patch),
                               tp, tape_bytes, grads, workspace, workspace_bytes, dropout, st);
  return HSENET_ERR_ARG;
}

size_t hsenet_packer_tape_bytes(int B, int precision, int out_dim) {
  if (B <= 0 || out_dim <= 0) return 0;
  if (precision == HSENET_PREC_BF16) return PackerTape<__nv_bfloat16>(nullptr, B, out_dim).total;
  if (precision == HSENET_PREC_FP32_VERIFY) return PackerTape<float>(nullptr, B, out_dim).total;
  return 0;
}

size_t hsenet_packer_train_workspace_bytes(int B, int precision, int out_dim) {
  if (B <= 0 || out_dim <= 0) return 0;
  if (precision == HSENET_PREC_BF16) return PackerTrainWs<__nv_bfloat16>(nullptr, B, out_dim).total;
  if (precision == HSENET_PREC_FP32_VERIFY) return PackerTrainWs<float>(nullptr, B, out_dim).total;
  return 0;
}

int hsenet_packer_forward_train(const hsenet_packer_weights* w, const void* hr, int B, int precision, void* out,
                                void* tape, size_t tape_bytes, void* workspace, size_t workspace_bytes,
                                const hsenet_dropout* dropout, hsenet_stream_t stream) {
  if (w == nullptr || hr == nullptr || out == nullptr || tape == nullptr || workspace == nullptr || B <= 0)
    return HSENET_ERR_ARG;
  if (dropout != nullptr && !(dropout->p_attn >= 0.f && dropout->p_attn < 1.f && dropout->p_out >= 0.f && dropout->p_out < 1.f))
    return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision == HSENET_PREC_BF16)
    return packer_forward_train<__nv_bfloat16>(w, static_cast<const __nv_bfloat16*>(hr), B,
                                               static_cast<__nv_bfloat16*>(out), tape, tape_bytes, workspace,
                                               workspace_bytes, dropout, st);
  if (precision == HSENET_PREC_FP32_VERIFY)
    return packer_forward_train<float>(w, static_cast<const float*>(hr), B, static_cast<float*>(out), tape, tape_bytes,
                                       workspace, workspace_bytes, dropout, st);
  return HSENET_ERR_ARG;
}

int hsenet_packer_backward(const hsenet_packer_weights* w, const hsenet_packer_weights_t* wt, const void* hr, int B,
                           int precision, const void* d_out, const void* tape, size_t tape_bytes,
                           const hsenet_packer_grads* grads, float* d_hr, void* workspace, size_t workspace_bytes,
                           const hsenet_dropout* dropout, hsenet_stream_t stream) {
  if (w == nullptr || wt == nullptr || hr == nullptr || d_out == nullptr || tape == nullptr || grads == nullptr ||
      workspace == nullptr || B <= 0)
    return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* tp = const_cast<void*>(tape);
  if (precision == HSENET_PREC_BF16)
    return packer_backward<__nv_bfloat16>(w, wt, static_cast<const __nv_bfloat16*>(hr), B,
                                          static_cast<const __nv_bfloat16*>(d_out), tp, tape_bytes, grads, d_hr,
                                          workspace, workspace_bytes, dropout, st);
  if (precision == HSENET_PREC_FP32_VERIFY)
    return packer_backward<float>(w, wt, static_cast<const float*>(hr), B, static_cast<const float*>(d_out), tp,
                                  tape_bytes, grads, d_hr, workspace, workspace_bytes, dropout, st);
  return HSENET_ERR_ARG;
}

int hsenet_dropout_mask(float p, unsigned long long seed, long long n, float* out, hsenet_stream_t stream) {
  if (out == nullptr || n < 0 || !(p >= 0.f && p < 1.f)) return HSENET_ERR_ARG;
  DropSpec d = make_dropspec(p, seed);
  if (!d.on()) { d.keep_thresh = 0xffffffffu; d.scale = 1.0f; }      // p = 0: all ones
  return dropout_mask(out, static_cast<long>(n), d, static_cast<cudaStream_t>(stream));
}

int hsenet_self_attention_ws(const void* qkv, void* out, float* lse, float* scratch, int B, int S, int precision,
                             hsenet_stream_t stream) {
  if (qkv == nullptr || out == nullptr) return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision == HSENET_PREC_BF16)
    return attention_bf16(static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), lse, scratch, B, S, st);
  if (precision == HSENET_PREC_FP32_VERIFY)
    return attention_f32(static_cast<const float*>(qkv), static_cast<float*>(out), lse, B, S, st);
  return HSENET_ERR_ARG;
}

int hsenet_self_attention_train(const void* qkv, void* out, float* lse, int B, int S, int precision,
                                hsenet_stream_t stream) {
  if (qkv == nullptr || out == nullptr || lse == nullptr) return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision == HSENET_PREC_BF16)
    return attention_bf16(static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), lse, nullptr, B, S, st);
  if (precision == HSENET_PREC_FP32_VERIFY)
    return attention_f32(static_cast<const float*>(qkv), static_cast<float*>(out), lse, B, S, st);
  return HSENET_ERR_ARG;
}

int hsenet_self_attention_backward(const void* qkv, const void* out, const void* d_out, const float* lse, float* dvec,
                                   void* d_qkv, int B, int S, int precision, hsenet_stream_t stream) {
  if (qkv == nullptr || out == nullptr || d_out == nullptr || lse == nullptr || dvec == nullptr || d_qkv == nullptr)
    return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision == HSENET_PREC_BF16)
    return attention_bwd_bf16(static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(out),
                              static_cast<const __nv_bfloat16*>(d_out), lse, dvec, static_cast<__nv_bfloat16*>(d_qkv), B,
                              S, st);
  if (precision == HSENET_PREC_FP32_VERIFY)
    return attention_bwd_f32(static_cast<const float*>(qkv), static_cast<const float*>(out),
                             static_cast<const float*>(d_out), lse, dvec, static_cast<float*>(d_qkv), B, S, st);
  return HSENET_ERR_ARG;
}

}  // extern "C"
