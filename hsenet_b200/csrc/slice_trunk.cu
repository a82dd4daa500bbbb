// Online 2D-slice branch: slice resize + ViT-B/16 trunk -> [B,32,768] slice features, on the
// kernels of the 3D towers.  Spec: ViT4LLM_v3_med2e3.forward (vit.py:805-808) and the offline extractor
// Data/data_processing/CT-RATE/CT-RATE_2D_to_npy_file.py:75-98 (model.visual.trunk of BiomedCLIP = timm
// vit_base_patch16_224, num_classes = 0).  See include/hsenet_b200.h (hsenet_trunk_weights) for the contract.
#include "common.cuh"
#include "composite.cuh"
#include "kernels.h"

namespace hs {
namespace {

constexpr int kTrunkSeq = 197;       // cls + 14 x 14 patches
constexpr int kTrunkPatches = 196;
constexpr int kTrunkPatchDim = 256;  // 16 x 16 pixels (three identical channels folded into the weight)

template <typename T>
struct TrunkWs {
  float* X;     // [M,768]  fp32 residual stream, M = B*32*197
  T* XN;        // [M,768]
  T* QKV;       // [M,2304]
  T* ATT;       // [M,768]
  T* H;         // [M,3072]  (also the patch matrix [B*32*196,256])
  float* KMAX;  // [B*32*12] key-norm scratch of the max-free attention softmax
  size_t total;
  TrunkWs(void* base, int B) {
    const size_t M = static_cast<size_t>(B) * kNSlice * kTrunkSeq;
    Bump b(base);
    X = b.take<float>(M * kHidden);
    XN = b.take<T>(M * kHidden);
    QKV = b.take<T>(M * 3 * kHidden);
    ATT = b.take<T>(M * kHidden);
    H = b.take<T>(M * kMlp);
    KMAX = b.take<float>(static_cast<size_t>(B) * kNSlice * kHeads);
    total = b.off;
  }
};

template <typename T>
int slice_trunk_forward(const hsenet_trunk_weights* w, const float* images, int B, float* out, void* workspace,
                        size_t workspace_bytes, cudaStream_t st) {
  TrunkWs<T> ws(workspace, B);
  if (workspace_bytes < ws.total) return HS_ERR_ARG;
  const int NS = B * kNSlice, M = NS * kTrunkSeq, Mp = NS * kTrunkPatches;
  T* P = ws.H;
  HS_TRY(slice_patches<T>(images, P, B, st));
  {
    GemmEpilogue ep;   // stem conv as a GEMM + bias + positional embedding, rows written behind each slice's cls row
    ep.bias = w->b_patch;
    ep.row_add = w->pos_patch;
    ep.rows_per_group = kTrunkPatches; ep.group_stride = kTrunkSeq; ep.group_offset = 1;
    ep.out_f32 = ws.X; ep.ld_f32 = kHidden;
    HS_TRY(Prec<T>::gemm(P, kTrunkPatchDim, w->w_patch_sum, kTrunkPatchDim, Mp, kHidden, kTrunkPatchDim, ep, st));
  }
  HS_TRY(write_cls_rows(ws.X, w->cls_pos0, NS, kTrunkSeq, st));
  for (int l = 0; l < w->num_layers; ++l) {
    const hsenet_block_weights& bw = w->blocks_host[l];
    HS_TRY(layernorm_rows<T>(ws.X, kHidden, bw.ln1_g, bw.ln1_b, M, ws.XN, kHidden, nullptr, kTrunkSeq, st, w->ln_eps));
    {
      GemmEpilogue ep;
      ep.bias = bw.b_qkv;
      set_act_out(ep, ws.QKV, 3 * kHidden);
      HS_TRY(Prec<T>::gemm(ws.XN, kHidden, bw.w_qkv, kHidden, M, 3 * kHidden, kHidden, ep, st));
    }
    HS_TRY(Prec<T>::attention(ws.QKV, ws.ATT, nullptr, ws.KMAX, NS, kTrunkSeq, st));
    {
      GemmEpilogue ep;
      ep.bias = bw.b_out; ep.resid = ws.X; ep.ld_resid = kHidden; ep.out_f32 = ws.X; ep.ld_f32 = kHidden;
      HS_TRY(Prec<T>::gemm(ws.ATT, kHidden, bw.w_out, kHidden, M, kHidden, kHidden, ep, st));
    }
    HS_TRY(layernorm_rows<T>(ws.X, kHidden, bw.ln2_g, bw.ln2_b, M, ws.XN, kHidden, nullptr, kTrunkSeq, st, w->ln_eps));
    {
      GemmEpilogue ep;
      ep.bias = bw.b_fc1; ep.gelu = 1;
      set_act_out(ep, ws.H, kMlp);
      HS_TRY(Prec<T>::gemm(ws.XN, kHidden, bw.w_fc1, kHidden, M, kMlp, kHidden, ep, st));
    }
    {
      GemmEpilogue ep;
      ep.bias = bw.b_fc2; ep.resid = ws.X; ep.ld_resid = kHidden; ep.out_f32 = ws.X; ep.ld_f32 = kHidden;
      HS_TRY(Prec<T>::gemm(ws.H, kMlp, bw.w_fc2, kMlp, M, kHidden, kMlp, ep, st));
    }
  }
  // global_pool = 'token': final norm of the cls row of every slice only
  return layernorm_rows<float>(ws.X, static_cast<long>(kTrunkSeq) * kHidden, w->norm_g, w->norm_b, NS, out, kHidden, nullptr, 1,
                               st, w->ln_eps);
}

}  // namespace
}  // namespace hs

using namespace hs;

extern "C" {

size_t hsenet_slice_trunk_workspace_bytes(int B, int precision) {
  if (B <= 0) return 0;
  if (precision == HSENET_PREC_BF16) return TrunkWs<__nv_bfloat16>(nullptr, B).total;
  if (precision == HSENET_PREC_FP32_VERIFY) return TrunkWs<float>(nullptr, B).total;
  return 0;
}

int hsenet_slice_trunk_forward(const hsenet_trunk_weights* w, const float* images, int B, int precision, float* out,
                               void* workspace, size_t workspace_bytes, hsenet_stream_t stream) {
  if (w == nullptr || images == nullptr || out == nullptr || workspace == nullptr || w->blocks_host == nullptr || B <= 0 ||
      w->num_layers < 0)
    return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision == HSENET_PREC_BF16)
    return slice_trunk_forward<__nv_bfloat16>(w, images, B, out, workspace, workspace_bytes, st);
  if (precision == HSENET_PREC_FP32_VERIFY)
    return slice_trunk_forward<float>(w, images, B, out, workspace, workspace_bytes, st);
  return HSENET_ERR_ARG;
}

}  // extern "C"
