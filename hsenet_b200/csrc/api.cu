// C ABI of libhsenet_sm100a.so (declared in include/hsenet_b200.h) and the composite forward passes built from the
// operator kernels.  No allocation, no synchronisation, no CPU compute: everything is enqueued on the caller's stream.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "composite.cuh"
#include "kernels.h"

#include <cstdlib>
#include <type_traits>

namespace hs {

static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- optional per-kernel-class event profiling (bench.py's roofline pass) ------------------------------------------
struct ProfRec { cudaEvent_t a, b; int cls; double flops, bytes; };
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;
static std::atomic<bool> g_prof_on{false};

ProfScope::ProfScope(int cls, double flops, double bytes, cudaStream_t st) : st_(st), idx_(-1) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRec r{nullptr, nullptr, cls, flops, bytes};
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
  idx_ = static_cast<long>(g_prof.size()) - 1;
}
ProfScope::~ProfScope() {
  if (idx_ < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[idx_].b, st_);
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("HSENET_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

bool first_use_on_device(unsigned char* slot) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return true;
  if (slot[dev]) return false;
  slot[dev] = 1;
  return true;
}

int num_sms() {
  static int cache[kMaxDevices] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  if (cache[dev] == 0) {
    int v = 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = v > 0 ? v : 148;
  }
  return cache[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 2D bf16 tensor map, 128-byte swizzle: dim0 = `inner` contiguous elements, dim1 = `outer` rows `ld_elems` apart.
int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                      uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return HS_ERR_DRIVER;
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {ld_elems * 2};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? HS_OK : HS_ERR_DRIVER;
}

namespace {

constexpr int kStatSlots = kHidden / 128;   // column slabs of a residual row: one partial-statistics slot each
constexpr float kLnEps = 1e-5f;        // nn.LayerNorm default (same constant as layernorm_rows)

// ---- ViT workspace layout -----------------------------------------------------------------------------------
template <typename T>
struct VitWs {
  float* X;      // [M,768]   fp32 residual stream
  T* XN;         // [M,768]   LayerNorm output / stage-2 bf16 patch embedding
  T* QKV;        // [M,2304]  (stage 2: aliased by Q fp32 [Mp,768])
  T* ATT;        // [M,768]   (stage 2: aliased by O [Mp,768])
  T* H;          // [M,3072]  (aliased by the im2col patches [Mp,1024] and, stage 2, XP fp32 [Mp,768])
  T* S16;        // [B*32,768]   slice features as act
  float* SKV;    // [B*32,1536]  Wk|Wv of the slice features
  float2* STATS; // [2, 6, M]  per-128-column partial (sum, sum of squares) of the residual rows entering norm1 / norm2, for
                 //   the LayerNorms folded into GEMMs (rewritten by every producing GEMM: no clearing, no atomics)
  float* KMAX;   // [B*12]  largest squared key norm per (volume, head): scratch of the max-free attention softmax
  size_t stats_bytes;
  size_t total;
  VitWs(void* base, int B) {
    const size_t M = static_cast<size_t>(B) * kSeq;
    Bump b(base);
    X = b.take<float>(M * kHidden);
    XN = b.take<T>(M * kHidden);
    QKV = b.take<T>(M * 3 * kHidden);
    ATT = b.take<T>(M * kHidden);
    H = b.take<T>(M * kMlp);
    S16 = b.take<T>(static_cast<size_t>(B) * kNSlice * kHidden);
    SKV = b.take<float>(static_cast<size_t>(B) * kNSlice * 2 * kHidden);
    stats_bytes = 2 * static_cast<size_t>(kStatSlots) * M * sizeof(float2);
    STATS = b.take<float2>(2 * static_cast<size_t>(kStatSlots) * M);
    KMAX = b.take<float>(static_cast<size_t>(B) * kHeads);
    total = b.off;
  }
};

// stage 2 keeps the fp32 patch embedding behind the (former) patch-matrix region of H
template <typename T>
float* stage2_xp(const VitWs<T>& ws, int Mp) {
  return reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws.H) + align_up(static_cast<size_t>(Mp) * kPatchDim * sizeof(T)));
}
// Epilogue of the patch-embedding GEMM of one tower: bias + positional embedding; stage 1 writes the rows behind the cls row
// of the residual stream (fuses torch.cat, vit.py:459-461), stage 2 into the fp32 + activation-dtype patch buffers.
template <typename T>
GemmEpilogue patch_epilogue(const hsenet_vit_weights* w, const VitWs<T>& ws, int B) {
  GemmEpilogue ep;
  ep.bias = w->b_patch;
  ep.row_add = w->pos_embed;
  ep.rows_per_group = kNPatch;
  if (w->stage == 1) {
    ep.group_stride = kSeq; ep.group_offset = 1;
    ep.out_f32 = ws.X; ep.ld_f32 = kHidden;
  } else {
    ep.group_stride = kNPatch; ep.group_offset = 0;
    ep.out_f32 = stage2_xp<T>(ws, B * kNPatch); ep.ld_f32 = kHidden;
    set_act_out(ep, ws.XN, kHidden);
  }
  return ep;
}

template <typename T>
int vit_forward(const hsenet_vit_weights* w, const float* images, const float* images_2d, int B, T* out_tokens,
                T* out_patch, float* hidden, float* scores, void* workspace, size_t workspace_bytes, int flags,
                cudaStream_t st) {
  VitWs<T> ws(workspace, B);
  if (workspace_bytes < ws.total) return HS_ERR_ARG;
  const int M = B * kSeq, Mp = B * kNPatch;
  T* P = ws.H;                                                            // im2col patches [Mp,1024]

  // K1: patch embedding with fused bias + positional embedding (+ cls-offset row remap).  bf16 mode: implicit-im2col
  // tf32 GEMM straight from the volume (patch_embed_tcgen05.cu) unless the caller already ran it for both towers at once
  // (hsenet_patch_embed_dual, flags & 1); verification mode: explicit im2col + fp32 GEMM.
  const GemmEpilogue pe = patch_epilogue<T>(w, ws, B);
  if (!(flags & HSENET_VIT_PATCH_DONE)) {
    if (std::is_same<T, __nv_bfloat16>::value && w->w_patch_f32 != nullptr) {
      HS_TRY(patch_embed_tf32(images, w->w_patch_f32, B, 1, pe, pe, st));
    } else {
      HS_TRY(im2col_patches<T>(images, B, P, st));
      HS_TRY(Prec<T>::gemm(P, kPatchDim, w->w_patch, kPatchDim, Mp, kHidden, kPatchDim, pe, st));
    }
  }
  if (w->stage == 2) {
    if (images_2d == nullptr) return HS_ERR_ARG;
    float* XP = stage2_xp<T>(ws, Mp);
    T* XPa = ws.XN;
    float* Q = reinterpret_cast<float*>(ws.QKV);
    T* O = ws.ATT;
    // K10: regular_attention(x, slices, slices)  (vit.py:50-64)
    HS_TRY(cast_rows<T>(images_2d, ws.S16, static_cast<long>(B) * kNSlice * kHidden, st));
    {
      GemmEpilogue ep;   // Wk | Wv of the 32 slice features
      ep.bias = w->b_skv; ep.out_f32 = ws.SKV; ep.ld_f32 = 2 * kHidden;
      HS_TRY(Prec<T>::gemm(ws.S16, kHidden, w->w_skv, kHidden, B * kNSlice, 2 * kHidden, kHidden, ep, st));
    }
    {
      GemmEpilogue ep;   // query_list = Wq(x)
      ep.bias = w->b_sq; ep.out_f32 = Q; ep.ld_f32 = kHidden;
      HS_TRY(Prec<T>::gemm(XPa, kHidden, w->w_sq, kHidden, Mp, kHidden, kHidden, ep, st));
    }
    HS_TRY(slice_cross_attention<T>(Q, ws.SKV, O, nullptr, B, st));
    {
      GemmEpilogue ep;   // query_list + output_linear(x)   (residual is the PROJECTED query, vit.py:62)
      ep.bias = w->b_so; ep.resid = Q; ep.ld_resid = kHidden; ep.out_f32 = Q; ep.ld_f32 = kHidden;
      HS_TRY(Prec<T>::gemm(O, kHidden, w->w_so, kHidden, Mp, kHidden, kHidden, ep, st));
    }
    // K11: LayerNorm -> patch_score_proj -> sigmoid -> x * score, written behind the cls row (vit.py:338-349)
    HS_TRY(score_and_scale(Q, w->sn_g, w->sn_b, w->w_score, w->b_score, XP, ws.X, scores, B, st));
  }
  HS_TRY(write_cls_rows(ws.X, w->cls_token, B, kSeq, st));

  // 12 x MONAI TransformerBlock (vit.py:463-466): x += attn(norm1(x)); x += mlp(norm2(x))
  // bf16 path with gain-folded weights supplied: norm2 of every layer and norm1 of layers >= 1 are folded into the
  // GEMMs around them (gemm_epilogue.cuh) -- the GEMM that produces x also writes bf16(x) into XN and the row
  // statistics, the consuming GEMM runs on XN with W' = gamma (.) W and normalises in its epilogue.
  constexpr bool kCanFold = std::is_same<T, __nv_bfloat16>::value;
  auto folds_ln1 = [&](int l) {
    return kCanFold && l >= 1 && l < w->num_layers && w->blocks_host[l].w_qkv_ln != nullptr &&
           w->blocks_host[l].cs_qkv != nullptr && w->blocks_host[l].b_qkv_ln != nullptr;   // selected by the caller
  };
  auto folds_ln2 = [&](int l) {
    return kCanFold && w->blocks_host[l].w_fc1_ln != nullptr &&
           w->blocks_host[l].cs_fc1 != nullptr && w->blocks_host[l].b_fc1_ln != nullptr;
  };
  float2* const stats1 = ws.STATS;                                        // rows entering norm1 of the next layer
  float2* const stats2 = ws.STATS + static_cast<size_t>(kStatSlots) * M;  // rows entering norm2 of this layer
  for (int l = 0; l < w->num_layers; ++l) {
    const hsenet_block_weights& bw = w->blocks_host[l];
    {
      GemmEpilogue ep;   // qkv, no bias
      set_act_out(ep, ws.QKV, 3 * kHidden);
      if (folds_ln1(l)) {
        ep.bias = bw.b_qkv_ln; ep.colsum = bw.cs_qkv; ep.stats_in = stats1; ep.ld_stats = M; ep.stats_slots = kStatSlots;
        ep.ln_inv_dim = 1.0f / kHidden; ep.ln_eps = kLnEps;
        HS_TRY(Prec<T>::gemm(ws.XN, kHidden, bw.w_qkv_ln, kHidden, M, 3 * kHidden, kHidden, ep, st));
      } else {
        ep.bias = bw.b_qkv;                 // NULL for the MONAI blocks (qkv_bias=False)
        HS_TRY(layernorm_rows<T>(ws.X, kHidden, bw.ln1_g, bw.ln1_b, M, ws.XN, kHidden, nullptr, kSeq, st));
        HS_TRY(Prec<T>::gemm(ws.XN, kHidden, bw.w_qkv, kHidden, M, 3 * kHidden, kHidden, ep, st));
      }
    }
    HS_TRY(Prec<T>::attention(ws.QKV, ws.ATT, nullptr, ws.KMAX, B, kSeq, st));
    {
      GemmEpilogue ep;   // out_proj + residual
      ep.bias = bw.b_out; ep.resid = ws.X; ep.ld_resid = kHidden; ep.out_f32 = ws.X; ep.ld_f32 = kHidden;
      if (folds_ln2(l)) {
        set_act_out(ep, ws.XN, kHidden);
        ep.stats_out = stats2; ep.ld_stats = M;
      }
      HS_TRY(Prec<T>::gemm(ws.ATT, kHidden, bw.w_out, kHidden, M, kHidden, kHidden, ep, st));
    }
    {
      GemmEpilogue ep;   // linear1 + exact GELU
      ep.gelu = 1;
      set_act_out(ep, ws.H, kMlp);
      if (folds_ln2(l)) {
        ep.bias = bw.b_fc1_ln; ep.colsum = bw.cs_fc1; ep.stats_in = stats2; ep.ld_stats = M; ep.stats_slots = kStatSlots;
        ep.ln_inv_dim = 1.0f / kHidden; ep.ln_eps = kLnEps;
        HS_TRY(Prec<T>::gemm(ws.XN, kHidden, bw.w_fc1_ln, kHidden, M, kMlp, kHidden, ep, st));
      } else {
        ep.bias = bw.b_fc1;
        HS_TRY(layernorm_rows<T>(ws.X, kHidden, bw.ln2_g, bw.ln2_b, M, ws.XN, kHidden, nullptr, kSeq, st));
        HS_TRY(Prec<T>::gemm(ws.XN, kHidden, bw.w_fc1, kHidden, M, kMlp, kHidden, ep, st));
      }
    }
    {
      GemmEpilogue ep;   // linear2 + residual
      ep.bias = bw.b_fc2; ep.resid = ws.X; ep.ld_resid = kHidden; ep.out_f32 = ws.X; ep.ld_f32 = kHidden;
      if (folds_ln1(l + 1)) {
        set_act_out(ep, ws.XN, kHidden);
        ep.stats_out = stats1; ep.ld_stats = M;
      }
      HS_TRY(Prec<T>::gemm(ws.H, kMlp, bw.w_fc2, kMlp, M, kHidden, kMlp, ep, st));
    }
    if (hidden != nullptr) {
      if (cudaMemcpyAsync(hidden + static_cast<size_t>(l) * M * kHidden, ws.X,
                          static_cast<size_t>(M) * kHidden * sizeof(float), cudaMemcpyDeviceToDevice,
                          st) != cudaSuccess)
        return HS_ERR_CUDA;
    }
  }
  // final LayerNorm (vit.py:467); the patch-only copy is the `[:, 1:]` view of ViT3DTower_dual_encoders (vit.py:934-936)
  if (out_tokens != nullptr || out_patch != nullptr)
    HS_TRY(layernorm_rows<T>(ws.X, kHidden, w->norm_g, w->norm_b, M, out_tokens, kHidden, out_patch, kSeq, st));
  return HS_OK;
}

// ---- packer workspace -----------------------------------------------------------------------------------------
template <typename T>
struct PackerWs {
  T* LR;       // [B*128,768]
  T* KV;       // [B*2048,1536]
  float* Q;    // [B*128,768]
  T* O;        // [B*128,768]
  T* A;        // [B*128,768]
  T* H;        // [B*128,out_dim]
  size_t total;
  PackerWs(void* base, int B, int out_dim) {
    const size_t n = static_cast<size_t>(B) * 128;
    Bump b(base);
    LR = b.take<T>(n * kHidden);
    KV = b.take<T>(static_cast<size_t>(B) * kNPatch * 2 * kHidden);
    Q = b.take<float>(n * kHidden);
    O = b.take<T>(n * kHidden);
    A = b.take<T>(n * kHidden);
    H = b.take<T>(n * out_dim);
    total = b.off;
  }
};

template <typename T>
int packer_forward(const hsenet_packer_weights* w, const T* hr, int B, void* out, int out_dtype,
                   int out_tokens_per_batch, int token_offset, void* workspace, size_t workspace_bytes,
                   cudaStream_t st) {
  const int D = w->out_dim;
  if (D <= 0 || D % 256 != 0) return HS_ERR_SHAPE;
  if (token_offset < 0 || token_offset + 128 > out_tokens_per_batch) return HS_ERR_ARG;
  PackerWs<T> ws(workspace, B, D);
  if (workspace_bytes < ws.total) return HS_ERR_ARG;
  const int Mp = B * kNPatch, Mw = B * 128;
  HS_TRY(packer_pool<T>(hr, ws.LR, B, st));                                  // K13
  {
    GemmEpilogue ep;   // K14: Wk | Wv over the HR tokens in natural order (window regrouping is pure indexing)
    ep.bias = w->b_kv; set_act_out(ep, ws.KV, 2 * kHidden);
    HS_TRY(Prec<T>::gemm(hr, kHidden, w->w_kv, kHidden, Mp, 2 * kHidden, kHidden, ep, st));
  }
  {
    GemmEpilogue ep;   // Wq(LR)
    ep.bias = w->b_q; ep.out_f32 = ws.Q; ep.ld_f32 = kHidden;
    HS_TRY(Prec<T>::gemm(ws.LR, kHidden, w->w_q, kHidden, Mw, kHidden, kHidden, ep, st));
  }
  HS_TRY(packer_window_attention<T>(ws.Q, ws.KV, ws.O, B, st));            // K15
  {
    GemmEpilogue ep;   // K16: Wq(LR) + output_linear(x)
    ep.bias = w->b_o; ep.resid = ws.Q; ep.ld_resid = kHidden; ep.out_f32 = ws.Q; ep.ld_f32 = kHidden;
    HS_TRY(Prec<T>::gemm(ws.O, kHidden, w->w_o, kHidden, Mw, kHidden, kHidden, ep, st));
  }
  HS_TRY(layernorm_rows<T>(ws.Q, kHidden, w->ln_g, w->ln_b, Mw, ws.A, kHidden, nullptr, 128, st));
  {
    GemmEpilogue ep;   // K17: proj_mpls.0 + GELU
    ep.bias = w->b_p0; ep.gelu = 1; set_act_out(ep, ws.H, D);
    HS_TRY(Prec<T>::gemm(ws.A, kHidden, w->w_p0, kHidden, Mw, D, kHidden, ep, st));
  }
  {
    GemmEpilogue ep;   // proj_mpls.2, written straight into this packer's slot of [B, tokens, D]
    ep.bias = w->b_p2;
    ep.rows_per_group = 128; ep.group_stride = out_tokens_per_batch; ep.group_offset = token_offset;
    if (out_dtype == HSENET_DTYPE_F32) {
      ep.out_f32 = static_cast<float*>(out); ep.ld_f32 = D;
    } else if (out_dtype == HSENET_DTYPE_BF16 && sizeof(T) == 2) {
      ep.out_bf16 = static_cast<__nv_bfloat16*>(out); ep.ld_bf16 = D;
    } else {
      return HS_ERR_ARG;
    }
    HS_TRY(Prec<T>::gemm(ws.H, D, w->w_p2, D, Mw, D, D, ep, st));
  }
  return HS_OK;
}

}  // namespace
}  // namespace hs

// =================================================================================================================
// C ABI
// =================================================================================================================
using namespace hs;

extern "C" {

const char* hsenet_version(void) { return "hsenet_b200 0.1 (sm_100a)"; }

const char* hsenet_error_string(int code) {
  switch (code) {
    case HSENET_OK: return "ok";
    case HSENET_ERR_SHAPE: return "unsupported shape";
    case HSENET_ERR_ALIGN: return "misaligned pointer or leading dimension";
    case HSENET_ERR_CUDA: return "CUDA launch failure";
    case HSENET_ERR_ARG: return "bad argument (null pointer, enum, or workspace too small)";
    case HSENET_ERR_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    default: return "unknown error";
  }
}

uint64_t hsenet_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void hsenet_profile_start(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_prof_on.store(true);
}

int hsenet_profile_stop(double* ms, double* flops, double* bytes, uint64_t* launches) {
  g_prof_on.store(false);
  if (cudaDeviceSynchronize() != cudaSuccess) return HSENET_ERR_CUDA;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int c = 0; c < HSENET_PROFILE_CLASSES; ++c) { ms[c] = 0; flops[c] = 0; bytes[c] = 0; launches[c] = 0; }
  for (auto& r : g_prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.cls >= 0 && r.cls < HSENET_PROFILE_CLASSES) {
      ms[r.cls] += t; flops[r.cls] += r.flops; bytes[r.cls] += r.bytes; launches[r.cls] += 1;
    }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof.clear();
  return HSENET_OK;
}

size_t hsenet_vit_workspace_bytes(int B, int precision, int stage) {
  (void)stage;
  if (B <= 0) return 0;
  if (precision == HSENET_PREC_BF16) return VitWs<__nv_bfloat16>(nullptr, B).total;
  if (precision == HSENET_PREC_FP32_VERIFY) return VitWs<float>(nullptr, B).total;
  return 0;
}

int hsenet_vit_forward(const hsenet_vit_weights* w, const float* images, const float* images_2d, int B,
                       int precision, void* out_tokens, void* out_patch, float* hidden_f32, float* scores_f32,
                       void* workspace, size_t workspace_bytes, int flags, hsenet_stream_t stream) {
  if (w == nullptr || images == nullptr || workspace == nullptr || w->blocks_host == nullptr) return HSENET_ERR_ARG;
  if (B <= 0 || w->num_layers < 0 || (w->stage != 1 && w->stage != 2)) return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision == HSENET_PREC_BF16)
    return vit_forward<__nv_bfloat16>(w, images, images_2d, B, static_cast<__nv_bfloat16*>(out_tokens),
                                      static_cast<__nv_bfloat16*>(out_patch), hidden_f32, scores_f32, workspace,
                                      workspace_bytes, flags, st);
  if (precision == HSENET_PREC_FP32_VERIFY)
    return vit_forward<float>(w, images, images_2d, B, static_cast<float*>(out_tokens),
                              static_cast<float*>(out_patch), hidden_f32, scores_f32, workspace, workspace_bytes,
                              flags & ~HSENET_VIT_PATCH_DONE, st);
  return HSENET_ERR_ARG;
}

int hsenet_patch_embed_dual(const hsenet_vit_weights* w1, const hsenet_vit_weights* w2, const float* w_stack_f32,
                            const float* images, int B, void* workspace1, size_t workspace1_bytes, void* workspace2,
                            size_t workspace2_bytes, hsenet_stream_t stream) {
  if (w1 == nullptr || w2 == nullptr || w_stack_f32 == nullptr || images == nullptr || workspace1 == nullptr ||
      workspace2 == nullptr || B <= 0)
    return HSENET_ERR_ARG;
  if (w1->stage != 1 || w2->stage != 2) return HSENET_ERR_ARG;
  VitWs<__nv_bfloat16> ws1(workspace1, B), ws2(workspace2, B);
  if (workspace1_bytes < ws1.total || workspace2_bytes < ws2.total) return HSENET_ERR_ARG;
  return patch_embed_tf32(images, w_stack_f32, B, 2, patch_epilogue<__nv_bfloat16>(w1, ws1, B),
                          patch_epilogue<__nv_bfloat16>(w2, ws2, B), static_cast<cudaStream_t>(stream));
}

size_t hsenet_packer_workspace_bytes(int B, int precision, int out_dim) {
  if (B <= 0 || out_dim <= 0) return 0;
  if (precision == HSENET_PREC_BF16) return PackerWs<__nv_bfloat16>(nullptr, B, out_dim).total;
  if (precision == HSENET_PREC_FP32_VERIFY) return PackerWs<float>(nullptr, B, out_dim).total;
  return 0;
}

int hsenet_packer_forward(const hsenet_packer_weights* w, const void* hr, int B, int precision, void* out,
                          int out_dtype, int out_tokens_per_batch, int token_offset, void* workspace,
                          size_t workspace_bytes, hsenet_stream_t stream) {
  if (w == nullptr || hr == nullptr || out == nullptr || workspace == nullptr || B <= 0) return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision == HSENET_PREC_BF16)
    return packer_forward<__nv_bfloat16>(w, static_cast<const __nv_bfloat16*>(hr), B, out, out_dtype,
                                         out_tokens_per_batch, token_offset, workspace, workspace_bytes, st);
  if (precision == HSENET_PREC_FP32_VERIFY)
    return packer_forward<float>(w, static_cast<const float*>(hr), B, out, out_dtype, out_tokens_per_batch,
                                 token_offset, workspace, workspace_bytes, st);
  return HSENET_ERR_ARG;
}

int hsenet_clip_image_head(const void* tokens, const void* w_proj, const float* b_proj, int B, int precision,
                           float* out, void* workspace, size_t workspace_bytes, hsenet_stream_t stream) {
  if (tokens == nullptr || w_proj == nullptr || out == nullptr || workspace == nullptr || B <= 0)
    return HSENET_ERR_ARG;
  if (workspace_bytes < static_cast<size_t>(B) * kHidden * sizeof(float)) return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* tmp = static_cast<float*>(workspace);
  GemmEpilogue ep;
  ep.bias = b_proj; ep.out_f32 = tmp; ep.ld_f32 = kHidden;
  int rc;
  if (precision == HSENET_PREC_BF16)
    rc = gemm_bf16(tokens, kSeq * kHidden, w_proj, kHidden, B, kHidden, kHidden, ep, st);
  else if (precision == HSENET_PREC_FP32_VERIFY)
    rc = gemm_f32(static_cast<const float*>(tokens), kSeq * kHidden, static_cast<const float*>(w_proj), kHidden, B,
                  kHidden, kHidden, ep, st);
  else
    return HSENET_ERR_ARG;
  if (rc != HS_OK) return rc;
  return l2_normalize_rows(tmp, out, B, kHidden, st);
}

int hsenet_linear(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias,
                  const float* resid, int ld_resid, int gelu, float* out_f32, int ld_f32, void* out_act, int ld_act,
                  int precision, hsenet_stream_t stream) {
  if (A == nullptr || W == nullptr || (out_f32 == nullptr && out_act == nullptr)) return HSENET_ERR_ARG;
  GemmEpilogue ep;
  ep.bias = bias; ep.resid = resid; ep.ld_resid = ld_resid; ep.gelu = gelu;
  ep.out_f32 = out_f32; ep.ld_f32 = ld_f32;
  ep.out_bf16 = static_cast<__nv_bfloat16*>(out_act); ep.ld_bf16 = ld_act;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision == HSENET_PREC_BF16) return gemm_bf16(A, lda, W, ldw, M, N, K, ep, st);
  if (precision == HSENET_PREC_FP32_VERIFY)
    return gemm_f32(static_cast<const float*>(A), lda, static_cast<const float*>(W), ldw, M, N, K, ep, st);
  return HSENET_ERR_ARG;
}

int hsenet_self_attention(const void* qkv, void* out, int B, int S, int precision, hsenet_stream_t stream) {
  if (qkv == nullptr || out == nullptr) return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (precision == HSENET_PREC_BF16)
    return attention_bf16(static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), nullptr, nullptr, B, S, st);
  if (precision == HSENET_PREC_FP32_VERIFY)
    return attention_f32(static_cast<const float*>(qkv), static_cast<float*>(out), nullptr, B, S, st);
  return HSENET_ERR_ARG;
}

int hsenet_layernorm(const float* x, const float* gamma, const float* beta, long rows, void* out, int out_dtype,
                     hsenet_stream_t stream) {
  if (x == nullptr || gamma == nullptr || beta == nullptr || out == nullptr) return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (out_dtype == HSENET_DTYPE_F32)
    return layernorm_rows<float>(x, kHidden, gamma, beta, rows, static_cast<float*>(out), kHidden, nullptr, 1, st);
  if (out_dtype == HSENET_DTYPE_BF16)
    return layernorm_rows<__nv_bfloat16>(x, kHidden, gamma, beta, rows, static_cast<__nv_bfloat16*>(out), kHidden,
                                         nullptr, 1, st);
  return HSENET_ERR_ARG;
}

int hsenet_patch_im2col(const float* images, int B, void* out, int out_dtype, hsenet_stream_t stream) {
  if (images == nullptr || out == nullptr) return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (out_dtype == HSENET_DTYPE_F32) return im2col_patches<float>(images, B, static_cast<float*>(out), st);
  if (out_dtype == HSENET_DTYPE_BF16)
    return im2col_patches<__nv_bfloat16>(images, B, static_cast<__nv_bfloat16*>(out), st);
  return HSENET_ERR_ARG;
}

int hsenet_packer_pool(const void* hr, void* lr, int B, int dtype, hsenet_stream_t stream) {
  if (hr == nullptr || lr == nullptr) return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == HSENET_DTYPE_F32) return packer_pool<float>(static_cast<const float*>(hr), static_cast<float*>(lr), B, st);
  if (dtype == HSENET_DTYPE_BF16)
    return packer_pool<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(hr), static_cast<__nv_bfloat16*>(lr), B, st);
  return HSENET_ERR_ARG;
}

int hsenet_packer_window_attention(const float* q, const void* kv, void* out, int B, int dtype,
                                   hsenet_stream_t stream) {
  if (q == nullptr || kv == nullptr || out == nullptr) return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == HSENET_DTYPE_F32)
    return packer_window_attention<float>(q, static_cast<const float*>(kv), static_cast<float*>(out), B, st);
  if (dtype == HSENET_DTYPE_BF16)
    return packer_window_attention<__nv_bfloat16>(q, static_cast<const __nv_bfloat16*>(kv),
                                                  static_cast<__nv_bfloat16*>(out), B, st);
  return HSENET_ERR_ARG;
}

int hsenet_slice_cross_attention(const float* q, const float* kv, void* out, float* attn, int B, int dtype,
                                 hsenet_stream_t stream) {
  if (q == nullptr || kv == nullptr || out == nullptr) return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == HSENET_DTYPE_F32) return slice_cross_attention<float>(q, kv, static_cast<float*>(out), attn, B, st);
  if (dtype == HSENET_DTYPE_BF16)
    return slice_cross_attention<__nv_bfloat16>(q, kv, static_cast<__nv_bfloat16*>(out), attn, B, st);
  return HSENET_ERR_ARG;
}

int hsenet_slice_extract(const float* images, void* out, int B, int out_h, int out_w, int out_dtype,
                         hsenet_stream_t stream) {
  if (images == nullptr || out == nullptr) return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (out_dtype == HSENET_DTYPE_F32) return slice_extract<float>(images, static_cast<float*>(out), B, out_h, out_w, st);
  if (out_dtype == HSENET_DTYPE_BF16)
    return slice_extract<__nv_bfloat16>(images, static_cast<__nv_bfloat16*>(out), B, out_h, out_w, st);
  return HSENET_ERR_ARG;
}

int hsenet_gather_rows(const void* in, int in_dtype, long batch_stride, long row_stride, int B, int rows, void* out,
                       int out_dtype, hsenet_stream_t stream) {
  if (in == nullptr || out == nullptr) return HSENET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (out_dtype == HSENET_DTYPE_F32)
    return gather_rows<float>(in, in_dtype, batch_stride, row_stride, B, rows, static_cast<float*>(out), st);
  if (out_dtype == HSENET_DTYPE_BF16)
    return gather_rows<__nv_bfloat16>(in, in_dtype, batch_stride, row_stride, B, rows,
                                      static_cast<__nv_bfloat16*>(out), st);
  return HSENET_ERR_ARG;
}

int hsenet_patch_gather_map(int32_t* out, hsenet_stream_t stream) {
  if (out == nullptr) return HSENET_ERR_ARG;
  return patch_gather_map(out, static_cast<cudaStream_t>(stream));
}
int hsenet_packer_window_map(int32_t* out, hsenet_stream_t stream) {
  if (out == nullptr) return HSENET_ERR_ARG;
  return packer_window_map(out, static_cast<cudaStream_t>(stream));
}
int hsenet_cast_bf16(const float* in, void* out, long n, hsenet_stream_t stream) {
  if (in == nullptr || out == nullptr) return HSENET_ERR_ARG;
  return cast_rows<__nv_bfloat16>(in, static_cast<__nv_bfloat16*>(out), n, static_cast<cudaStream_t>(stream));
}
int hsenet_fold_layernorm(const float* w, const float* gamma, const float* beta, const float* bias, int N, int K,
                          void* w_folded_bf16, float* colsum, float* bias_folded, hsenet_stream_t stream) {
  if (w == nullptr || gamma == nullptr || beta == nullptr || w_folded_bf16 == nullptr || colsum == nullptr ||
      bias_folded == nullptr || N < 0 || K < 0)
    return HSENET_ERR_ARG;
  return fold_layernorm(w, gamma, beta, bias, N, K, static_cast<__nv_bfloat16*>(w_folded_bf16), colsum, bias_folded,
                        static_cast<cudaStream_t>(stream));
}
int hsenet_hu_resample(const float* raw, int n0, int n1, int n2, float slope, float intercept, float hu_min,
                       float hu_max, float* out, int o0, int o1, int o2, float* scratch, hsenet_stream_t stream) {
  if (raw == nullptr || out == nullptr) return HSENET_ERR_ARG;
  return hu_resample(raw, n0, n1, n2, slope, intercept, hu_min, hu_max, out, o0, o1, o2, scratch,
                     static_cast<cudaStream_t>(stream));
}
int hsenet_minmax(const float* x, long n, float* minmax2, int32_t* scratch2, hsenet_stream_t stream) {
  if (x == nullptr || minmax2 == nullptr || scratch2 == nullptr) return HSENET_ERR_ARG;
  return minmax(x, n, minmax2, scratch2, static_cast<cudaStream_t>(stream));
}
int hsenet_foreground_bbox(const float* x, int d0, int d1, int d2, const float* minmax2, int32_t* bbox6,
                           hsenet_stream_t stream) {
  if (x == nullptr || minmax2 == nullptr || bbox6 == nullptr) return HSENET_ERR_ARG;
  return foreground_bbox(x, d0, d1, d2, minmax2, bbox6, static_cast<cudaStream_t>(stream));
}
int hsenet_crop_normalize_resize(const float* x, int d0, int d1, int d2, const float* minmax2, const int32_t* bbox6,
                                 float* out, int o0, int o1, int o2, hsenet_stream_t stream) {
  if (x == nullptr || minmax2 == nullptr || bbox6 == nullptr || out == nullptr) return HSENET_ERR_ARG;
  return crop_normalize_resize(x, d0, d1, d2, minmax2, bbox6, out, o0, o1, o2, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
