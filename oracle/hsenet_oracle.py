"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the HSENet visual-encoding hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this file; the product package (hsenet_b200/) never does.

A plain fp32 PyTorch/numpy restatement of the reference's algorithm, written as stateless functions
over ``state_dict``-style weight mappings so that the same random-init weights can be fed to the CUDA
path and to this checker.  Every function cites the reference lines it follows (paths relative to
/root/reference/Preprint/LaMed/src/model/ unless stated).  The restatement itself is pinned against the
reference's own files (run through oracle/monai_shim.py) by tests/test_oracle.py::test_restatement_*
and against committed fixtures generated from those files (tests/golden/make_golden.py).

Parity status at the third-party boundary: MONAI 1.3.0 (requirements.txt:62) is absent; its blocks are
restated from the published source -- "parity unpinned" by any reference-held golden vector (the
reference has no tests).  See oracle/monai_shim.py for the indirect pins.
"""
from __future__ import annotations

import math
from typing import Mapping, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# Geometry hard-coded by the reference (spatial_pooling_projector.py:140, lamed_arch.py:125, vit.py:194).
IMG = (32, 256, 256)
PATCH = (4, 16, 16)
GRID = (8, 16, 16)
N_PATCH = 2048
HIDDEN = 768
HEADS = 12
HEAD_DIM = 64
MLP_DIM = 3072
PATCH_DIM = 1024
N_SLICE = 32
LN_EPS = 1e-5  # torch.nn.LayerNorm default, used by MONAI TransformerBlock, vit.py:310/445, packer :60


# --------------------------------------------------------------------------------------------------
# Integer maps (must be reproduced bit-exactly by the CUDA path)
# --------------------------------------------------------------------------------------------------
def patch_gather_map() -> np.ndarray:
    """[2048, 1024] int32: flat voxel index (within one 1x32x256x256 volume) feeding patch token t,
    feature f.  Closed form of MONAI's perceptron rearrange quoted at vit.py:437
    ``'b c (h p1) (w p2) (d p3) -> b (h w d) (p1 p2 p3 c)'`` with p = (4,16,16) on [B,1,32,256,256]:
    token t = (dz*16 + hy)*16 + wx, feature f = (p1*16 + p2)*16 + p3,
    voxel = (dz*4+p1)*65536 + (hy*16+p2)*256 + (wx*16+p3)."""
    t = np.arange(N_PATCH, dtype=np.int64)[:, None]
    f = np.arange(PATCH_DIM, dtype=np.int64)[None, :]
    dz, hy, wx = t // 256, (t // 16) % 16, t % 16
    p1, p2, p3 = f // 256, (f // 16) % 16, f % 16
    vox = (dz * 4 + p1) * 65536 + (hy * 16 + p2) * 256 + (wx * 16 + p3)
    return vox.astype(np.int32)


def packer_window_map() -> np.ndarray:
    """[128, 16] int32: HR token index of member e of packer window n.
    Closed form of the reshape/permute at spatial_pooling_projector.py:70-71 with kernel (1,4,4) on the
    8x16x16 grid (:140): n = dz*16 + wy*4 + hx, e = sw*4 + sh -> dz*256 + (4*wy+sw)*16 + (4*hx+sh)."""
    n = np.arange(128, dtype=np.int64)[:, None]
    e = np.arange(16, dtype=np.int64)[None, :]
    dz, wy, hx = n // 16, (n // 4) % 4, n % 4
    sw, sh = e // 4, e % 4
    return (dz * 256 + (4 * wy + sw) * 16 + (4 * hx + sh)).astype(np.int32)


def patchify(x: torch.Tensor) -> torch.Tensor:
    """[B,1,32,256,256] -> [B,2048,1024] by the gather map above (== the einops Rearrange, vit.py:437)."""
    b = x.shape[0]
    idx = torch.from_numpy(patch_gather_map().astype(np.int64)).reshape(-1)
    return x.reshape(b, -1)[:, idx].reshape(b, N_PATCH, PATCH_DIM)


# --------------------------------------------------------------------------------------------------
# MONAI blocks (restated; see oracle/monai_shim.py)
# --------------------------------------------------------------------------------------------------
def _lin(x, sd, prefix, bias=True):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"] if bias else None)


def _ln(x, sd, prefix):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], LN_EPS)


def patch_embedding(sd: Mapping[str, torch.Tensor], x: torch.Tensor, prefix="patch_embedding") -> torch.Tensor:
    """MONAI PatchEmbeddingBlock(perceptron) as constructed at vit.py:428-437 and called at vit.py:455/325:
    Linear(1024->768)(patchify(x)) + position_embeddings; dropout p=0."""
    p = patchify(x.float())
    return _lin(p, sd, prefix + ".patch_embeddings.1") + sd[prefix + ".position_embeddings"]


def self_attention(sd, x, prefix):
    """MONAI SABlock: qkv (no bias, vit.py passes qkv_bias=False), feature order (qkv, head, d),
    softmax(q k^T / sqrt(64)) v, heads re-concatenated, out_proj."""
    b, n, _ = x.shape
    qkv = F.linear(x, sd[prefix + ".qkv.weight"])                       # [B,N,2304]
    qkv = qkv.reshape(b, n, 3, HEADS, HEAD_DIM).permute(2, 0, 3, 1, 4)  # qkv b l(heads) h(tokens) d
    q, k, v = qkv[0], qkv[1], qkv[2]
    att = (torch.einsum("blxd,blyd->blxy", q, k) * (HEAD_DIM ** -0.5)).softmax(dim=-1)
    o = torch.einsum("bhxy,bhyd->bhxd", att, v)
    o = o.permute(0, 2, 1, 3).reshape(b, n, HIDDEN)
    return _lin(o, sd, prefix + ".out_proj")


def mlp(sd, x, prefix):
    """MONAI MLPBlock: linear1 -> exact-erf GELU -> linear2."""
    return _lin(F.gelu(_lin(x, sd, prefix + ".linear1")), sd, prefix + ".linear2")


def transformer_block(sd, x, prefix):
    """MONAI TransformerBlock (pre-LN): x += attn(norm1(x)); x += mlp(norm2(x))  (loop vit.py:463-466)."""
    x = x + self_attention(sd, _ln(x, sd, prefix + ".norm1"), prefix + ".attn")
    x = x + mlp(sd, _ln(x, sd, prefix + ".norm2"), prefix + ".mlp")
    return x


def _num_layers(sd) -> int:
    n = 0
    while f"blocks.{n}.norm1.weight" in sd:
        n += 1
    return n


# --------------------------------------------------------------------------------------------------
# Reference modules
# --------------------------------------------------------------------------------------------------
def vit_stage1(sd, x: torch.Tensor):
    """ViT_stage1.forward, vit.py:449-469: patch embed -> prepend cls (459-461) -> blocks -> LayerNorm (467).
    Returns (x [B,2049,768], hidden_states list)."""
    h = patch_embedding(sd, x)
    if "cls_token" in sd:
        h = torch.cat((sd["cls_token"].expand(h.shape[0], -1, -1), h), dim=1)
    hidden = []
    for i in range(_num_layers(sd)):
        h = transformer_block(sd, h, f"blocks.{i}")
        hidden.append(h)
    return _ln(h, sd, "norm"), hidden


def single_head_attention(q, k, v, drop_mask=None):
    """``attention`` helper, vit.py:25-33 == spatial_pooling_projector.py:8-16: one head, d_k = full
    embedding width (768), no mask.  ``drop_mask`` restates ``p_attn = dropout(p_attn)`` (vit.py:31-32) of train mode
    with an explicit keep mask (0 or 1/(1-p), shape of the probabilities) instead of torch's RNG; None = eval."""
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(q.size(-1))
    p = F.softmax(scores, dim=-1)
    if drop_mask is not None:
        p = p * drop_mask
    return torch.matmul(p, v), p


def regular_attention(sd, query, key, value, prefix="slice_guided_attention", drop=None):
    """regular_attention.forward, vit.py:50-64.  NB the residual is the *projected* query (:62).  ``drop`` = (mask on the
    probabilities [B,2048,32], mask on the output projection [B,2048,768]) restates the two nn.Dropout members of train
    mode (:60 ``dropout``, :62 ``dropout_2``) with explicit keep masks."""
    ql = _lin(query, sd, prefix + ".Wq")
    kl = _lin(key, sd, prefix + ".Wk")
    vl = _lin(value, sd, prefix + ".Wv")
    x, attn = single_head_attention(ql, kl, vl, None if drop is None else drop[0])
    x = _lin(x, sd, prefix + ".output_linear")
    if drop is not None:
        x = x * drop[1]
    x = _ln(ql + x, sd, prefix + ".norm")
    return x, attn


def patch_scores(sd, xp: torch.Tensor, image_2d: torch.Tensor, drop=None):
    """vit.py:332-339: slice features [B,32,768] guide a cross attention from the 2048 patch tokens;
    patch_score_proj (768->1) then Sigmoid -> scores [B,2048]."""
    b = xp.shape[0]
    sem = image_2d.float().reshape(b, N_SLICE, -1)
    ps, att = regular_attention(sd, xp, sem, sem, drop=drop)
    s = _lin(ps, sd, "patch_score_proj").reshape(b, xp.shape[1])
    return torch.sigmoid(s), att


def vit_stage2(sd, x: torch.Tensor, image_2d: torch.Tensor, drop=None):
    """ViT_stage2.forward, vit.py:315-357: patch embed -> scores (332-339) -> x*score (345) -> prepend cls
    (347-349) -> blocks (351-354) -> LayerNorm (355)."""
    xp = patch_embedding(sd, x)
    scores, _ = patch_scores(sd, xp, image_2d, drop=drop)
    h = xp * scores.unsqueeze(-1)
    if "cls_token" in sd:
        h = torch.cat((sd["cls_token"].expand(h.shape[0], -1, -1), h), dim=1)
    hidden = []
    for i in range(_num_layers(sd)):
        h = transformer_block(sd, h, f"blocks.{i}")
        hidden.append(h)
    return _ln(h, sd, "norm"), hidden


def _sub(sd, prefix):
    p = prefix + "."
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}


def dual_tower(sd, images, images_2d, select_feature="patch", remain="dual_vits"):
    """ViT3DTower_dual_encoders.forward, vit.py:926-948."""
    f1, _ = vit_stage1(_sub(sd, "vision_tower_stage1"), images)
    f2, _ = vit_stage2(_sub(sd, "vision_tower_stage2"), images, images_2d)
    if select_feature == "patch":
        f1, f2 = f1[:, 1:], f2[:, 1:]
    elif select_feature != "cls_patch":
        raise ValueError(f"Unexpected select feature: {select_feature}")
    if remain == "dual_vits":
        return f1, f2
    if remain == "3d_vit":
        return f1
    if remain == "2e3_vit":
        return f2
    return None


def packer_pool(feats: torch.Tensor) -> torch.Tensor:
    """spatial_pooling_projector.py:140-141: view(B,8,16,16,768); avg_pool3d kernel (1,4,4) -> [B,128,768]."""
    b = feats.shape[0]
    hr = feats.reshape(b, 8, 16, 16, HIDDEN)
    lr = F.avg_pool3d(hr.permute(0, 4, 1, 2, 3), kernel_size=(1, 4, 4)).permute(0, 2, 3, 4, 1)
    return lr.reshape(b, 128, HIDDEN)


def resolution_attention_v3(sd, lr: torch.Tensor, hr: torch.Tensor, prefix="resolution_attention", drop=None):
    """resolution_attention_v3.forward, spatial_pooling_projector.py:62-83 with kernel (1,4,4):
    each of the 128 pooled tokens attends (single head, d_k = 768) to its own 16 HR tokens;
    output_linear; LayerNorm(Wq(LR) + out)."""
    b = lr.shape[0]
    win = torch.from_numpy(packer_window_map().astype(np.int64))          # [128,16]
    hr_win = hr.reshape(b, N_PATCH, HIDDEN)[:, win]                        # [B,128,16,768]
    q = _lin(lr.reshape(b, 128, 1, HIDDEN), sd, prefix + ".Wq")
    k = _lin(hr_win, sd, prefix + ".Wk")
    v = _lin(hr_win, sd, prefix + ".Wv")
    # train mode (drop = (mask [B,128,16], mask [B,128,768])): dropout on the probabilities (:76) and dropout_2 (:78)
    x, _ = single_head_attention(q, k, v, None if drop is None else drop[0].reshape(b, 128, 1, 16))
    x = x.reshape(b, 128, HIDDEN)
    q = q.reshape(b, 128, HIDDEN)
    x = _lin(x, sd, prefix + ".output_linear")
    if drop is not None:
        x = x * drop[1]
    return _ln(q + x, sd, prefix + ".norm")


def visual_packer(sd, feats: torch.Tensor, drop=None) -> torch.Tensor:
    """VisualPacker_3d_phi_v3.forward, spatial_pooling_projector.py:138-146 -> [B,128,out_dim]."""
    feats = feats.float()
    lr = packer_pool(feats)
    a = resolution_attention_v3(sd, lr, feats, drop=drop)
    h = F.gelu(_lin(a, sd, "proj_mpls.0"))
    return _lin(h, sd, "proj_mpls.2")


def encode_images(tower_sd, proj_sd, proj2_sd, images, images_2d):
    """LamedMetaForCausalLM.encode_images, lamed_arch.py:122-141, dual-tower branch (125-132):
    packer(stage-1 feats) ++ packer2(stage-2 feats) along tokens -> [B,256,out_dim].
    ``proj2_sd=None`` reproduces the shared-projector fallback (128-131)."""
    f1, f2 = dual_tower(tower_sd, images, images_2d)
    o1 = visual_packer(proj_sd, f1)
    o2 = visual_packer(proj2_sd if proj2_sd is not None else proj_sd, f2)
    return torch.cat([o1, o2], dim=1)


def clip_encode_image(sd, vit_out: torch.Tensor) -> torch.Tensor:
    """M3DCLIP_stage1.encode_image tail, CLIP_stage1.py:100-101 + ``[:, 0]`` (:117):
    mm_vision_proj over all tokens, L2 normalise, keep the cls row -> [B,768]."""
    f = F.normalize(_lin(vit_out, sd, "mm_vision_proj"), dim=-1)
    return f[:, 0]


def contrastive_logits(all_image: torch.Tensor, all_text: torch.Tensor, logit_scale: torch.Tensor,
                       labels: torch.Tensor):
    """image_text_contrastive_learning, CLIP_stage1.py:141-155 (gather_loss=True, local_loss=False)."""
    lpi = logit_scale * all_image @ all_text.T
    lpt = lpi.T
    loss = (F.cross_entropy(lpi, labels) + F.cross_entropy(lpt, labels)) / 2
    return loss, lpi, lpt


def gather_features_single(image_features, text_features, world: Sequence = None):
    """gather_features, utils/dist_utils.py:280-306: concatenation of every rank's [B_loc,768] blocks in
    rank order.  ``world`` is the list of (image, text) pairs of all ranks (single-process restatement)."""
    if world is None:
        return image_features, text_features
    return torch.cat([w[0] for w in world], 0), torch.cat([w[1] for w in world], 0)


def slice_extract(images: torch.Tensor, out_hw=(224, 224)) -> torch.Tensor:
    """K12 online variant, vit.py:529-531 / 805-807: trilinear (32,256,256)->(32,224,224)
    (align_corners=False), expand to 3 channels, -> [B*32,3,224,224]."""
    b = images.shape[0]
    r = F.interpolate(images.float(), size=(IMG[0],) + tuple(out_hw), mode="trilinear", align_corners=False)
    r = r.expand(-1, 3, -1, -1, -1).permute(0, 2, 1, 3, 4)
    return r.reshape(b * IMG[0], 3, out_hw[0], out_hw[1]).contiguous()


# --------------------------------------------------------------------------------------------------
# Volume ingest: Data/data_processing/CT-RATE/CT-RATE_nii_to_3D_volume_npy_file.py
# PARITY UNPINNED against the script itself: MONAI's CropForeground / Resize sources are not in the reference tree (restated
# here from their documented behaviour: select_fn x > 0, margin 0, Resize(mode='bilinear') == trilinear with
# align_corners=False on a 3-D volume), the script resamples in float64, and it needs nibabel + the metadata CSV.
# --------------------------------------------------------------------------------------------------
def preprocess_resampled_shape(raw_shape, xy_spacing, z_spacing, target=(1.5, 0.75, 0.75)):
    """new_shape of resize_array (script lines 30-35) for the (2,0,1)-transposed volume."""
    n0, n1, n2 = raw_shape
    cur = (z_spacing, xy_spacing, xy_spacing)
    orig = (n2, n0, n1)
    return tuple(int(orig[i] * (cur[i] / target[i])) for i in range(3))


def foreground_bbox(x: torch.Tensor):
    """MONAI generate_spatial_bounding_box(select_fn=is_positive, margin=0) on a [z,a,b] volume: (lo, hi) with hi exclusive;
    the full extent when nothing is positive."""
    fg = x > 0
    if not bool(fg.any()):
        return (0, 0, 0), tuple(x.shape)
    lo, hi = [], []
    for d in range(3):
        other = tuple(i for i in range(3) if i != d)
        idx = torch.nonzero(fg.any(dim=other)).reshape(-1)
        lo.append(int(idx[0]))
        hi.append(int(idx[-1]) + 1)
    return tuple(lo), tuple(hi)


def preprocess_volume(raw: torch.Tensor, slope, intercept, xy_spacing, z_spacing, out_size=(32, 256, 256),
                      dtype=torch.float32, return_intermediates=False):
    """nii_img_to_tensor + transform (script lines 41-117) on an in-memory voxel array raw [n0,n1,n2]."""
    x = (slope * raw.to(dtype) + intercept).clamp(-1000, 200)                       # lines 80-86
    x = x.permute(2, 0, 1)[None, None]                                              # lines 90, 93-94
    new_shape = preprocess_resampled_shape(raw.shape, xy_spacing, z_spacing)
    x = F.interpolate(x, size=new_shape, mode="trilinear", align_corners=False)[0, 0]   # lines 37, 97
    x = x.float()                                                                   # line 102
    mn, mx = x.min(), x.max()                                                       # lines 111-112
    xn = (x - mn) / torch.clamp(mx - mn, min=1e-8)                                  # lines 113-114
    lo, hi = foreground_bbox(xn)                                                    # CropForeground, line 122
    crop = xn[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
    out = F.interpolate(crop[None, None], size=tuple(out_size), mode="trilinear", align_corners=False)[0]   # Resize
    if return_intermediates:
        return out, x, (float(mn), float(mx)), lo + hi
    return out


# --------------------------------------------------------------------------------------------------
# Parity metrics used by every test (north_star tolerances)
# --------------------------------------------------------------------------------------------------
def parity_metrics(got: torch.Tensor, ref: torch.Tensor) -> dict:
    g = got.detach().double().cpu().reshape(-1)
    r = ref.detach().double().cpu().reshape(-1)
    cos = float(torch.dot(g, r) / (g.norm() * r.norm() + 1e-300))
    max_rel = float((g - r).abs().max() / (r.abs().max() + 1e-300))
    rms_rel = float((g - r).norm() / (r.norm() + 1e-300))
    return {"cos": cos, "max_rel": max_rel, "rms_rel": rms_rel}


# --------------------------------------------------------------------------------------------------
# Online 2D-slice branch: ViT4LLM_v3_med2e3.forward, vit.py:805-808, and the offline extractor
# Data/data_processing/CT-RATE/CT-RATE_2D_to_npy_file.py:75-98 (`model.visual.trunk` of BiomedCLIP).
# PARITY UNPINNED: the trunk is timm 1.0.11's VisionTransformer (vit_base_patch16_224, num_classes=0) created by
# open_clip 2.24.0 (requirements.txt:73,142); neither package nor the BiomedCLIP weights are available offline.  Restated
# here from timm's published forward: conv stem (16x16/16) -> cls token prepended -> + pos_embed -> 12 pre-LN blocks
# (LayerNorm eps 1e-6, qkv WITH bias in (qkv, head, d) order, scale 64^-0.5, exact GELU) -> LayerNorm -> token pooling.
# --------------------------------------------------------------------------------------------------
TRUNK_LN_EPS = 1e-6


def _ln6(x, sd, prefix):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], TRUNK_LN_EPS)


def timm_vit_trunk(sd, x2d: torch.Tensor) -> torch.Tensor:
    """timm VisionTransformer.forward (forward_features + token pooling, head = Identity): [N,3,224,224] -> [N,768]."""
    n = x2d.shape[0]
    x = F.conv2d(x2d, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=16)    # [N,768,14,14]
    x = x.flatten(2).transpose(1, 2)                                                            # [N,196,768]
    x = torch.cat((sd["cls_token"].expand(n, -1, -1), x), dim=1) + sd["pos_embed"]
    i = 0
    while f"blocks.{i}.norm1.weight" in sd:
        p = f"blocks.{i}"
        h = _ln6(x, sd, p + ".norm1")
        qkv = F.linear(h, sd[p + ".attn.qkv.weight"], sd[p + ".attn.qkv.bias"])
        qkv = qkv.reshape(n, -1, 3, HEADS, HEAD_DIM).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        att = ((q * HEAD_DIM ** -0.5) @ k.transpose(-2, -1)).softmax(dim=-1)
        o = (att @ v).transpose(1, 2).reshape(n, -1, HIDDEN)
        x = x + F.linear(o, sd[p + ".attn.proj.weight"], sd[p + ".attn.proj.bias"])
        h = _ln6(x, sd, p + ".norm2")
        h = F.gelu(F.linear(h, sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"]))
        x = x + F.linear(h, sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"])
        i += 1
    return _ln6(x, sd, "norm")[:, 0]


def slice_branch(sd, images: torch.Tensor) -> torch.Tensor:
    """vit.py:805-808 + :816: resize to (32,224,224), expand to 3 channels, trunk, view(batch, slice_num, -1)."""
    b = images.shape[0]
    return timm_vit_trunk(sd, slice_extract(images)).reshape(b, IMG[0], -1)
