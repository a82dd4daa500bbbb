"""Generates tests/golden/*.npz by running the reference's OWN files (vit.py, spatial_pooling_projector.py, loaded
unmodified from /root/reference through oracle/monai_shim.py) on the deterministic recipe of recipe.py.

    python tests/golden/make_golden.py          (needs /root/reference; run in the build container only)

The fixtures are small (sub-sampled rows + per-row norms) and travel to the GPU box, where /root/reference is absent.
They pin (a) the oracle restatement (tests/test_oracle.py) and (b) the CUDA path (tests/test_gpu_golden.py) to the
reference's outputs."""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import reference_loader as RL  # noqa: E402
from recipe import PACKER_ROWS, SAMPLE_ROWS, recipe_inputs, recipe_state_dict, recipe_tokens  # noqa: E402

GEOM = dict(in_channels=1, img_size=(32, 256, 256), patch_size=(4, 16, 16), pos_embed="perceptron",
            spatial_dims=3, classification=True)


def load_recipe(module, seed):
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    module.load_state_dict(recipe_state_dict(shapes, seed), strict=True)
    return module.eval()


def summarise(x, rows):
    x = x.detach().float()
    return {"rows": x[:, rows].numpy(), "row_norms": x.norm(dim=-1).numpy(),
            "mean": np.float64(x.double().mean().item()), "abs_max": np.float32(x.abs().max().item())}


def main():
    assert RL.available(), "reference not found"
    torch.set_num_threads(os.cpu_count() or 8)
    vit, pk = RL.vit(), RL.packer()
    out = {}
    with torch.no_grad(), RL.quiet():
        for layers in (2, 12):
            x, s = recipe_inputs(1)
            m1 = load_recipe(vit.ViT_stage1(num_layers=layers, **GEOM), seed=layers)
            y1, h1 = m1(x)
            d = summarise(y1, SAMPLE_ROWS)
            d["hidden0_rows"] = h1[0][:, SAMPLE_ROWS].numpy()
            np.savez_compressed(os.path.join(HERE, f"vit_stage1_L{layers}.npz"), **d)
            m2 = load_recipe(vit.ViT_stage2(num_layers=layers, **GEOM), seed=100 + layers)
            y2, _ = m2(x, s)
            d = summarise(y2, SAMPLE_ROWS)
            # scores as the reference computes them (vit.py:332-339)
            xp = m2.patch_embedding(x.clone())
            ps, att = m2.slice_guided_attention(xp, s.view(1, 32, -1), s.view(1, 32, -1))
            d["scores"] = m2.patch_score_norm(m2.patch_score_proj(ps).view(1, 2048)).numpy()
            d["attn_rows"] = att[:, SAMPLE_ROWS[:4]].numpy()
            np.savez_compressed(os.path.join(HERE, f"vit_stage2_L{layers}.npz"), **d)
            out[layers] = (y1, y2)
        # packer on recipe tokens and on the stage-1 features (non-contiguous [:,1:] view, as the tower hands over)
        p = load_recipe(pk.VisualPacker_3d_phi_v3((32, 256, 256), (4, 16, 16), 768, 3072, "mlp", 2), seed=7)
        t = recipe_tokens(2)
        d = summarise(p(t), PACKER_ROWS)
        d2 = summarise(p(out[2][0][:, 1:]), PACKER_ROWS)
        d.update({"feat_" + k: v for k, v in d2.items()})
        np.savez_compressed(os.path.join(HERE, "packer.npz"), **d)
        # dual tower wrapper: shapes / dispatch only (weights random): key counts are the reference's own KAT
        cfg = RL.TowerConfig()
        tower = vit.ViT3DTower_dual_encoders(cfg)
        keys = list(tower.state_dict().keys())
        # integer maps derived by executing the reference's einops / reshape-permute code on index tensors
        idx = torch.arange(32 * 256 * 256, dtype=torch.float32).reshape(1, 1, 32, 256, 256)
        pmap = m1.patch_embedding.patch_embeddings[0](idx)[0].to(torch.int32).numpy()           # [2048,1024]
        hr_idx = torch.arange(2048, dtype=torch.float32).reshape(1, 8, 16, 16, 1)
        S_d, S_w, S_h = 1, 4, 4
        hr = hr_idx.reshape(1, 8, 1, 16, 1, 16, 1, 1).view(1, 8 // S_d, S_d, 16 // S_w, S_w, 16 // S_h, S_h, 1)
        hr = hr.permute(0, 2, 4, 6, 1, 3, 5, 7).contiguous().view(1, S_d * S_w * S_h, 128, 1).permute(0, 2, 1, 3)
        wmap = hr[0, :, :, 0].to(torch.int32).numpy()                                              # [128,16]
        np.savez_compressed(
            os.path.join(HERE, "maps.npz"),
            patch_map_sha256=np.frombuffer(hashlib.sha256(pmap.tobytes()).digest(), dtype=np.uint8),
            patch_map_rows=pmap[[0, 1, 17, 255, 256, 1000, 2047]],
            patch_map_row_ids=np.array([0, 1, 17, 255, 256, 1000, 2047]),
            window_map=wmap,
            tower_keys=np.array(keys),
            n_stage1=np.int64(sum("vision_tower_stage1" in k for k in keys)),
            n_stage2=np.int64(sum("vision_tower_stage2" in k for k in keys)),
            packer_keys=np.array(list(p.state_dict().keys())))
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
