"""Small end-to-end invocation for compute-sanitizer: 2-layer towers (so that the folded LayerNorm epilogues run) + packer at
B = 1 in both precisions, the dual tower (shared patch embedding), and one training step of ViT_stage2 / the packer in
train() mode (dropout kernels, attention backward, split-K weight gradients)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hsenet_b200 as H
geom = dict(in_channels=1, img_size=(32, 256, 256), patch_size=(4, 16, 16), pos_embed="perceptron", spatial_dims=3,
            classification=True, num_layers=2)
torch.manual_seed(0)
dev = torch.device("cuda:0")
v1 = H.ViT_stage1(**geom).eval().requires_grad_(False).to(dev)
v2 = H.ViT_stage2(**geom).eval().requires_grad_(False).to(dev)
pk = H.VisualPacker_3d_phi_v3((32, 256, 256), (4, 16, 16), 768, 3072, "mlp", 2).eval().requires_grad_(False).to(dev)
for m in (v1, v2):
    m.use_cuda_graph = False
x = torch.rand(1, 1, 32, 256, 256, device=dev); s = torch.randn(1, 32, 768, device=dev)
with torch.no_grad():
    for prec in ("bf16", "fp32_verify"):
        with H.precision(prec):
            a, _ = v1(x); b, _ = v2(x, s); c = pk(v1.last_patch_tokens)
            torch.cuda.synchronize()
            print(prec, float(a.float().abs().mean()), float(b.float().abs().mean()), float(c.float().abs().mean()))
    os.environ["HSENET_ATT_MAXFREE"] = "1"
    with H.precision("bf16"):
        a, _ = v1(x)
    os.environ["HSENET_ATT_MAXFREE"] = "0"
    print("maxfree", float(a.float().abs().mean()))
# training step in train() mode: dropout masks, attention backward, weight gradients
for prec in ("bf16", "fp32_verify"):
    t2 = H.ViT_stage2(**dict(geom, num_layers=1)).to(dev).train()
    p2 = H.VisualPacker_3d_phi_v3((32, 256, 256), (4, 16, 16), 768, 3072, "mlp", 2).to(dev).train()
    with H.precision(prec):
        y, _ = t2(x, s)
        z = p2(t2.last_patch_tokens)
        (y.float().mean() + z.float().mean()).backward()
    torch.cuda.synchronize()
    g = [p.grad for p in list(t2.parameters()) + list(p2.parameters())]
    print(prec, "train", all(t is not None and bool(torch.isfinite(t).all()) for t in g))
print("done")
