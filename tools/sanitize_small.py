"""Small end-to-end invocation for compute-sanitizer (memcheck): 1-layer towers + packer, B = 1, both precisions."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hsenet_b200 as H
geom = dict(in_channels=1, img_size=(32, 256, 256), patch_size=(4, 16, 16), pos_embed="perceptron", spatial_dims=3,
            classification=True, num_layers=1)
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
print("done")
