import cProfile, pstats, io, os, sys, time, torch
sys.path.insert(0, os.getcwd())
import hsenet_b200 as H
torch.manual_seed(0)
dev = torch.device("cuda:0")
enc = H.HSENetVisualEncoder(H.VisionConfig()).eval().requires_grad_(False).to(dev)
x = torch.rand(8, 1, 32, 256, 256, device=dev); s = torch.randn(8, 32, 768, device=dev)
def step():
    with torch.no_grad():
        return enc.vision_tower(x, s)
for conc in (True, False):
    enc.vision_tower.concurrent_towers = conc
    for _ in range(5): step()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(20): step()
    host = (time.perf_counter() - t) / 20 * 1e3
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t) / 20 * 1e3
    print(f"concurrent={conc}: host enqueue {host:.2f} ms/step, wall {tot:.2f} ms/step")
enc.vision_tower.concurrent_towers = True
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
pr.disable(); torch.cuda.synchronize()
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(22); print(st.getvalue()[:4500])
