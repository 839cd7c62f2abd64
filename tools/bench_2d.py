"""2-D extractor fwd+bwd (graph replay) for 1 pair (batch 2 = L,R) vs 2 pairs (batch 4): is it latency-bound?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eval_driving_safety_b200 import dsgn
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
cfg = dsgn.default_cfg()
model = dsgn.build_model(cfg, seed=1, device=dev)
fe = model.feature_extraction
for npairs in (1, 2, 4):
    x = torch.randn(2 * npairs, 3, 384, 1248, device=dev).contiguous(memory_format=torch.channels_last)
    gf = gr = None
    def work():
        a = x.detach().requires_grad_(True)
        f, r = fe(a, rpn_samples=npairs)
        global gf, gr
        if gf is None:
            gf, gr = torch.randn_like(f), torch.randn_like(r)
        torch.autograd.grad([f, r], a, [gf, gr])
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): work()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            work()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): g.replay()
        e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10
    print("pairs %d: %.3f ms per replay = %.3f ms per pair" % (npairs, t, t / npairs))
