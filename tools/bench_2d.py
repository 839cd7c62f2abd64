"""2-D extractor: (a) per-layer conv timings, own tcgen05 kernels (3xTF32 / plain TF32) vs cuDNN TF32,
(b) whole extractor fwd+bwd per pair under graph replay for each implementation."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from eval_driving_safety_b200 import dsgn, ops
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)


def timeit(fn, iters=20):
    """ms per call with the launches replayed from a CUDA graph (the host's ~30 us per call would otherwise hide
    every kernel shorter than that)."""
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(iters):
                fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * iters)


# (n, cin, cout, h, w, k, stride, dil)
LAYERS = [(2, 32, 32, 192, 624, 3, 1, 1), (2, 32, 64, 192, 624, 3, 2, 1), (2, 64, 64, 96, 312, 3, 1, 1),
          (2, 128, 128, 96, 312, 3, 1, 1), (2, 128, 128, 96, 312, 3, 1, 2), (2, 320, 128, 96, 312, 3, 1, 1),
          (1, 320, 128, 192, 304, 3, 1, 1), (1, 128, 64, 192, 304, 3, 1, 1), (2, 128, 32, 96, 312, 1, 1, 1)]
print("halo kernel:", os.environ.get("B2_CONV2D_HALO", "1"))
print("layer: ms fwd (TFLOP/s-equivalent, graph replay)  own 3xTF32 | own TF32 | cuDNN TF32")
for (n, ci, co, h, w, k, s, d) in LAYERS:
    x = torch.randn(n, ci, h, w, device=dev).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(co, ci, k, k, device=dev) / (ci * k * k) ** 0.5)
    wcl = wt.contiguous(memory_format=torch.channels_last)
    fl = 2 * ci * co * k * k * n * ((h - 1) // s + 1) * ((w - 1) // s + 1)
    res = []
    for split in (True, False):
        ops.set_conv2d_split(split)
        res.append(timeit(lambda: ops.conv2d(x, wt, None, s, d)))
    ops.set_conv2d_split(True)
    res.append(timeit(lambda: F.conv2d(x, wcl, None, s, d * (k // 2), d)))
    print("%s: " % ((n, ci, co, h, w, k, s, d),) + " | ".join("%.4f ms (%.0f TF/s)" % (t, fl / t / 1e9) for t in res), flush=True)

cfg = dsgn.default_cfg()
model = dsgn.build_model(cfg, seed=1, device=dev)
fe = model.feature_extraction
for impl, split in (("b2", True), ("b2", False), ("cudnn", True)):
    dsgn.set_backbone_impl(impl); ops.set_conv2d_split(split)
    for npairs in (1, 2):
        x = torch.randn(2 * npairs, 3, 384, 1248, device=dev)
        gf = gr = None
        def work():
            global gf, gr
            a = x.detach().requires_grad_(True)
            f, r = fe(a, rpn_samples=npairs)
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
        print("extractor %s split=%s pairs %d: %.3f ms per replay = %.3f ms per pair" % (impl, split, npairs, t, t / npairs), flush=True)
dsgn.set_backbone_impl("b2"); ops.set_conv2d_split(True)
