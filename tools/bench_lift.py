"""Lifting kernels at KITTI size: fused 4-lane forward vs the two generic kernels; wide-lane gather backward (B2_GS_BWD_WIDE=0
in the environment selects the one-float4-per-lane variant).  CUDA events, L2 flushed between launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eval_driving_safety_b200 import dsgn, ops, synthetic
dev = torch.device("cuda", 0)
flush = torch.empty(64 * 1024 * 1024, device=dev)
cfg = dsgn.default_cfg()
fu, b, P, PR = synthetic.make_calib(1)
g = torch.Generator().manual_seed(0)
psv = torch.randn(1, 48, 96, 312, 64, generator=g).to(dev).permute(0, 4, 1, 2, 3)
img = torch.randn(1, 96, 312, 32, generator=g).to(dev).permute(0, 3, 1, 2)
grid3 = dsgn.lifting_grid(cfg, P, (96, 312)).to(dev).contiguous()
grid2 = grid3[..., :2].contiguous().view(1, 192 * 20, 304, 2)
plan3, plan2 = ops.GridPlan(grid3, (48, 96, 312), True), ops.GridPlan(grid2, (96, 312), True)
gout = torch.randn(1, 192, 20, 304, 96, generator=g).to(dev).permute(0, 4, 1, 2, 3)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


ref = None
for fused in (False, True):
    ops.LIFT_FUSED = fused
    out = ops.lift(psv, img, grid3, plan3, plan2, True)
    if ref is None:
        ref = out.clone()
    t = timeit(lambda: ops.lift(psv, img, grid3, plan3, plan2, True))
    print("lift fwd fused=%s: %.4f ms  (bit-identical to unfused: %s)" % (fused, t, torch.equal(out, ref)), flush=True)
ops.LIFT_FUSED = True
a, c = psv.detach().requires_grad_(True), img.detach().requires_grad_(True)
o = ops.lift(a, c, grid3, plan3, plan2, True)
torch.autograd.grad(o, [a, c], gout, retain_graph=True)
with ops.profile() as prof:
    for _ in range(5):
        flush.fill_(1.0)
        torch.autograd.grad(o, [a, c], gout, retain_graph=True)
s = prof.summary()
print("lift bwd (B2_GS_BWD_WIDE=%s): 3-D %.4f ms, 2-D %.4f ms" % (os.environ.get("B2_GS_BWD_WIDE", "1"),
      s["grid_sample3d_bwd"]["ms"] / 5, s["grid_sample2d_bwd"]["ms"] / 5), flush=True)
