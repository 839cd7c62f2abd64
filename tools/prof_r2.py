"""One launch each of the round-2 kernels at KITTI sizes, for `ncu --set full` (tools/run_gpu_batch.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eval_driving_safety_b200 import attack, dsgn, ops, synthetic
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
cfg = dsgn.default_cfg()
fu, b, P, PR = synthetic.make_calib(1)


def cl3(n, c, d, h, w):
    return torch.randn(n, d, h, w, c, generator=g).to(dev).permute(0, 4, 1, 2, 3)


def cl2(n, c, h, w):
    return torch.randn(n, h, w, c, generator=g).to(dev).permute(0, 3, 1, 2)


psv, img = cl3(1, 64, 48, 96, 312), cl2(1, 32, 96, 312)
grid3 = dsgn.lifting_grid(cfg, P, (96, 312)).to(dev).contiguous()
grid2 = grid3[..., :2].contiguous().view(1, 192 * 20, 304, 2)
plan3, plan2 = ops.GridPlan(grid3, (48, 96, 312), True), ops.GridPlan(grid2, (96, 312), True)
a, c = psv.detach().requires_grad_(True), img.detach().requires_grad_(True)
gout = cl3(1, 96, 192, 20, 304)
x64, w64 = cl2(2, 64, 96, 312), (torch.randn(64, 64, 3, 3, generator=g) / 24).to(dev)
x128, w128 = cl2(2, 128, 96, 312), (torch.randn(128, 128, 3, 3, generator=g) / 34).to(dev)
xd, wd = cl3(1, 128, 24, 48, 156), (torch.randn(27, 64, 128, generator=g) * 0.02).to(dev)
xs, ws = cl3(1, 64, 48, 96, 312), (torch.randn(27, 128, 64, generator=g) * 0.02).to(dev)
shifts = dsgn.plane_shifts(cfg, fu, b).to(dev)
L, Rr = cl2(1, 32, 96, 312).requires_grad_(True), cl2(1, 32, 96, 312).requires_grad_(True)
gc = cl3(1, 64, 48, 96, 312)
xi = [torch.randn(1, 3, 384, 1248, generator=g).to(dev) for _ in range(12)]
for _ in range(2):
    o = ops.lift(a, c, grid3, plan3, plan2, True)
    torch.autograd.grad(o, [a, c], gout)
    ops.conv2d(x64, w64)
    ops.conv2d(x128, w128)
    ops._conv_call(xd, wd, 2, 1, 0)          # transposed 128 -> 64 (class-stacked kernel)
    ops._conv_call(xs, ws, 2, 0, 0)          # stride-2 64 -> 128 (generic kernel)
    cv = ops.build_cost_volume(L, Rr, shifts)
    torch.autograd.grad(cv, [L, Rr], gc)
    attack._pgd_update_sets(xi[0:4], xi[4:8], xi[8:12], xi[0:4], 0.0075, 0.03, attack.IMAGENET_MEAN, attack.IMAGENET_STD, 0.0, 1.0)
torch.cuda.synchronize()
