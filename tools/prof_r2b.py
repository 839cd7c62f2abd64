"""One launch each of the kernels changed at the end of round 2, at their full sizes, for `ncu --set full`:
fused lifting forward (6 resident blocks per SM), x4 depth head fwd / bwd, warp-per-pixel RoIAlign backward (pyramid,
R = 256, C = 256, P = 7), conv2d with the GroupNorm-sum epilogues (forward statistics, backward sums)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eval_driving_safety_b200 import dsgn, ops, synthetic, stereo_rcnn as S
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
cost1 = torch.randn(1, 1, 48, 96, 312, generator=g).to(dev).requires_grad_(True)
feats = [torch.randn(1, 256, h, w, generator=g).to(dev).requires_grad_(True) for (h, w) in ((150, 497), (75, 249), (38, 125), (19, 63))]
rois, _ = S.synthetic_rois(256, 600, 1987, seed=0)
rois = rois.to(dev)
x64 = cl2(2, 64, 96, 312).requires_grad_(True)
wa, wb = (torch.randn(64, 64, 3, 3, generator=g) / 24).to(dev), (torch.randn(64, 64, 3, 3, generator=g) / 24).to(dev)
gam, bet = torch.ones(64, device=dev), torch.zeros(64, device=dev)
for _ in range(2):
    ops.lift(psv, img, grid3, plan3, plan2, True)
    d = ops.depth_head(cost1, (192, 384, 1248), cfg.min_depth, cfg.depth_interval)
    torch.autograd.grad(d, cost1, torch.ones_like(d))
    pooled = S.pyramid_roi_feat(feats, rois, 600.0, 7)
    torch.autograd.grad(pooled, feats, torch.ones_like(pooled))
    ya, part = ops.conv2d_with_stats(x64, wa)
    h = ops.groupnorm_act(ya, gam, bet, 32, 1e-5, relu=True, partial=part)
    yb = ops.conv2d(h, wb)
    torch.autograd.grad(yb, x64, torch.ones_like(yb))
torch.cuda.synchronize()
