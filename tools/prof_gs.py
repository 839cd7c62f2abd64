import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from eval_driving_safety_b200 import dsgn, ops, synthetic
dev = torch.device("cuda", 0); g = torch.Generator().manual_seed(0)
cfg = dsgn.default_cfg(); fu, b, P, PR = synthetic.make_calib(1)
psv = torch.randn(1, 48, 96, 312, 64, generator=g).to(dev).permute(0, 4, 1, 2, 3).requires_grad_(True)
grid3 = dsgn.lifting_grid(cfg, P, (96, 312)).to(dev).contiguous()
plan3 = ops.GridPlan(grid3, (48, 96, 312), True)
for _ in range(2):
    out = ops.grid_sample(psv, grid3, True, plan3)
    (gi,) = torch.autograd.grad(out, psv, torch.ones_like(out))
torch.cuda.synchronize()
