"""Cost-volume forward / backward at KITTI size (CUDA events, L2 flushed between launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eval_driving_safety_b200 import dsgn, ops, synthetic
dev = torch.device("cuda", 0)
flush = torch.empty(64 * 1024 * 1024, device=dev)
g = torch.Generator().manual_seed(0)
cfg = dsgn.default_cfg()
fu, b, P, PR = synthetic.make_calib(1)
shifts = dsgn.plane_shifts(cfg, fu, b).to(dev)
L = torch.randn(1, 96, 312, 32, generator=g).to(dev).permute(0, 3, 1, 2).requires_grad_(True)
R = torch.randn(1, 96, 312, 32, generator=g).to(dev).permute(0, 3, 1, 2).requires_grad_(True)
gc = torch.randn(1, 48, 96, 312, 64, generator=g).to(dev).permute(0, 4, 1, 2, 3)
cv = ops.build_cost_volume(L, R, shifts)
for _ in range(3):
    torch.autograd.grad(cv, [L, R], gc, retain_graph=True)
with ops.profile() as prof:
    for _ in range(10):
        flush.fill_(1.0)
        ops.build_cost_volume(L, R, shifts)
        flush.fill_(1.0)
        torch.autograd.grad(cv, [L, R], gc, retain_graph=True)
s = prof.summary()
for k in ("cost_volume_fwd", "cost_volume_bwd"):
    print("%s: %.4f ms  %.0f GB/s (%.2f of 6543)" % (k, s[k]["ms"] / 10, s[k]["per_s"] / 1e9, s[k]["per_s"] / 1e9 / 6543), flush=True)
