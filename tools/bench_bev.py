"""bev_pool forward / backward at the KITTI voxel-grid size (CUDA events, L2 flushed between launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eval_driving_safety_b200 import ops
g = torch.Generator().manual_seed(0)
v = torch.randn(1, 192, 20, 304, 64, generator=g).cuda().permute(0, 4, 1, 2, 3).requires_grad_(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, n=6):
    r = []
    for _ in range(n + 2):
        flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); r.append(e0.elapsed_time(e1))
    return sorted(r[2:])[n // 2]
bev = ops.bev_pool(v, 4)
gb = torch.randn_like(bev)
print("bev fwd %.3f ms" % t(lambda: ops.bev_pool(v.detach(), 4)))
print("bev fwd+bwd %.3f ms" % t(lambda: torch.autograd.grad(ops.bev_pool(v, 4), v, gb)))
