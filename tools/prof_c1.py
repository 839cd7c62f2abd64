"""One forward + backward of the Cout = 1 head at KITTI size, for `ncu --set full -k regex:c1`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eval_driving_safety_b200 import ops
g = torch.Generator().manual_seed(0)
x = torch.randn(1, 48, 96, 312, 64, generator=g).cuda().permute(0, 4, 1, 2, 3).requires_grad_(True)
w = (torch.randn(1, 64, 3, 3, 3, generator=g) * 0.05).cuda()
for _ in range(2):
    y = ops.conv3d_c1(x, w)
    torch.autograd.grad(y, x, torch.ones_like(y))
torch.cuda.synchronize()
