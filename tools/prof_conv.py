"""One launch each of the stride-1 (N-stacked) and the transposed (class-stacked) tcgen05 kernels at KITTI sizes, for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eval_driving_safety_b200 import ops
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
def cl(shape):
    n, c, d, h, w = shape
    return torch.randn(n, d, h, w, c, generator=g).to(dev).permute(0, 4, 1, 2, 3)
x1 = cl((1, 64, 48, 96, 312)); w1 = (torch.randn(27, 64, 64, generator=g) * 0.02).to(dev)
x2 = cl((1, 128, 24, 48, 156)); w2 = (torch.randn(27, 64, 128, generator=g) * 0.02).to(dev)
for _ in range(2):
    ops._conv_call(x1, w1, 1, 0, 0)
    ops._conv_call(x2, w2, 2, 1, 0)
torch.cuda.synchronize()
