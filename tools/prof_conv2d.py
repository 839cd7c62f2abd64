"""One launch per (shape, operand split) of the stride-1 3x3 conv2d kernel at extractor sizes, for `ncu --set full`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eval_driving_safety_b200 import ops
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)


def cl2(n, c, h, w):
    return torch.randn(n, h, w, c, generator=g).to(dev).permute(0, 3, 1, 2)


cases = [(cl2(2, c, h, w), (torch.randn(c, c, 3, 3, generator=g) / (3 * c ** 0.5)).to(dev))
         for (c, h, w) in ((32, 192, 624), (64, 96, 312), (128, 96, 312))]
for rep in range(2):
    for split in (1, 0):
        ops.set_conv2d_split(split)
        for x, w in cases:
            ops.conv2d(x, w)
torch.cuda.synchronize()
