import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eval_driving_safety_b200 import ops
g = torch.Generator().manual_seed(0)
x = torch.randn(1, 48, 96, 312, 64, generator=g).cuda().permute(0, 4, 1, 2, 3).requires_grad_(True)
gamma, beta = torch.ones(64).cuda(), torch.zeros(64).cuda()
gy = torch.randn(1, 48, 96, 312, 64, generator=g).cuda().permute(0, 4, 1, 2, 3)
for relu, res in ((True, None), (False, x.detach() * 0.5)):
    for it in range(3):
        y = ops.groupnorm_act(x, gamma, beta, 32, 1e-5, relu=relu, res=res)
        (gx,) = torch.autograd.grad(y, x, gy)
torch.cuda.synchronize()
print("ok")
