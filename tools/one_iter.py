"""One full-size pair-iteration (after one warm-up) for ncu launch lists / profiles."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eval_driving_safety_b200 import attack, dsgn, ops, synthetic
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
cfg = dsgn.default_cfg()
model = dsgn.build_model(cfg, seed=1, device=dev)
calib = synthetic.make_calib(1)
p = synthetic.make_pair(0)
xL, xR, disp = p["imgL"].to(dev), p["imgR"].to(dev), p["disp_L"].to(dev)
labels = {k: v.to(dev) for k, v in synthetic.make_labels(cfg, 1, 7).items()}
mean = torch.tensor(attack.IMAGENET_MEAN, device=dev).view(1, 3, 1, 1)
std = torch.tensor(attack.IMAGENET_STD, device=dev).view(1, 3, 1, 1)
cL, cR = xL * std + mean, xR * std + mean
for it in range(iters):
    if it == iters - 1:
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("timed_iteration")
        torch.cuda.cudart().cudaProfilerStart()
    a, b = xL.detach().requires_grad_(True), xR.detach().requires_grad_(True)
    out = model(a, b, calib[0], calib[1], calib[2], calibs_Proj_R=calib[3])
    loss = dsgn.attack_loss(cfg, out, disp, labels)
    gL, gR = torch.autograd.grad(loss, [a, b])
    attack.pgd_step_pair(xL, gL.contiguous(), cL, xR, gR.contiguous(), cR, 0.0075, 0.03, inplace=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", loss.item(), "launches", ops.LAUNCH_COUNT)
