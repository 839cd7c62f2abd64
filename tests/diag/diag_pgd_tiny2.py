"""Free-running tiny PGD in b2 mode, instrumented: is the gradient inside the loop the one a fresh call gives,
and is the update what the oracle's rule makes of that gradient?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import rel_err
from oracle import attack_ref as A, dsgn_ref as R
from eval_driving_safety_b200 import attack, dsgn, ops, synthetic
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
H, W = 32, 64
cfg_r, cfg_p = R.tiny_cfg(), dsgn.tiny_cfg()
ref = R.build_model(cfg_r, seed=1)
model = dsgn.StereoNet(cfg_p); model.load_state_dict(ref.state_dict()); model = model.freeze().cuda()
pair = synthetic.make_pair(0, H, W, max_depth=8.4)
calib = synthetic.make_calib(1, scale=H / 384, cu=W / 2, cv=H / 2)
labels = R.make_labels(cfg_r, 1, 7); lab = {k: v.cuda() for k, v in labels.items()}
eps, alpha, K = 0.03, 0.03 / 4, 3
ops.set_conv_impl(1)
disp = pair["disp_L"].cuda()


def grads(xL, xR):
    a, b = xL.detach().clone().requires_grad_(True), xR.detach().clone().requires_grad_(True)
    o = model(a, b, *calib[:3], calibs_Proj_R=calib[3])
    l = dsgn.attack_loss(cfg_p, o, disp, lab)
    return torch.autograd.grad(l, [a, b]) + (l,)


def cpu_grads(xL, xR):
    a, b = xL.clone().requires_grad_(True), xR.clone().requires_grad_(True)
    out = ref(a, b, *calib[:3], calibs_Proj_R=calib[3])
    loss = R.attack_loss(cfg_r, out, pair["disp_L"], labels)
    return torch.autograd.grad(loss, [a, b]) + (loss,)


for bb, split in (("b2", 1), ("b2", 3), ("b2", 0), ("cudnn", 1)):
    dsgn.set_backbone_impl(bb); ops.set_conv2d_split(split)
    print("=== backbone %s split %d" % (bb, split))
    xL, xR = pair["imgL"].cuda().clone(), pair["imgR"].cuda().clone()
    cL, cR = A.denormalize(pair["imgL"]).cuda(), A.denormalize(pair["imgR"]).cuda()
    for it in range(K):
        gL, gR, l = grads(xL, xR)
        gL2, gR2, _ = grads(xL, xR)
        gLc, gRc, lc = cpu_grads(xL.cpu(), xR.cpu())
        print(" iter %d: loss gpu %.6f cpu(on the gpu iterate) %.6f | grad rel vs cpu L %.2e R %.2e | repeat bitwise %s | sign agree all px L %.5f R %.5f" % (
            it, l.item(), lc.item(), rel_err(gL.cpu(), gLc), rel_err(gR.cpu(), gRc), torch.equal(gL, gL2) and torch.equal(gR, gR2),
            (gL.cpu().sign() == gLc.sign()).float().mean().item(), (gR.cpu().sign() == gRc.sign()).float().mean().item()))
        nL_ref = A.pgd_step_linf(xL.cpu(), gL.cpu(), cL.cpu(), alpha, eps)
        nR_ref = A.pgd_step_linf(xR.cpu(), gR.cpu(), cR.cpu(), alpha, eps)
        xL, xR = attack.pgd_step_pair(xL, gL.contiguous(), cL, xR, gR.contiguous(), cR, alpha, eps, inplace=True)
        print("          update kernel vs oracle rule on the same gradient: L %s R %s" % (torch.equal(xL.cpu(), nL_ref), torch.equal(xR.cpu(), nR_ref)))
dsgn.set_backbone_impl("b2"); ops.set_conv2d_split(1)
