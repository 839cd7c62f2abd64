"""Why does the 3-iteration tiny-config PGD differ?  Teacher-forced per-iteration gradient comparison
(GPU path fed the ORACLE's iterate) for each 2-D backbone implementation."""
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
xL, xR = pair["imgL"].clone(), pair["imgR"].clone()
cleanL, cleanR = A.denormalize(xL), A.denormalize(xR)
for it in range(K):
    a, b = xL.clone().requires_grad_(True), xR.clone().requires_grad_(True)
    out = ref(a, b, *calib[:3], calibs_Proj_R=calib[3])
    loss = R.attack_loss(cfg_r, out, pair["disp_L"], labels)
    gL, gR = torch.autograd.grad(loss, [a, b])
    print("iter %d oracle loss %.6f |g|max %.3e  frac |g| < 1e-3 max: %.3f  exact zeros %.3f" % (
        it, loss.item(), gL.abs().max().item(), (gL.abs() < 1e-3 * gL.abs().max()).float().mean().item(), (gL == 0).float().mean().item()))
    for bb in ("b2", "cudnn"):
        dsgn.set_backbone_impl(bb)
        ac, bc = xL.cuda().requires_grad_(True), xR.cuda().requires_grad_(True)
        o = model(ac, bc, *calib[:3], calibs_Proj_R=calib[3])
        l = dsgn.attack_loss(cfg_p, o, pair["disp_L"].cuda(), lab)
        g1, g2 = torch.autograd.grad(l, [ac, bc])
        print("   %-6s loss %.6f  gradL rel %.2e gradR rel %.2e  sign agree all px L %.4f R %.4f" % (
            bb, l.item(), rel_err(g1.cpu(), gL), rel_err(g2.cpu(), gR), (g1.cpu().sign() == gL.sign()).float().mean().item(),
            (g2.cpu().sign() == gR.sign()).float().mean().item()))
    xL = A.pgd_step_linf(xL, gL, cleanL, alpha, eps)
    xR = A.pgd_step_linf(xR, gR, cleanR, alpha, eps)
# free-running product loop, both backbones
disp = pair["disp_L"].cuda()
loss_fn = lambda out: dsgn.attack_loss(cfg_p, out, disp, lab)
for bb in ("b2", "cudnn"):
    dsgn.set_backbone_impl(bb)
    aL, aR, losses = attack.pgd_attack(model, loss_fn, pair["imgL"].cuda(), pair["imgR"].cuda(), calib, K, alpha, eps)
    dL_ref, dL = A.denormalize(xL) - cleanL, A.denormalize(aL.cpu()) - cleanL
    print("free-running %s: same %.4f  losses %s" % (bb, ((dL - dL_ref).abs() < 1e-6).float().mean().item(), losses.tolist()))
dsgn.set_backbone_impl("b2")
