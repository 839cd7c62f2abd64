import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import rel_err
from oracle import attack_ref as A, dsgn_ref as R
from eval_driving_safety_b200 import attack, dsgn, ops, synthetic
torch.set_num_threads(os.cpu_count() or 1)
cfg_r, cfg_p = R.default_cfg(), dsgn.default_cfg()
ref = R.build_model(cfg_r, seed=1)
for p in ref.parameters(): p.requires_grad_(False)
model = dsgn.StereoNet(cfg_p); model.load_state_dict(ref.state_dict()); model = model.freeze().cuda()
pair = synthetic.make_pair(0); calib = synthetic.make_calib(1); labels = R.make_labels(cfg_r, 1, 7)
xL, xR = pair["imgL"].clone().requires_grad_(True), pair["imgR"].clone().requires_grad_(True)
out_r = ref(xL, xR, *calib[:3], calibs_Proj_R=calib[3])
loss_r = R.attack_loss(cfg_r, out_r, pair["disp_L"], labels)
gL_r, gR_r = torch.autograd.grad(loss_r, [xL, xR])
print("cpu loss", loss_r.item(), "|g| max %.3e median %.3e" % (gL_r.abs().max().item(), gL_r.abs().median().item()), flush=True)
lab = {k: v.cuda() for k, v in labels.items()}
adv_r = A.pgd_step_linf(pair["imgL"], gL_r, A.denormalize(pair["imgL"]), 8 / 255, 8 / 255)
clean = pair["imgL"].cuda() * torch.tensor(A.IMAGENET_STD).view(1, 3, 1, 1).cuda() + torch.tensor(A.IMAGENET_MEAN).view(1, 3, 1, 1).cuda()
for impl, tf32, bb, split, split_bwd in ((0, False, 'b2', 1, 0), (0, False, 'b2', 1, 1), (0, False, 'b2', 0, 0)):
    ops.set_conv_impl(impl); dsgn.set_backbone_impl(bb); ops.set_conv2d_split(split); ops.CONV2D_SPLIT_BWD = split_bwd
    torch.backends.cudnn.allow_tf32 = tf32; torch.backends.cuda.matmul.allow_tf32 = tf32
    a, b = pair["imgL"].cuda().requires_grad_(True), pair["imgR"].cuda().requires_grad_(True)
    out = model(a, b, *calib[:3], calibs_Proj_R=calib[3])
    loss = dsgn.attack_loss(cfg_p, out, pair["disp_L"].cuda(), lab)
    gL, gR = torch.autograd.grad(loss, [a, b])
    big = gL_r.abs() > 1e-2 * gL_r.abs().max()
    adv = attack.pgd_step(pair["imgL"].cuda(), gL.contiguous(), clean, 8 / 255, 8 / 255)
    same = ((adv.cpu() - adv_r).abs() < 1e-5).float().mean().item()
    print("split_bwd=%s" % split_bwd, end=" ")
    print("impl=%d cudnn_tf32=%s backbone=%s split=%s same_px %.5f: depth %.2e cls %.2e reg %.2e loss %.2e gradL %.2e gradR %.2e sign@1%%max %.5f sign(all) %.5f" % (
        impl, tf32, bb, split, same, rel_err(out["depth_preds"].cpu(), out_r["depth_preds"]), rel_err(out["bbox_cls"].cpu(), out_r["bbox_cls"]),
        rel_err(out["bbox_reg"].cpu(), out_r["bbox_reg"]),
        abs(loss.item() - loss_r.item()) / abs(loss_r.item()), rel_err(gL.cpu(), gL_r), rel_err(gR.cpu(), gR_r),
        (gL.cpu().sign() == gL_r.sign())[big].float().mean().item(), (gL.cpu().sign() == gL_r.sign()).float().mean().item()), flush=True)
    del out, loss, gL, gR
