"""GPU diagnostics (scratch): is the affine-GN gradient gap conditioning or a bug?  fp64 CPU oracle as arbiter."""
import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
from eval_driving_safety_b200 import ops, dsgn, synthetic
from oracle import dsgn_ref as R
from helpers import rel_err, max_err
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
H, W = 32, 64
cfg_r, cfg_p = R.tiny_cfg(), dsgn.tiny_cfg()
pair = synthetic.make_pair(0, H, W, max_depth=8.4)
calib = synthetic.make_calib(1, scale=H / 384, cu=W / 2, cv=H / 2)
labels = R.make_labels(cfg_r, 1, 7)
for affine in (False, True):
    ref = R.build_model(cfg_r, seed=1)
    g = torch.Generator().manual_seed(5)
    if affine:
        for m in ref.modules():
            if isinstance(m, torch.nn.GroupNorm):
                m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
                m.bias.data = 0.1 * torch.randn(m.bias.shape, generator=g)
    model = dsgn.StereoNet(cfg_p); model.load_state_dict(ref.state_dict()); model = model.freeze().cuda()
    ref64 = copy.deepcopy(ref).double()

    def cpu_grads(m, dt):
        xL, xR = pair['imgL'].to(dt).requires_grad_(True), pair['imgR'].to(dt).requires_grad_(True)
        # dsgn_ref helpers build fp32 grids/depths; run them in the model dtype by casting calibration
        out = m(xL, xR, *calib[:3], calibs_Proj_R=calib[3])
        loss = R.attack_loss(cfg_r, out, pair['disp_L'].to(dt), {k: v.to(dt) for k, v in labels.items()})
        return torch.autograd.grad(loss, [xL, xR]) + (loss,)
    g32 = cpu_grads(ref, torch.float32)
    try:
        g64 = cpu_grads(ref64, torch.float64)
    except Exception as e:
        print("fp64 oracle failed:", e); g64 = None
    for impl in (1, 0):
        ops.set_conv_impl(impl)
        xLc, xRc = pair['imgL'].cuda().requires_grad_(True), pair['imgR'].cuda().requires_grad_(True)
        outg = model(xLc, xRc, *calib[:3], calibs_Proj_R=calib[3])
        lossg = dsgn.attack_loss(cfg_p, outg, pair['disp_L'].cuda(), {k: v.cuda() for k, v in labels.items()})
        gg = torch.autograd.grad(lossg, [xLc, xRc])
        msg = "affine=%s impl=%d: gpu vs cpu32 gL %.2e" % (affine, impl, rel_err(gg[0].cpu(), g32[0]))
        if g64 is not None:
            msg += " | gpu vs cpu64 %.2e | cpu32 vs cpu64 %.2e" % (rel_err(gg[0].cpu(), g64[0]), rel_err(g32[0], g64[0]))
        for frac in (1e-1, 1e-2, 1e-3):
            refg = g64[0] if g64 is not None else g32[0]
            big = refg.abs() > frac * refg.abs().max()
            msg += " | sign@%.0e %.4f" % (frac, (gg[0].cpu().sign() == refg.sign())[big].float().mean().item())
        print(msg, flush=True)
