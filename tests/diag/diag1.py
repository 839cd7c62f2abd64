"""GPU diagnostics (scratch): tcgen05 conv vs SIMT conv, pyramid roi per level, e2e backward per stage."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
from eval_driving_safety_b200 import ops, dsgn, synthetic, stereo_rcnn
from oracle import dsgn_ref as R, attack_ref as A
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import rel_err, max_err

what = sys.argv[1:] or ["conv", "roi", "e2e"]
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

if "conv" in what:
    cases = [(64, 64, 1, False, (1, 16, 8)), (64, 64, 1, False, (4, 6, 8)), (64, 64, 1, False, (3, 20, 19)),
             (96, 64, 1, False, (4, 4, 8)), (64, 96, 1, False, (2, 5, 9)), (128, 128, 1, False, (2, 4, 6)),
             (64, 128, 2, False, (4, 6, 8)), (128, 128, 2, False, (4, 4, 4)), (64, 128, 2, False, (6, 36, 20)),
             (128, 128, 2, True, (2, 3, 4)), (128, 64, 2, True, (2, 2, 6)), (128, 64, 2, True, (3, 18, 10))]
    for cin, cout, stride, tr, sp in cases:
        g = torch.Generator().manual_seed(cin + cout)
        x = torch.randn(2, cin, *sp, generator=g).cuda()
        w = (torch.randn((cin, cout, 3, 3, 3) if tr else (cout, cin, 3, 3, 3), generator=g) * 0.05).cuda()
        ref = ops.conv3d(x, w, stride, tr, impl=1)
        torch.cuda.synchronize()
        try:
            out = ops.conv3d(x, w, stride, tr, impl=0)
            torch.cuda.synchronize()
            print("conv", cin, cout, stride, tr, sp, "rel_err tcgen05 vs simt = %.3e  max %.3e" % (rel_err(out, ref), max_err(out, ref)), flush=True)
        except Exception as e:
            print("conv", cin, cout, stride, tr, sp, "FAILED:", e, flush=True)
            break

if "roi" in what:
    g = torch.Generator().manual_seed(22)
    sizes = [(150, 497), (75, 249), (38, 125), (19, 63)]
    feats = [torch.randn(1, 8, h, w, generator=g) for h, w in sizes]
    x1, y1 = torch.rand(40, generator=g) * 1500, torch.rand(40, generator=g) * 400
    w = torch.exp(torch.rand(40, generator=g) * 6.0) + 2
    h = torch.exp(torch.rand(40, generator=g) * 5.0) + 2
    rois = torch.stack([torch.zeros(40), x1, y1, x1 + w, y1 + h], 1)
    print("levels cpu", A.roi_levels(rois).tolist())
    print("levels gpu", stereo_rcnn.roi_levels(rois.cuda()).tolist())
    from torchvision.ops import roi_align
    for i, f in enumerate(feats):
        scale = f.size(2) / 600.0
        ref = roi_align(f, rois, (7, 7), scale, 0, False)
        out = ops.roi_align(f.cuda(), rois.cuda(), 7, scale).cpu()
        err = (out - ref).abs().amax((1, 2, 3))
        print("level", i, "max err per roi", [round(v, 5) for v in err.tolist() if v > 1e-4], "widths", [round(float(w[j]), 1) for j in range(40) if err[j] > 1e-4])

if "e2e" in what:
    H, W = 32, 64
    cfg_r, cfg_p = R.tiny_cfg(), dsgn.tiny_cfg()
    ref = R.build_model(cfg_r, seed=1)
    if "affine" in what:
        ga_ = torch.Generator().manual_seed(5)
        for m in ref.modules():
            if isinstance(m, torch.nn.GroupNorm):
                m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=ga_)
                m.bias.data = 0.1 * torch.randn(m.bias.shape, generator=ga_)
    model = dsgn.StereoNet(cfg_p); model.load_state_dict(ref.state_dict()); model = model.freeze().cuda()
    pair = synthetic.make_pair(0, H, W, max_depth=8.4)
    calib = synthetic.make_calib(1, scale=H / 384, cu=W / 2, cv=H / 2)
    ops.set_conv_impl(1)
    g = torch.Generator().manual_seed(3)
    fL, fR = torch.randn(1, 32, 8, 16, generator=g), torch.randn(1, 32, 8, 16, generator=g)
    rL = torch.randn(1, 32, 8, 16, generator=g)
    # psv stage backward
    a, b = fL.clone().requires_grad_(True), fR.clone().requires_grad_(True)
    cost_r, out_r, cost1_r = ref.psv_stage(a, b, calib[0], calib[1])
    go, g1 = torch.randn(out_r.shape, generator=g), torch.randn(cost1_r.shape, generator=g)
    ga_r, gb_r = torch.autograd.grad((out_r * go).sum() + (cost1_r * g1).sum(), [a, b])
    ac, bc = fL.cuda().requires_grad_(True), fR.cuda().requires_grad_(True)
    cost_g, out_g, cost1_g = model.psv_stage(ac, bc, calib[0], calib[1])
    ga, gb = torch.autograd.grad((out_g * go.cuda()).sum() + (cost1_g * g1.cuda()).sum(), [ac, bc])
    print("psv fwd out %.2e cost1 %.2e | bwd gL %.2e gR %.2e" % (rel_err(out_g.cpu(), out_r), rel_err(cost1_g.cpu(), cost1_r), rel_err(ga.cpu(), ga_r), rel_err(gb.cpu(), gb_r)))
    # lift + bev backward
    o, r2 = out_r.detach().clone().requires_grad_(True), rL.clone().requires_grad_(True)
    heads_r = ref.bev_stage(ref.lift(o, r2, calib[2]))
    gh = [torch.randn(t.shape, generator=g) for t in heads_r]
    go_r, gr_r = torch.autograd.grad(sum((t * q).sum() for t, q in zip(heads_r, gh)), [o, r2])
    oc, rc = out_r.detach().cuda().requires_grad_(True), rL.cuda().requires_grad_(True)
    heads_g = model.bev_stage(model.lift(oc, rc, calib[2]))
    go_g, gr_g = torch.autograd.grad(sum((t * q.cuda()).sum() for t, q in zip(heads_g, gh)), [oc, rc])
    print("bev fwd %s | bwd g_psv %.2e g_img %.2e" % (["%.2e" % rel_err(x.cpu(), y) for x, y in zip(heads_g, heads_r)], rel_err(go_g.cpu(), go_r), rel_err(gr_g.cpu(), gr_r)))
    # depth head backward (torch ops both sides)
    c1 = cost1_r.detach().clone().requires_grad_(True)
    d_r = ref.depth_head(c1, (H, W)); gd = torch.randn(d_r.shape, generator=g)
    (gc_r,) = torch.autograd.grad((d_r * gd).sum(), c1)
    c1c = cost1_r.detach().cuda().requires_grad_(True)
    d_g = model.depth_head(c1c, (H, W))
    (gc_g,) = torch.autograd.grad((d_g * gd.cuda()).sum(), c1c)
    print("depth head fwd %.2e bwd %.2e" % (rel_err(d_g.cpu(), d_r), rel_err(gc_g.cpu(), gc_r)))
    # feature extraction backward
    xi = pair['imgL'].clone().requires_grad_(True)
    f_r, p_r = ref.feature_extraction(xi); gf = torch.randn(f_r.shape, generator=g)
    (gx_r,) = torch.autograd.grad((f_r * gf).sum() + (p_r * gf).sum(), xi)
    xc = pair['imgL'].cuda().requires_grad_(True)
    f_g, p_g = model.feature_extraction(xc)
    (gx_g,) = torch.autograd.grad((f_g * gf.cuda()).sum() + (p_g * gf.cuda()).sum(), xc)
    print("feat fwd %.2e bwd %.2e" % (rel_err(f_g.cpu(), f_r), rel_err(gx_g.cpu(), gx_r)))
    # full model
    labels = R.make_labels(cfg_r, 1, 7)
    xL, xR = pair['imgL'].clone().requires_grad_(True), pair['imgR'].clone().requires_grad_(True)
    out = ref(xL, xR, *calib[:3], calibs_Proj_R=calib[3]); loss = R.attack_loss(cfg_r, out, pair['disp_L'], labels)
    gL_r, gR_r = torch.autograd.grad(loss, [xL, xR])
    xLc, xRc = pair['imgL'].cuda().requires_grad_(True), pair['imgR'].cuda().requires_grad_(True)
    outg = model(xLc, xRc, *calib[:3], calibs_Proj_R=calib[3])
    lossg = dsgn.attack_loss(cfg_p, outg, pair['disp_L'].cuda(), {k: v.cuda() for k, v in labels.items()})
    gL_g, gR_g = torch.autograd.grad(lossg, [xLc, xRc])
    print("full: loss %.6f vs %.6f gL %.2e gR %.2e" % (lossg.item(), loss.item(), rel_err(gL_g.cpu(), gL_r), rel_err(gR_g.cpu(), gR_r)))
    for frac in (1e-1, 1e-2, 1e-3):
        tau = frac * gL_r.abs().max()
        big = gL_r.abs() > tau
        print("  sign agree where |g| > %.0e*max: %.5f (%.3f of pixels)" % (frac, (gL_g.cpu().sign() == gL_r.sign())[big].float().mean().item(), big.float().mean().item()))
    print("  |g| quantiles", torch.quantile(gL_r.abs().flatten(), torch.tensor([0.01, 0.1, 0.5, 0.9, 0.99, 1.0])).tolist())

if "bev" in what:
    H, W = 32, 64
    cfg_r, cfg_p = R.tiny_cfg(), dsgn.tiny_cfg()
    ref = R.build_model(cfg_r, seed=1)
    g = torch.Generator().manual_seed(5)
    if "affine" in what:
        for m in ref.modules():
            if isinstance(m, torch.nn.GroupNorm):
                m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
                m.bias.data = 0.1 * torch.randn(m.bias.shape, generator=g)
    model = dsgn.StereoNet(cfg_p); model.load_state_dict(ref.state_dict()); model = model.freeze().cuda()
    ops.set_conv_impl(1)
    vox = torch.randn(1, 96, 16, 4, 16, generator=g)

    def cmp(name, fr, fg, x):
        a = x.clone().requires_grad_(True); ya = fr(a); gy = torch.randn(ya.shape, generator=g)
        (ga,) = torch.autograd.grad((ya * gy).sum(), a)
        b = x.cuda().requires_grad_(True); yb = fg(b)
        (gb,) = torch.autograd.grad((yb * gy.cuda()).sum(), b)
        print("%-28s fwd %.2e bwd %.2e" % (name, rel_err(yb.cpu(), ya), rel_err(gb.cpu(), ga)), flush=True)
        return ya.detach()

    v1 = cmp("rpn3d_conv(96->64)+gn+relu", lambda t: ref.rpn3d_conv(t), lambda t: dsgn.run_convbn_3d(model.rpn3d_conv[0], t, relu=True), vox)
    hg_r, hg_g = ref.rpn3d_hg, model.rpn3d_hg
    o1 = cmp("hg.conv1 s2 64->128", lambda t: hg_r.conv1(t), lambda t: dsgn.run_convbn_3d(hg_g.conv1[0], t, relu=True), v1)
    pre = cmp("hg.conv2 128->128", lambda t: F.relu(hg_r.conv2(t)), lambda t: dsgn.run_convbn_3d(hg_g.conv2, t, relu=True), o1)
    o3 = cmp("hg.conv3 s2", lambda t: hg_r.conv3(t), lambda t: dsgn.run_convbn_3d(hg_g.conv3[0], t, relu=True), pre)
    o4 = cmp("hg.conv4", lambda t: hg_r.conv4(t), lambda t: dsgn.run_convbn_3d(hg_g.conv4[0], t, relu=True), o3)
    post = cmp("hg.conv5 deconv+res+relu", lambda t: F.relu(hg_r.conv5(t) + pre), lambda t: dsgn.run_convbn_3d(hg_g.conv5, t, relu=True, res=pre.cuda()), o4)
    o6 = cmp("hg.conv6 deconv+res", lambda t: hg_r.conv6(t) + v1, lambda t: dsgn.run_convbn_3d(hg_g.conv6, t, relu=False, res=v1.cuda()), post)
    cmp("whole hg + res", lambda t: hg_r(t) + t, lambda t: hg_g(t, res=t), v1)

    def tail_r(t):
        v = F.avg_pool3d(t, (1, cfg_r.y_pool, 1)); n, c, zz, yy, xx = v.shape
        bev = ref.bev_conv(v.permute(0, 1, 3, 2, 4).reshape(n, c * yy, zz, xx))
        return torch.cat([ref.bbox_cls(bev), ref.bbox_reg(bev), ref.bbox_centerness(bev)], 1)

    def tail_g(t):
        v = F.avg_pool3d(t, (1, cfg_p.y_pool, 1)); n, c, zz, yy, xx = v.shape
        bev = model.bev_conv(v.permute(0, 1, 3, 2, 4).reshape(n, c * yy, zz, xx))
        return torch.cat([model.bbox_cls(bev), model.bbox_reg(bev), model.bbox_centerness(bev)], 1)
    cmp("avgpool + bev 2d heads", tail_r, tail_g, o6)
    cmp("avgpool + bev (cl3 input)", tail_r, lambda t: tail_g(t.contiguous(memory_format=torch.channels_last_3d)), o6)

if "pyr" in what:
    g = torch.Generator().manual_seed(22)
    sizes = [(150, 497), (75, 249), (38, 125), (19, 63)]
    feats = [torch.randn(1, 8, h, w, generator=g) for h, w in sizes]
    x1, y1 = torch.rand(40, generator=g) * 1500, torch.rand(40, generator=g) * 400
    w = torch.exp(torch.rand(40, generator=g) * 6.0) + 2
    h = torch.exp(torch.rand(40, generator=g) * 5.0) + 2
    rois = torch.stack([torch.zeros(40), x1, y1, x1 + w, y1 + h], 1)
    refp = A.pyramid_roi_feat(feats, rois, 600.0, 7)
    outp = stereo_rcnn.pyramid_roi_feat([f.cuda() for f in feats], rois.cuda(), 600.0, 7).cpu()
    print("pyramid err per roi", [(i, round(v, 4)) for i, v in enumerate((outp - refp).abs().amax((1, 2, 3)).tolist()) if v > 1e-4])
