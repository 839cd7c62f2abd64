"""GPU tests of the round-2 host-side additions, all through the C ABI: module swapping on the oracle's own
model, the graph-captured universal-patch iteration (device-resident patch positions, split update for the
all-reduce), the Stereo R-CNN patch loop, per-calibration graphs, the asynchronous image writer, and real
multi-rank sharding invariance when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import max_err, rel_err
from oracle import attack_ref as A
from oracle import dsgn_ref as R

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H, W = 32, 64


@pytest.fixture(scope="module")
def tiny(built_lib):
    from eval_driving_safety_b200 import dsgn, ops, synthetic
    ops.set_conv_impl(1)
    cfg_r, cfg_p = R.tiny_cfg(), dsgn.tiny_cfg()
    ref = R.build_model(cfg_r, seed=1)
    model = dsgn.StereoNet(cfg_p)
    model.load_state_dict(ref.state_dict())
    model = model.freeze().cuda()
    pair = synthetic.make_pair(0, H, W, max_depth=8.4)
    calib = synthetic.make_calib(1, scale=H / 384, cu=W / 2, cv=H / 2)
    labels = R.make_labels(cfg_r, 1, 7)
    yield dict(cfg_r=cfg_r, cfg_p=cfg_p, ref=ref, model=model, pair=pair, calib=calib, labels=labels,
               labels_gpu={k: v.cuda() for k, v in labels.items()})
    ops.set_conv_impl(0)


def _ref_grads(s, xL, xR, labels=None):
    xL, xR = xL.clone().requires_grad_(True), xR.clone().requires_grad_(True)
    out = s["ref"](xL, xR, *s["calib"][:3], calibs_Proj_R=s["calib"][3])
    loss = R.attack_loss(s["cfg_r"], out, s["pair"]["disp_L"], labels or s["labels"])
    gL, gR = torch.autograd.grad(loss, [xL, xR])
    return out, loss, gL, gR


# ---------------------------------------------------------------- module seams (ADVICE)
def test_swap_modules_on_the_oracle_model_reproduces_the_product_model(tiny, monkeypatch):
    """``swap_modules(upstream model)`` puts every Conv3d / ConvTranspose3d / Conv2d / GroupNorm
    on the sm_100a kernels; the oracle's StereoNetRef (stock cost volume / grid_sample around them) then has to
    give the oracle's outputs and input gradient, and agree with ``dsgn.StereoNet``."""
    import copy
    from eval_driving_safety_b200 import dsgn, modules, ops
    s = tiny
    ops.set_conv_impl(1)
    out_r, loss_r, gL_r, gR_r = _ref_grads(s, s["pair"]["imgL"], s["pair"]["imgR"])       # the CPU oracle itself
    # the oracle builds its calibration-only tensors on the CPU: move them along for the CUDA run of its forward
    for name in ("plane_shifts", "lifting_grid", "full_depths"):
        orig = getattr(R, name)
        monkeypatch.setattr(R, name, (lambda f: lambda *a, **k: f(*a, **k).cuda())(orig))
    swapped = modules.swap_modules(copy.deepcopy(s["ref"])).cuda()
    assert sum(isinstance(m, modules.Conv3dSm100) for m in swapped.modules()) == 15
    assert not any(type(m) in (torch.nn.Conv3d, torch.nn.ConvTranspose3d, torch.nn.Conv2d, torch.nn.GroupNorm)
                   for m in swapped.modules())
    a, b = s["pair"]["imgL"].cuda().requires_grad_(True), s["pair"]["imgR"].cuda().requires_grad_(True)
    out_s = swapped(a, b, *s["calib"][:3], calibs_Proj_R=s["calib"][3])
    loss_s = R.attack_loss(s["cfg_r"], out_s, s["pair"]["disp_L"].cuda(), s["labels_gpu"])
    gL_s, gR_s = torch.autograd.grad(loss_s, [a, b])
    a2, b2 = s["pair"]["imgL"].cuda().requires_grad_(True), s["pair"]["imgR"].cuda().requires_grad_(True)
    out_p = s["model"](a2, b2, *s["calib"][:3], calibs_Proj_R=s["calib"][3])
    loss_p = dsgn.attack_loss(s["cfg_p"], out_p, s["pair"]["disp_L"].cuda(), s["labels_gpu"])
    gL_p, gR_p = torch.autograd.grad(loss_p, [a2, b2])
    for k in ("depth_preds", "bbox_cls", "bbox_reg", "bbox_centerness"):
        assert rel_err(out_s[k].cpu(), out_r[k]) < 1e-4, k                   # vs the oracle
        assert rel_err(out_s[k], out_p[k]) < 1e-4, k                          # vs the product model
    assert abs(loss_s.item() - loss_r.item()) < 1e-4 * abs(loss_r.item())
    assert rel_err(gL_s.cpu(), gL_r) < 5e-3 and rel_err(gR_s.cpu(), gR_r) < 5e-3
    assert rel_err(gL_s, gL_p) < 5e-3 and rel_err(gR_s, gR_p) < 5e-3


def test_dropins_reject_unsupported_configurations_at_call_time():
    from eval_driving_safety_b200 import modules as M
    x = torch.zeros(1, 32, 4, 8, 8, device="cuda")
    with pytest.raises(RuntimeError):
        M.Conv3dSm100(32, 32, 3, 1, 1, bias=True).cuda()(x)
    with pytest.raises(RuntimeError):
        M.Conv3dSm100(32, 32, 3, 1, 0, bias=False).cuda()(x)
    with pytest.raises(RuntimeError):
        M.ConvTranspose3dSm100(32, 32, 3, 2, 1, output_padding=0, bias=False).cuda()(x)


# ---------------------------------------------------------------- universal patch: kernels
def test_patch_kernels_with_device_resident_centres_and_split_update():
    from eval_driving_safety_b200 import attack
    g = torch.Generator().manual_seed(4)
    h, w, r = 96, 160, 9
    img = torch.randn(1, 3, h, w, generator=g).cuda()
    patch = torch.randn(1, 3, 2 * r + 1, 2 * r + 1, generator=g).cuda()
    gl, gr = torch.randn(1, 3, h, w, generator=g).cuda() * 1e-4, torch.randn(1, 3, h, w, generator=g).cuda() * 1e-4
    cl, cr = [40, 80], [40, 16]
    i1 = attack.patch_apply(img.clone(), patch, cl, r)
    i2 = attack.patch_apply(img.clone(), patch, torch.tensor(cl, dtype=torch.int32, device="cuda"), r)
    assert torch.equal(i1, i2)
    assert torch.equal(i1.cpu(), A.patch_apply(img.cpu(), patch.cpu(), cl, r))
    for lo, hi in ((None, None), ([-0.5, -0.4, -0.3], [0.5, 0.4, 0.3])):
        p1 = attack.patch_update(patch.clone(), gl, gr, cl, cr, r, 1e3, 8 / 255, lo, hi)
        p2 = attack.patch_update(patch.clone(), gl, gr, torch.tensor(cl + cr, dtype=torch.int32, device="cuda"), None, r,
                                 1e3, 8 / 255, lo, hi)
        delta = torch.empty_like(patch)
        p3 = patch.clone()
        attack.patch_update(p3, gl, gr, cl, cr, r, 1e3, 8 / 255, delta_out=delta)
        assert torch.equal(p3, patch)                                          # delta mode leaves the patch alone
        attack.patch_axpy(p3, delta, lo, hi)
        assert torch.equal(p1, p2) and torch.equal(p1, p3)
        assert torch.equal(p1.cpu(), A.patch_update(patch.cpu(), gl.cpu(), gr.cpu(), cl, cr, r, 1e3, 8 / 255, lo, hi))
    # a box that leaves the frame: the host-centre entry point refuses, the device-centre one contributes nothing
    with pytest.raises(RuntimeError):
        attack.patch_update(patch.clone(), gl, gr, [2, 80], cr, r, 1e3, 8 / 255)
    delta = torch.ones_like(patch)
    attack.patch_update(patch.clone(), gl, gr, torch.tensor([2, 80] + cr, dtype=torch.int32, device="cuda"), None, r, 1e3,
                        8 / 255, delta_out=delta)
    assert delta.abs().max() == 0


def test_patch_iteration_graph_targeted_attack(tiny):
    """attack/DSGN/patch_attack.py:336-430 on the tiny model: fake ground truth -> labels, blend, forward/backward,
    crop, clipped descent.  The CUDA-graph engine (device-resident centres, one graph for every image) must equal
    the eager loop bit for bit, the split update with a (world = 1) all-reduce hook must equal the fused one, and
    both must follow the oracle's loop."""
    from eval_driving_safety_b200 import attack, dsgn, engine, parallel, synthetic
    s = tiny
    radius, iters, alpha, eps = 3, 2, 1e3, 8 / 255
    dim = 2 * radius + 1
    bbox, box3d = synthetic.make_targets(4, 1)
    attack.inject_fake_gt(bbox, box3d)
    box3d[0, 3:6] = torch.tensor([0.3, 0.5, 5.1])                      # inside the tiny world grid
    labels = synthetic.labels_from_box3d(s["cfg_p"], box3d)
    assert labels["cls"].sum() > 0
    labels_gpu = {k: v.cuda() for k, v in labels.items()}
    g = torch.Generator().manual_seed(9)
    patch0 = torch.randn(1, 3, dim, dim, generator=g) * 0.5
    images = [(synthetic.make_pair(i, H, W, max_depth=8.4), [14 + i, 40 - i], [14 + i, 30 - i]) for i in range(2)]
    # oracle: sequential over the images, ``iters`` iterations each.  Before every update the oracle's state is also given
    # to the product path for ONE update (teacher-forced): on this chaotic 32x64 network only single steps from a common
    # state are comparable (alpha = 1e3 saturates the clip, an entry is -+eps * sign(g): one flipped near-zero gradient moves
    # it by 2 eps and the trajectories part -- measured 52-80 % identical entries after 4 free-running updates, depending on
    # nothing but the summation order inside the 2-D convs)
    loss_of = lambda disp: (lambda out: dsgn.attack_loss(s["cfg_p"], out, disp, labels_gpu))
    patch_r, step_same = patch0.clone(), []
    for pair, cl, cr in images:
        imgL, imgR = pair["imgL"].clone(), pair["imgR"].clone()
        for _ in range(iters):
            pg = patch_r.clone().cuda()
            attack.patch_attack_step(s["model"], loss_of(pair["disp_L"].cuda()), imgL.cuda(), imgR.cuda(), s["calib"], pg, cl, cr,
                                     radius, iters=1, alpha=alpha, eps=eps)
            imgL, imgR = A.patch_apply(imgL, patch_r, cl, radius), A.patch_apply(imgR, patch_r, cr, radius)
            a, b = imgL.clone().requires_grad_(True), imgR.clone().requires_grad_(True)
            out = s["ref"](a, b, *s["calib"][:3], calibs_Proj_R=s["calib"][3])
            gL, gR = torch.autograd.grad(R.attack_loss(s["cfg_r"], out, pair["disp_L"], labels), [a, b])
            patch_r = A.patch_update(patch_r, gL, gR, cl, cr, radius, alpha, eps)
            assert (pg.cpu() - patch_r).abs().max() <= 2 * eps + 1e-6
            step_same.append(((pg.cpu() - patch_r).abs() < 1e-5).float().mean().item())
    assert min(step_same) > 0.9, step_same
    results = []
    for use_graph, hook in ((False, None), (True, None), (True, parallel.allreduce_patch_delta)):
        patch = patch0.clone().cuda()
        ex = (images[0][0]["imgL"].cuda(), images[0][0]["imgR"].cuda(), images[0][0]["disp_L"].cuda())
        eng = engine.PatchIterationGraph(s["model"], s["cfg_p"], labels_gpu, s["calib"], patch, radius, ex, alpha=alpha,
                                         eps=eps, use_graph=use_graph, allreduce=hook)
        assert torch.equal(patch.cpu(), patch0)                          # capture / warm-up did not train the patch
        for pair, cl, cr in images:
            eng.load(pair["imgL"].cuda(), pair["imgR"].cuda(), pair["disp_L"].cuda(), cl, cr)
            for _ in range(iters):
                loss = eng.iterate()
        torch.cuda.synchronize()
        assert torch.isfinite(loss)
        results.append(patch.cpu().clone())
    # the graph engine (device-resident centres) == the eager loop == the split update behind an all-reduce hook, bit for bit
    assert torch.equal(results[0], results[1]) and torch.equal(results[0], results[2])
    assert (results[0] - patch_r).abs().max() <= 2 * eps * iters * len(images) + 1e-6
    with pytest.raises(RuntimeError):
        eng.load(ex[0], ex[1], ex[2], [1, 40], [1, 30])                  # box leaves the frame


def test_runner_patch_resumes_and_saves(built_lib, tmp_path):
    """runner patch: fake ground truth, epoch0 resume (attack/DSGN/patch_attack.py:211-234), save :438-443."""
    import numpy as np
    from eval_driving_safety_b200 import runner
    patch = runner.main(["patch", "--tiny", "--pairs", "2", "--epochs", "1", "--iter", "2", "--ratio", "0.2",
                         "--save-dir", str(tmp_path)])
    assert patch.shape == (1, 3, 7, 7) and 0 < patch.abs().max() <= 2 * 2 * (8 / 255) + 1e-6
    assert (tmp_path / "epoch0" / "patch.npy").exists() and (tmp_path / "epoch1" / "patch.npy").exists()
    saved = np.load(tmp_path / "epoch1" / "patch.npy")
    assert np.array_equal(saved, patch.cpu().numpy())
    # second run resumes from epoch0: seed it with the trained patch and check the start value is used
    np.save(tmp_path / "epoch0" / "patch.npy", saved)
    patch2 = runner.main(["patch", "--tiny", "--pairs", "2", "--epochs", "1", "--iter", "2", "--ratio", "0.2",
                          "--save-dir", str(tmp_path)])
    assert not torch.equal(patch2.cpu(), patch.cpu())
    assert (patch2.cpu() - torch.from_numpy(saved)).abs().max() <= 2 * 2 * (8 / 255) + 1e-6


# ---------------------------------------------------------------- Stereo R-CNN patch loop
def test_stereo_rcnn_patch_loop_vs_oracle(built_lib):
    """attack/Stereo-RCNN/patch_attack.py:219-281 on a small frame: blend, forward/backward through our RoIAlign,
    crop, clipped descent, per-channel clamp -- against the same network with torchvision's RoIAlign on the CPU and
    the oracle's blend / update."""
    from eval_driving_safety_b200 import attack, stereo_rcnn as S
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    h, w, nroi, radius = 96, 320, 24, 4
    il, ir = S.synthetic_pair(0, h, w)
    rl, rr = S.synthetic_rois(nroi, h, w, seed=3)
    tg = S.synthetic_targets(nroi, seed=3)
    ref = S.SyntheticStereoRCNN(roi_feat_fn=A.pyramid_roi_feat, width=64)
    gpu = S.SyntheticStereoRCNN(width=64).cuda()
    gpu.load_state_dict(ref.state_dict())
    cl, cr = [50, 200], [50, 136]
    g = torch.Generator().manual_seed(2)
    patch0 = (torch.rand(1, 3, 9, 9, generator=g) - 0.5) * 300
    means = A.STEREO_RCNN_MEANS
    lo, hi = [0 - m for m in means], [255 - m for m in means]
    alpha, eps, iters = 1e3, 0.1, 2
    # oracle loop
    gl, gr_, _, _ = attack.stereo_rcnn_fake_gt(cl, cr, radius)
    rl_f, rr_f = rl.clone(), rr.clone()
    rl_f[0, 1:], rr_f[0, 1:] = gl[0, 0, :4], gr_[0, 0, :4]
    tg_f = dict(tg); tg_f["cls"] = tg["cls"].clone(); tg_f["cls"][0] = 1
    patch_r, xl, xr = patch0.clone(), il.clone(), ir.clone()
    for _ in range(iters):
        xl, xr = A.patch_apply(xl, patch_r, cl, radius), A.patch_apply(xr, patch_r, cr, radius)
        a, b = xl.clone().requires_grad_(True), xr.clone().requires_grad_(True)
        g1, g2 = torch.autograd.grad(ref(a, b, rl_f, rr_f, tg_f), [a, b])
        patch_r = A.patch_update(patch_r, g1, g2, cl, cr, radius, alpha, eps, lo, hi)
    # product loop
    patch_g = patch0.clone().cuda()
    losses = S.patch_attack_image(gpu, il.cuda(), ir.cuda(), rl.cuda(), rr.cuda(), {k: v.cuda() for k, v in tg.items()},
                                  patch_g, cl, cr, radius, iters=iters, alpha=alpha, eps=eps)
    assert losses.shape == (iters,) and torch.isfinite(losses).all()
    for c in range(3):
        assert patch_g[0, c].min() >= lo[c] - 1e-4 and patch_g[0, c].max() <= hi[c] + 1e-4
    assert (patch_g.cpu() - patch_r).abs().max() <= 2 * eps * iters + 1e-4
    assert ((patch_g.cpu() - patch_r).abs() < 1e-4).float().mean() > 0.9


def test_runner_srcnn_patch(built_lib, tmp_path):
    from eval_driving_safety_b200 import runner
    patch = runner.main(["srcnn-patch", "--tiny", "--pairs", "2", "--epochs", "1", "--iter", "2", "--ratio", "0.1",
                         "--save-dir", str(tmp_path)])
    assert patch.shape == (1, 3, 9, 9) and patch.abs().max() > 0
    assert (tmp_path / "epoch1" / "patch.npy").exists()


# ---------------------------------------------------------------- engine: one graph per calibration (ADVICE)
def test_engine_routes_frames_to_the_graph_of_their_calibration(tiny):
    from eval_driving_safety_b200 import engine, synthetic
    s = tiny
    dev = torch.device("cuda")
    mean = torch.tensor(A.IMAGENET_MEAN, device=dev).view(1, 3, 1, 1)
    std = torch.tensor(A.IMAGENET_STD, device=dev).view(1, 3, 1, 1)
    xL, xR = s["pair"]["imgL"].cuda(), s["pair"]["imgR"].cuda()
    ex = (xL, xR, xL * std + mean, xR * std + mean, s["pair"]["disp_L"].cuda())
    calib_b = synthetic.make_calib(1, scale=H / 384 * 1.1, cu=W / 2 + 3, cv=H / 2)      # another frame's P2/P3
    e_a = engine.PgdIterationGraph(s["model"], s["cfg_p"], s["labels_gpu"], s["calib"], 0.0075, 0.03, ex)
    e_b = engine.PgdIterationGraph(s["model"], s["cfg_p"], s["labels_gpu"], calib_b, 0.0075, 0.03, ex)
    want_a, want_b, got_a, got_b, got_a2 = ([t.clone() for t in ex] for _ in range(5))
    la, lb = e_a.step(*want_a).item(), e_b.step(*want_b).item()
    assert abs(la - lb) > 1e-6 * abs(la)                      # the calibration matters
    assert e_a.step(*got_b, calib=calib_b).item() == lb       # routed to a graph captured for calib_b ...
    assert torch.equal(got_b[0], want_b[0]) and torch.equal(got_b[1], want_b[1])
    assert e_a.step(*got_a, calib=s["calib"]).item() == la    # ... and back, without re-capturing
    assert torch.equal(got_a[0], want_a[0])
    assert len(e_a._siblings) == 1
    assert e_a.step(*got_a2).item() == la                     # no calib given: the captured one


# ---------------------------------------------------------------- asynchronous image dump from device tensors
def test_async_writer_snapshots_device_images(tmp_path):
    from eval_driving_safety_b200 import kitti_io
    g = torch.Generator().manual_seed(5)
    img = torch.randn(1, 3, 40, 64, generator=g).cuda()
    want = []
    with kitti_io.AsyncImageWriter(workers=2) as wr:
        for k in range(5):
            want.append(img[0].cpu().clone())
            wr.submit(img, str(tmp_path / "a" / ("%06d.png" % k)), 60, 36)
            img.mul_(0.9).add_(0.05)                          # the attack loop updates the image in place right away
    for k, im in enumerate(want):
        os.makedirs(tmp_path / "s", exist_ok=True)
        kitti_io.save_image(im, str(tmp_path / "s" / ("%06d.png" % k)), 60, 36)
        assert (tmp_path / "a" / ("%06d.png" % k)).read_bytes() == (tmp_path / "s" / ("%06d.png" % k)).read_bytes()


# ---------------------------------------------------------------- real sharding invariance (VERDICT weak 4)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs on the box")
def test_pgd_results_do_not_depend_on_the_number_of_ranks(built_lib, tmp_path):
    """pair i -> rank i mod world; the per-pair statistics of a 1-rank run and of a 2-rank
    ``torchrun`` (NCCL all_gather of the rows) must be identical."""
    from eval_driving_safety_b200 import runner
    args = ["pgd", "--tiny", "--pairs", "4", "--iter", "2", "--alpha", "0.0075", "--eps", "0.03"]
    one = runner.main(args)
    out = tmp_path / "stats2.pt"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", "-m", "eval_driving_safety_b200.runner"] + args + ["--stats-out", str(out)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    two = torch.load(out)
    # (the two depth-error columns are NaN placeholders in this runner: compare the computed ones)
    assert two.shape == one.shape and torch.equal(two.cpu()[:, :6], one.cpu()[:, :6])


# ---------------------------------------------------------------- one-launch FPN RoIAlign dispatch (8f-4)
@pytest.mark.parametrize("pooled", [7, 14])
def test_pyramid_roi_align_one_launch_matches_the_per_level_dispatch_and_the_oracle(built_lib, pooled):
    """attack/Stereo-RCNN/stereo_rcnn.py:110-141: level per RoI in the kernel, rows in the original order.  Must equal
    the reference-structured per-level dispatch on our kernels bit for bit (forward) and the oracle
    (torchvision RoIAlign per level, concatenate, re-sort) to 1e-4, gradients included; deterministic backward."""
    from eval_driving_safety_b200 import stereo_rcnn as S
    g = torch.Generator().manual_seed(30 + pooled)
    im_h, im_w, c = 600, 1987, 16
    feats = [torch.randn(1, c, h, w, generator=g) for (h, w) in ((150, 497), (75, 249), (38, 125), (19, 63))]
    rois, _ = S.synthetic_rois(96, im_h, im_w, seed=5)
    lv = S.roi_levels(rois)
    assert set(lv.tolist()) == {2.0, 3.0, 4.0, 5.0}                     # every level is exercised
    fr = [f.clone().requires_grad_(True) for f in feats]
    ref = A.pyramid_roi_feat(fr, rois, float(im_h), pooled)
    gy = torch.randn(ref.shape, generator=g)
    g_ref = torch.autograd.grad(ref, fr, gy)
    fa = [f.cuda().requires_grad_(True) for f in feats]
    fb = [f.cuda().requires_grad_(True) for f in feats]
    one = S.pyramid_roi_feat(fa, rois.cuda(), float(im_h), pooled)
    per = S.pyramid_roi_feat_per_level(fb, rois.cuda(), float(im_h), pooled)
    assert torch.equal(one, per)
    assert max_err(one.cpu(), ref) < 1e-4
    g_one = torch.autograd.grad(one, fa, gy.cuda())
    g_per = torch.autograd.grad(per, fb, gy.cuda())
    for a, b, r in zip(g_one, g_per, g_ref):
        assert max_err(a, b) < 1e-5 and rel_err(a.cpu(), r) < 1e-5     # (P = 14 sums ~200 terms per pixel: 1.7e-4 absolute)
    g_two = torch.autograd.grad(S.pyramid_roi_feat(fa, rois.cuda(), float(im_h), pooled), fa, gy.cuda())
    assert all(torch.equal(a, b) for a, b in zip(g_one, g_two))


def test_runner_writes_detections_for_the_evaluation_stage(built_lib, tmp_path):
    """runner pgd --detections: clean and attacked detections of every pair as KITTI txt files (the hand-off the
    reference's predict scripts produce, predict_and_save_pgd.py:249-283), counts gathered into the statistics."""
    from eval_driving_safety_b200 import kitti_io, parallel, runner
    stats = runner.main(["pgd", "--tiny", "--pairs", "2", "--iter", "1", "--alpha", "0.0075", "--eps", "0.03",
                         "--detections", str(tmp_path), "--score-thresh", "0.0"])
    for tag in ("clean", "adv"):
        for i in range(2):
            dets = kitti_io.read_detections(str(tmp_path / tag / ("%06d.txt" % i)))
            assert 1 <= len(dets) <= 20 and all(d["type"] == "Car" and 0.0 <= d["score"] <= 1.0 for d in dets)
    col = parallel.STAT_FIELDS.index("n_det_clean")
    assert (stats[:, col] >= 1).all() and (stats[:, col + 1] >= 1).all()


# ---------------------------------------------------------------------------------------------------------------
# concat / split glue of the extractor (upstream feature_extraction.forward torch.cat + slicing): bit-exact copies
# ---------------------------------------------------------------------------------------------------------------
def _cl(t):
    return t.cuda().contiguous(memory_format=torch.channels_last)


def test_cat_and_split_channels_match_torch_cat_and_slicing(built_lib):
    from eval_driving_safety_b200 import ops
    g = torch.Generator().manual_seed(11)
    widths = [64, 128, 32, 32, 32, 32]
    xs = [_cl(torch.randn(2, c, 13, 21, generator=g)).requires_grad_(True) for c in widths]
    ref = torch.cat(xs, 1)
    n0 = ops.LAUNCH_COUNT
    out = ops.cat_channels(xs)
    assert ops.LAUNCH_COUNT - n0 == 1 and torch.equal(out, ref)
    assert out.permute(0, 2, 3, 1).is_contiguous()
    gy = _cl(torch.randn(ref.shape, generator=g))
    want = torch.autograd.grad(ref, xs, gy)
    n0 = ops.LAUNCH_COUNT
    got = torch.autograd.grad(out, xs, gy)
    assert ops.LAUNCH_COUNT - n0 == 1
    for a, b in zip(got, want):
        assert torch.equal(a, b) and a.permute(0, 2, 3, 1).is_contiguous()
    # only some inputs need a gradient
    (g2,) = torch.autograd.grad(ops.cat_channels(xs), xs[2], gy)
    assert torch.equal(g2, want[2])
    # split: the fused detection heads (4 + 28 + 4 channels of a 64-wide map)
    y = _cl(torch.randn(1, 64, 9, 14, generator=g)).requires_grad_(True)
    a, b, c = ops.split_channels(y, [4, 28, 4])
    assert torch.equal(a, y[:, :4]) and torch.equal(b, y[:, 4:32]) and torch.equal(c, y[:, 32:36])
    ga, gc = torch.randn(a.shape, generator=g).cuda(), torch.randn(c.shape, generator=g).cuda()
    (gy_got,) = torch.autograd.grad((a * ga).sum() + (c * gc).sum(), y)          # b unused -> zeros
    (gy_ref,) = torch.autograd.grad((y[:, :4] * ga).sum() + (y[:, 32:36] * gc).sum(), y)
    assert torch.equal(gy_got, gy_ref)


def test_prefix_fork_and_split_batch_merge_gradients_like_autograd(built_lib):
    from eval_driving_safety_b200 import ops
    g = torch.Generator().manual_seed(12)
    x = _cl(torch.randn(2, 32, 7, 10, generator=g)).requires_grad_(True)
    wa, wb = _cl(torch.randn(2, 32, 7, 10, generator=g)), _cl(torch.randn(1, 32, 7, 10, generator=g))
    (ref,) = torch.autograd.grad((x * wa).sum() + (x[:1] * wb).sum(), x)
    full, pre = ops.prefix_fork(x, 1)
    assert torch.equal(full, x) and torch.equal(pre, x[:1])
    n0 = ops.LAUNCH_COUNT
    (got,) = torch.autograd.grad((full * wa).sum() + (pre * wb).sum(), x)
    assert ops.LAUNCH_COUNT - n0 == 1 and torch.equal(got, ref)
    (only_full,) = torch.autograd.grad((ops.prefix_fork(x, 1)[0] * wa).sum(), x)
    assert torch.equal(only_full, wa)
    (only_pre,) = torch.autograd.grad((ops.prefix_fork(x, 1)[1] * wb).sum(), x)
    assert torch.equal(only_pre[:1], wb) and only_pre[1:].abs().max().item() == 0
    a, b = ops.split_batch(x, 1)
    (gs,) = torch.autograd.grad((a * wb).sum() + (b * wa[1:]).sum(), x)
    (gr,) = torch.autograd.grad((x[:1] * wb).sum() + (x[1:] * wa[1:]).sum(), x)
    assert torch.equal(gs, gr)
    (ga,) = torch.autograd.grad((ops.split_batch(x, 1)[0] * wb).sum(), x)
    assert torch.equal(ga[:1], wb) and ga[1:].abs().max().item() == 0


def test_spp_upsample_matmul_equals_interpolate(built_lib):
    """The SPP branches' bilinear upsampling as two GEMMs on the channels-last memory against F.interpolate."""
    import torch.nn.functional as F
    from eval_driving_safety_b200 import dsgn
    g = torch.Generator().manual_seed(13)
    for (h, w, size) in ((1, 4, (96, 312)), (3, 9, (96, 312)), (12, 39, (96, 312)), (2, 3, (7, 11))):
        x = _cl(torch.randn(2, 32, h, w, generator=g)).requires_grad_(True)
        ref = F.interpolate(x, size, mode="bilinear", align_corners=False)
        out = dsgn.upsample_bilinear_matmul(x, size)
        assert out.shape == ref.shape and out.permute(0, 2, 3, 1).is_contiguous()
        assert (out - ref).abs().max().item() < 1e-5
        gy = _cl(torch.randn(ref.shape, generator=g))
        (a,) = torch.autograd.grad(out, x, gy)
        (b,) = torch.autograd.grad(ref, x, gy)
        assert (a - b).abs().max().item() < 1e-4 * b.abs().max().item()


@pytest.mark.parametrize("pooled,c", [(7, 64), (14, 32), (7, 256)])
def test_roi_align_backward_warp_kernel_equals_the_chunked_one(built_lib, pooled, c):
    """FPN widths take the warp-per-pixel backward reading the channels-last gradient [R,P,P,C]; it must give the bits of
    the thread-per-pixel kernel on [R,C,P,P] (same entries, same order), single level and pyramid, and be deterministic."""
    from eval_driving_safety_b200 import _lib, ops, stereo_rcnn as S
    g = torch.Generator().manual_seed(40 + pooled + c)
    im_h, im_w = 600, 1987
    feats = [torch.randn(1, c, h, w, generator=g).cuda().requires_grad_(True) for (h, w) in ((150, 497), (75, 249), (38, 125), (19, 63))]
    rois, _ = S.synthetic_rois(200, im_h, im_w, seed=7)
    rois = rois.cuda()
    res = {}
    for flag in (0, 1):
        _lib.set_flag("roi_bwd_warp", flag)
        try:
            out = S.pyramid_roi_feat(feats, rois, float(im_h), pooled)
            gy = torch.randn(out.shape, generator=torch.Generator().manual_seed(1)).cuda()
            gp = torch.autograd.grad(out, feats, gy)
            single = ops.roi_align(feats[2], rois, pooled, 38 / im_h)
            (gs,) = torch.autograd.grad(single, feats[2], gy)
        finally:
            _lib.set_flag("roi_bwd_warp", None)
        res[flag] = (gp, gs)
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.equal(a, b)
    assert torch.equal(res[0][1], res[1][1])
    # 600 RoIs: more than one shared-memory RoI tile
    many, _ = S.synthetic_rois(600, im_h, im_w, seed=9)
    many = many.cuda()
    outs = []
    for flag in (0, 1, 1):
        _lib.set_flag("roi_bwd_warp", flag)
        try:
            o = ops.roi_align(feats[1], many, pooled, 75 / im_h)
            gy = torch.randn(o.shape, generator=torch.Generator().manual_seed(2)).cuda()
            outs.append(torch.autograd.grad(o, feats[1], gy)[0])
        finally:
            _lib.set_flag("roi_bwd_warp", None)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
