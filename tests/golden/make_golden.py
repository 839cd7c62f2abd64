"""Generate the committed golden vectors by EXECUTING THE REFERENCE'S OWN LINES.

The reference ships no tests or fixtures for the attack path, and its scripts
cannot be imported (module-level argparse + un-vendored dsgn imports).  The
update arithmetic, however, is plain torch code: this script slices the cited
line ranges out of the files under /root/reference at run time, dedents them and
``exec``s them verbatim on seeded inputs.  Nothing is copied into the repo; only
the resulting input/output tensors are stored (tests/golden/*.npz).

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
"""
import os
import random
import textwrap

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("B2_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def ref_lines(rel, lo, hi):
    with open(os.path.join(REF, rel)) as f:
        lines = f.readlines()
    return textwrap.dedent("".join(lines[lo - 1:hi]))


def leaf_with_grad(x, g):
    x = x.clone().requires_grad_(True)
    x.grad = g.clone()
    return x


def dsgn_pgd():
    """attack/DSGN/pgd_attack.py: mean/std :153-154, (de)normalize :196-207, step :339-354."""
    src = "attack/DSGN/pgd_attack.py"
    ns = {"torch": torch}
    exec(ref_lines(src, 153, 154), ns)
    exec(ref_lines(src, 196, 207), ns)
    cases = {}
    for k, (alpha, eps, gscale) in enumerate([(1 / 255, 0.3, 1.0), (0.03 / 4, 0.03, 1.0), (8 / 255, 8 / 255, 0.0),
                                              (0.5, 0.0, 1.0), (1.0, 0.25, 1.0)]):
        g = torch.Generator().manual_seed(100 + k)
        clean01 = torch.rand(1, 3, 16, 24, generator=g)
        x01 = (clean01 + 0.2 * (torch.rand(1, 3, 16, 24, generator=g) - 0.5)).clamp(0, 1)
        gl = torch.randn(1, 3, 16, 24, generator=g) * gscale
        gl[0, :, :2] = 0.0                                    # sign(0) == 0 rows
        env = dict(ns)
        norm = env["normalize"]
        xL = norm(x01.clone())
        xR = norm(x01.flip(-1).clone())
        env.update(alpha=alpha, eps=eps,
                   imgL=leaf_with_grad(xL, gl), imgR=leaf_with_grad(xR, gl.flip(-2)),
                   clean_imgL_data=clean01.clone(), clean_imgR_data=clean01.flip(-1).clone(),
                   ori_imgL_data=clean01.clone(), ori_imgR_data=clean01.flip(-1).clone())
        exec(ref_lines(src, 339, 354), env)
        cases["pgd%d" % k] = dict(alpha=alpha, eps=eps, xL=xL, xR=xR, gL=gl, gR=gl.flip(-2), cleanL=clean01,
                                  cleanR=clean01.flip(-1), outL=env["imgL"].detach(), outR=env["imgR"].detach())
    return cases


def stereo_rcnn_pgd():
    """attack/Stereo-RCNN/pgd_attack.py:177-217 (0-255 mean-subtracted BGR)."""
    src = "attack/Stereo-RCNN/pgd_attack.py"
    cases = {}
    for k, (alpha, eps) in enumerate([(1.0, 255 * 0.03), (2.0, 255 * 0.3)]):
        g = torch.Generator().manual_seed(200 + k)
        means = torch.tensor([102.9801, 115.9465, 122.7717]).view(1, 3, 1, 1)
        clean = torch.rand(1, 3, 12, 20, generator=g) * 255 - means
        x = clean + (torch.rand(1, 3, 12, 20, generator=g) - 0.5) * 10
        gr = torch.randn(1, 3, 12, 20, generator=g)
        env = dict(torch=torch, alpha=alpha, eps=eps, im_left_data=leaf_with_grad(x, gr),
                   im_right_data=leaf_with_grad(x.flip(-1), gr.flip(-1)),
                   clean_im_left_data=clean.clone(), clean_im_right_data=clean.flip(-1).clone())
        exec(ref_lines(src, 177, 217), env)
        cases["srcnn%d" % k] = dict(alpha=alpha, eps=eps, x=x, g=gr, clean=clean, out=env["im_left_data"],
                                    outR=env["im_right_data"])
    return cases


def dsgn_patch():
    """attack/DSGN/patch_attack.py: generate_round_mask :237-256, pad :326-333, blend :373-376,
    crop + update :416-430 -- executed at the reference's hard-coded 384x1248 frame."""
    src = "attack/DSGN/patch_attack.py"
    ns = {"np": np, "random": random, "torch": torch, "nn": nn}
    exec(ref_lines(src, 237, 256), ns)
    random.seed(7)
    radius = 38
    center_l, center_r, mask_l, mask_r = ns["generate_round_mask"](radius)
    g = torch.Generator().manual_seed(300)
    patch = torch.randn(1, 3, 77, 77, generator=g)
    imgL = torch.randn(1, 3, 384, 1248, generator=g)
    imgR = torch.randn(1, 3, 384, 1248, generator=g)
    env = dict(ns)
    env.update(center_l=center_l, center_r=center_r, radius=radius, patch=patch.clone(),
               mask_l=torch.from_numpy(mask_l), mask_r=torch.from_numpy(mask_r),
               imgL=imgL.clone(), imgR=imgR.clone())
    exec(ref_lines(src, 326, 333), env)     # padding_l / padding_r
    exec(ref_lines(src, 369, 370), env)     # patch_l / patch_r
    exec(ref_lines(src, 373, 376), env)     # blend into imgL.data / imgR.data
    blendL, blendR = env["imgL"].clone(), env["imgR"].clone()
    gl = torch.randn(1, 3, 384, 1248, generator=g) * 1e-5
    gr = torch.randn(1, 3, 384, 1248, generator=g) * 1e-5
    env["imgL"] = leaf_with_grad(blendL, gl)
    env["imgR"] = leaf_with_grad(blendR, gr)
    env.update(alpha=1e3, eps=8 / 255)
    exec(ref_lines(src, 416, 430), env)     # grad clone, crop, patch -= clamp(...)
    box = lambda t, c: t[:, :, c[0] - radius - 2:c[0] + radius + 3, c[1] - radius - 2:c[1] + radius + 3].clone()
    # keep the fixture small: only the patch boxes (+2 px margin) of the big frames are stored
    return {"patch": dict(center_l=np.array(center_l), center_r=np.array(center_r), radius=radius, patch=patch,
                          imgL_box=box(imgL, center_l), imgR_box=box(imgR, center_r),
                          blendL_box=box(blendL, center_l), blendR_box=box(blendR, center_r),
                          gL_box=box(gl, center_l), gR_box=box(gr, center_r), patch_out=env["patch"],
                          mask_l_sum=float(mask_l.sum()), alpha=1e3, eps=8 / 255)}


def stereo_rcnn_patch_clamp():
    """attack/Stereo-RCNN/patch_attack.py:260-281: crop, update + per-channel range clamp."""
    src = "attack/Stereo-RCNN/patch_attack.py"
    g = torch.Generator().manual_seed(400)
    radius, H, W = 30, 600, 1987
    center_l, center_r = [300, 900], [300, 836]
    patch = (torch.rand(1, 3, 61, 61, generator=g) - 0.5) * 300
    gl = torch.randn(1, 3, H, W, generator=g) * 1e-2
    gr = torch.randn(1, 3, H, W, generator=g) * 1e-2
    env = dict(torch=torch, patch=patch.clone(), alpha=1e3, eps=0.3 * 255, radius=radius, center_l=center_l,
               center_r=center_r, im_left_data_grad=gl.clone(), im_right_data_grad=gr.clone())
    exec(ref_lines(src, 260, 281), env)
    box = lambda t, c: t[:, :, c[0] - radius:c[0] + radius + 1, c[1] - radius:c[1] + radius + 1].clone()
    return {"srcnn_patch": dict(patch=patch, gL_box=box(gl, center_l), gR_box=box(gr, center_r),
                                patch_out=env["patch"], alpha=1e3, eps=0.3 * 255, radius=radius)}


def roi_levels():
    """attack/Stereo-RCNN/stereo_rcnn.py:113-119 FPN level assignment (natural log)."""
    src = "attack/Stereo-RCNN/stereo_rcnn.py"
    g = torch.Generator().manual_seed(500)
    x1 = torch.rand(64, generator=g) * 1500
    y1 = torch.rand(64, generator=g) * 400
    w = torch.exp(torch.rand(64, generator=g) * 6.5)
    h = torch.exp(torch.rand(64, generator=g) * 5.5)
    rois = torch.stack([torch.zeros(64), x1, y1, x1 + w, y1 + h], 1)
    env = dict(torch=torch, rois=rois, im_info=torch.tensor([[600., 1987., 1.]]))
    exec(ref_lines(src, 113, 119), env)
    return {"roi_levels": dict(rois=rois, levels=env["roi_level"])}


def handoff():
    """Hand-off formats, by executing the reference's own lines:
    tensor2im attack/DSGN/pgd_attack.py:153-178, detection line attack/DSGN/predict_and_save_pgd.py:273-283."""
    import io
    import json
    ns = {"np": np, "torch": torch}
    exec(ref_lines("attack/DSGN/pgd_attack.py", 153, 154), ns)
    exec(ref_lines("attack/DSGN/pgd_attack.py", 157, 178), ns)
    g = torch.Generator().manual_seed(7)
    img = (torch.rand(3, 5, 7, generator=g) - torch.tensor(ns["mean"]).view(3, 1, 1)) / torch.tensor(ns["std"]).view(3, 1, 1)
    img[0, 0, 0] = (1.0 - 0.485) / 0.229          # exactly 1.0 after denormalisation -> 255
    img[1, 0, 1] = (0.999 - 0.456) / 0.224
    im = ns["tensor2im"](img.clone())
    body = ref_lines("attack/DSGN/predict_and_save_pgd.py", 273, 283)
    cases = []
    rnd = random.Random(3)
    for k in range(8):
        cls = [1, 2, 3, 2, 1, 2, 0, 2][k]
        bbox = [rnd.uniform(0, 1200) for _ in range(4)]
        h, w, l = (rnd.uniform(0.5, 4.5) for _ in range(3))
        c3 = [rnd.uniform(-30, 30), rnd.uniform(-2, 3), rnd.uniform(2, 60)]
        ry = rnd.uniform(-3.2, 3.2)
        score = rnd.random()
        if k == 6:                                  # the reference's "no 3-D box" defaults (:267-270)
            h, w, l, c3, ry = 0., 0., 0., [0., 0., 0.], 0.
        f = io.StringIO()
        env = dict(np=np, cls=cls, bbox=torch.tensor(bbox), h=h, w=w, l=l, box_center3d=torch.tensor(c3), ry=ry,
                   score=torch.tensor(score), f=f)
        exec(body, env)
        cases.append(dict(cls=cls, bbox=[float(v) for v in torch.tensor(bbox)], hwl=[h, w, l],
                          center3d=[float(v) for v in torch.tensor(c3)], ry=ry, score=float(torch.tensor(score)),
                          line=f.getvalue()))
    path = os.path.join(OUT, "handoff.json")
    with open(path, "w") as fh:
        json.dump(dict(tensor2im=dict(img=img.tolist(), out=im.tolist()), detections=cases), fh)
    print("wrote", path, os.path.getsize(path), "bytes")


class _Data:
    """Stand-in for a BoxList field: the reference only touches ``.data``."""

    def __init__(self, t):
        self.data = t


def dsgn_fake_gt():
    """attack/DSGN/patch_attack.py:336-354: zero the real ground truth, box 0 = the fake car."""
    import types
    g = torch.Generator().manual_seed(600)
    bbox, box3d = torch.rand(5, 4, generator=g) * 1000, torch.randn(5, 7, generator=g) * 10
    targets = [types.SimpleNamespace(bbox=_Data(bbox.clone()), box3d=_Data(box3d.clone()))]
    env = dict(torch=torch, targets=targets)
    exec(ref_lines("attack/DSGN/patch_attack.py", 336, 354), env)
    return {"fake_gt": dict(bbox_in=bbox, box3d_in=box3d, bbox_out=targets[0].bbox.data, box3d_out=targets[0].box3d.data)}


def dsgn_init_patch_resume():
    """attack/DSGN/patch_attack.py:211-234: resume from epoch0/patch.npy, cv2.INTER_LINEAR resize of a 61x61
    patch (trained on Stereo R-CNN) to the 77x77 of ratio 0.2; and the fresh zero patch."""
    import tempfile
    import cv2
    ns = {"np": np, "os": os, "cv2": cv2}
    exec(ref_lines("attack/DSGN/patch_attack.py", 211, 234), ns)
    g = torch.Generator().manual_seed(601)
    src = (torch.rand(1, 3, 61, 61, generator=g) - 0.5).numpy().astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "epoch0"))
        np.save(os.path.join(d, "epoch0", "patch.npy"), src)
        dim, radius, out = ns["init_patch"](0.2, d)
    with tempfile.TemporaryDirectory() as d:
        dim0, radius0, fresh = ns["init_patch"](0.2, os.path.join(d, "new"))
    return {"init_patch": dict(src=src, dim=dim, radius=radius, out=np.asarray(out, dtype=np.float32),
                               fresh_dim=dim0, fresh_radius=radius0, fresh_sum=float(np.abs(fresh).sum()),
                               fresh_shape=np.array(fresh.shape))}


def dsgn_loss_assembly():
    """attack/DSGN/pgd_attack.py:269 (mask) and :310-319 (depth term) executed on an eval-mode output dict
    (``depth_preds`` = ONE [1,H,W] tensor, which those lines iterate over the batch dimension)."""
    import types
    import warnings
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(602)
    cfg = types.SimpleNamespace(min_depth=2.0, max_depth=40.4, PlaneSweepVolume=True, loss_disp=True)
    disp_true = 2.0 + 38.4 * torch.rand(1, 12, 20, generator=g)
    disp_true = disp_true * (torch.rand(1, 12, 20, generator=g) >= 0.3)
    disp_true[0, 0, 0] = 40.4           # boundary: included (<=)
    disp_true[0, 0, 1] = 2.0            # boundary: excluded (>)
    pred = disp_true + torch.randn(1, 12, 20, generator=g) * 1.5      # residuals on both sides of smooth-L1's |x| = 1
    env = dict(torch=torch, F=F, cfg=cfg, disp_true=disp_true, outputs={"depth_preds": pred.clone()}, loss=0., losses=dict())
    exec(ref_lines("attack/DSGN/pgd_attack.py", 269, 270), env)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exec(ref_lines("attack/DSGN/pgd_attack.py", 310, 319), env)
    return {"loss_assembly": dict(disp_true=disp_true, pred=pred, mask=env["mask"].float(), loss=env["loss"].detach())}


def stereo_rcnn_fake_gt():
    """attack/Stereo-RCNN/patch_attack.py:188-207: the image's only ground-truth box = the patch square."""
    g = torch.Generator().manual_seed(603)
    center_l, center_r, radius = [311, 905], [311, 841], 30
    data = [None, None, None] + [torch.rand(1, 30, 5, generator=g) * 500 for _ in range(3)] + [None, None, _Data(torch.tensor(7))]
    env = dict(torch=torch, data=data, center_l=center_l, center_r=center_r, radius=radius)
    exec(ref_lines("attack/Stereo-RCNN/patch_attack.py", 188, 207), env)
    return {"srcnn_fake_gt": dict(center_l=np.array(center_l), center_r=np.array(center_r), radius=radius, gt_left=data[3],
                                  gt_right=data[4], gt_merge=data[5], num_boxes=data[8].data)}


def round2():
    allc = {}
    for fn in (dsgn_fake_gt, dsgn_init_patch_resume, dsgn_loss_assembly, stereo_rcnn_fake_gt):
        allc.update(fn())
    flat = {}
    for case, d in allc.items():
        for k, v in d.items():
            flat["%s/%s" % (case, k)] = v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    path = os.path.join(OUT, "attack_round2.npz")
    np.savez_compressed(path, **flat)
    print("wrote", path, os.path.getsize(path), "bytes,", len(flat), "arrays")


def main():
    round2()
    handoff()
    allc = {}
    for fn in (dsgn_pgd, stereo_rcnn_pgd, dsgn_patch, stereo_rcnn_patch_clamp, roi_levels):
        allc.update(fn())
    flat = {}
    for case, d in allc.items():
        for k, v in d.items():
            flat["%s/%s" % (case, k)] = v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    path = os.path.join(OUT, "attack_update.npz")
    np.savez_compressed(path, **flat)
    print("wrote", path, os.path.getsize(path), "bytes,", len(flat), "arrays")


if __name__ == "__main__":
    main()
