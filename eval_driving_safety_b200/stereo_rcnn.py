"""Stereo R-CNN side of the hot path (BASELINE config 5): the FPN-level RoIAlign
dispatch of attack/Stereo-RCNN/stereo_rcnn.py:110-141 on the sm_100a RoIAlign
kernels, and the 0-255-space PGD step of attack/Stereo-RCNN/pgd_attack.py:177-217
(``attack.stereo_rcnn_pgd_step``)."""
import torch

from . import ops


def roi_levels(rois):
    """stereo_rcnn.py:113-119 -- natural log (not log2) reproduced on purpose."""
    h = rois[:, 4] - rois[:, 2] + 1
    w = rois[:, 3] - rois[:, 1] + 1
    lvl = torch.round(torch.log(torch.sqrt(h * w) / 224.0) + 4)
    return lvl.clamp(2, 5)


def pyramid_roi_feat(feat_maps, rois, im_h, pooled):
    """``_StereoRCNN.PyramidRoI_Feat`` (stereo_rcnn.py:110-141) in ONE launch: the level of every RoI is evaluated
    inside the kernel and its pooled row lands in the original RoI order, so the per-level ``nonzero`` index lists
    (a host synchronisation each), the concatenation and the re-sort of the reference disappear.
    ``pooled`` = 7 (cfg.POOLING_SIZE) or 14 for the keypoint branch (:44-45)."""
    return ops.pyramid_roi_align(feat_maps, rois, im_h, pooled)


def pyramid_roi_feat_per_level(feat_maps, rois, im_h, pooled):
    """The reference's own structure (one RoIAlign per level, concatenate, restore the order); kept as the
    cross-check of the one-launch dispatch."""
    lvl = roi_levels(rois)
    feats, idxs = [], []
    for i, l in enumerate(range(2, 6)):
        sel = (lvl == l).nonzero().reshape(-1)
        if sel.numel() == 0:
            continue
        idxs.append(sel)
        scale = feat_maps[i].size(2) / im_h
        feats.append(ops.roi_align(feat_maps[i], rois[sel], pooled, scale))
    feat = torch.cat(feats, 0)
    order = torch.sort(torch.cat(idxs, 0))[1]
    return feat[order]


# ---------------------------------------------------------------------------------------------
# Stereo-R-CNN-shaped stand-in for BASELINE config 5.  The real detector (ResNet-101 + FPN +
# stereo RPN + proposal/target layers) lives in the un-vendored HKUST-Aerial-Robotics/Stereo-RCNN
# tree (attack/Stereo-RCNN/README.md:12) and cannot run here; what the reference owns on this
# path is the RoIAlign dispatch (stereo_rcnn.py:110-141, 250-262), the uncertainty-weighted loss
# (pgd_attack.py:165-171) and the 0-255-space PGD step (:177-217).  This network reproduces that
# data flow with a small stock-torch FPN so the sm_100a RoIAlign kernels run forward AND backward
# inside a PGD loop at the reference's 600x1987 frame size.
# ---------------------------------------------------------------------------------------------
import torch.nn as nn
import torch.nn.functional as F


class SyntheticStereoRCNN(nn.Module):
    def __init__(self, roi_feat_fn=None, width=256, seed=1):
        super().__init__()
        torch.manual_seed(seed)
        self.roi_feat_fn = roi_feat_fn or pyramid_roi_feat
        c = width
        self.stem = nn.Sequential(nn.Conv2d(3, 32, 7, 2, 3), nn.ReLU(inplace=True), nn.MaxPool2d(3, 2, 1))
        self.c2 = nn.Sequential(nn.Conv2d(32, 64, 3, 1, 1), nn.ReLU(inplace=True))
        self.c3 = nn.Sequential(nn.Conv2d(64, 128, 3, 2, 1), nn.ReLU(inplace=True))
        self.c4 = nn.Sequential(nn.Conv2d(128, 256, 3, 2, 1), nn.ReLU(inplace=True))
        self.c5 = nn.Sequential(nn.Conv2d(256, 256, 3, 2, 1), nn.ReLU(inplace=True))
        self.lat = nn.ModuleList([nn.Conv2d(k, c, 1) for k in (64, 128, 256, 256)])
        self.smooth = nn.ModuleList([nn.Conv2d(c, c, 3, 1, 1) for _ in range(4)])
        self.head = nn.Sequential(nn.Linear(2 * c, 256), nn.ReLU(inplace=True))
        self.cls_score = nn.Linear(256, 2)
        self.bbox_pred = nn.Linear(256, 6)
        self.dim_orien = nn.Linear(256, 5)
        self.kpts = nn.Conv2d(c, 6, 3, 1, 1)
        self.register_buffer("uncert", torch.rand(6))              # pgd_attack.py:47 (loaded from ckpt upstream)
        for p in self.parameters():
            p.requires_grad_(False)
        self.eval()

    @staticmethod
    def _upsample_add(x, y):
        """stereo_rcnn.py:105-108"""
        return F.interpolate(x, size=y.shape[-2:], mode='bilinear', align_corners=False) + y

    def fpn(self, im):
        c2 = self.c2(self.stem(im)); c3 = self.c3(c2); c4 = self.c4(c3); c5 = self.c5(c4)
        p5 = self.lat[3](c5)
        p4 = self._upsample_add(p5, self.lat[2](c4))
        p3 = self._upsample_add(p4, self.lat[1](c3))
        p2 = self._upsample_add(p3, self.lat[0](c2))
        return [s(p) for s, p in zip(self.smooth, (p2, p3, p4, p5))]

    def forward(self, im_left, im_right, rois_left, rois_right, targets):
        im_h = float(im_left.shape[-2])
        fl, fr = self.fpn(im_left), self.fpn(im_right)
        # stereo_rcnn.py:250-251: 7x7 left and right features concatenated on channels
        sem = torch.cat((self.roi_feat_fn(fl, rois_left, im_h, 7), self.roi_feat_fn(fr, rois_right, im_h, 7)), 1)
        h = self.head(sem.mean((2, 3)))
        dense = self.roi_feat_fn(fl, rois_left, im_h, 14)          # :260 keypoint branch, 14x14
        kp = self.kpts(dense).sum(2)                               # :263 sum over rows
        losses = [F.cross_entropy(self.cls_score(h), targets['cls']),
                  F.smooth_l1_loss(self.bbox_pred(h), targets['bbox']),
                  F.smooth_l1_loss(self.dim_orien(h), targets['dim']),
                  F.cross_entropy(kp.flatten(0, 1), targets['kpts'].flatten()),
                  self.cls_score(h).logsumexp(1).mean() * 0.1,
                  self.bbox_pred(h).abs().mean() * 0.1]
        u = self.uncert
        # pgd_attack.py:165-171: sum_i L_i * exp(-u_i) + u_i
        return sum(l * torch.exp(-u[i]) + u[i] for i, l in enumerate(losses))


def synthetic_rois(n_rois, im_h, im_w, seed):
    """Seeded RoIs [R,5] spread over the FPN levels 2..5, right view shifted."""
    g = torch.Generator().manual_seed(seed)
    side = torch.exp(torch.rand(n_rois, generator=g) * 4.2 + 2.5)            # 12 .. 800 px
    w = (side * (0.6 + 0.8 * torch.rand(n_rois, generator=g))).clamp(8, im_w * 0.8)
    h = (side * (0.4 + 0.4 * torch.rand(n_rois, generator=g))).clamp(8, im_h * 0.8)
    x1 = torch.rand(n_rois, generator=g) * (im_w - w - 1)
    y1 = torch.rand(n_rois, generator=g) * (im_h - h - 1)
    left = torch.stack([torch.zeros(n_rois), x1, y1, x1 + w, y1 + h], 1)
    disp = torch.rand(n_rois, generator=g) * 60
    right = left.clone()
    right[:, 1] = (left[:, 1] - disp).clamp_min(0)
    right[:, 3] = (left[:, 3] - disp).clamp_min(8)
    return left, right


def synthetic_targets(n_rois, seed):
    g = torch.Generator().manual_seed(seed + 1)
    return {'cls': torch.randint(0, 2, (n_rois,), generator=g), 'bbox': torch.randn(n_rois, 6, generator=g),
            'dim': torch.randn(n_rois, 5, generator=g), 'kpts': torch.randint(0, 14, (n_rois, 6), generator=g)}


def synthetic_pair(i, im_h=600, im_w=1987, means=(102.9801, 115.9465, 122.7717)):
    """Mean-subtracted 0-255 BGR frames, the space of attack/Stereo-RCNN/pgd_attack.py."""
    g = torch.Generator().manual_seed(2000 + i)
    img = torch.rand(1, 3, im_h, im_w, generator=g)
    img = F.avg_pool2d(F.pad(img, (2, 2, 2, 2), mode='replicate'), 5, 1) * 255
    m = torch.tensor(means).view(1, 3, 1, 1)
    return img - m, torch.roll(img, -24, 3) - m


def pgd_attack(model, im_left, im_right, rois_left, rois_right, targets, iters, alpha, eps255):
    """The loop of attack/Stereo-RCNN/pgd_attack.py:151-217 (explicit clone of the clean images,
    SURVEY App. A).  Returns (adv_left, adv_right, losses)."""
    from . import attack
    clean_l, clean_r = im_left.clone(), im_right.clone()
    xl, xr, losses = im_left.clone(), im_right.clone(), []
    for _ in range(iters):
        xl.requires_grad_(True); xr.requires_grad_(True)
        loss = model(xl, xr, rois_left, rois_right, targets)
        gl, gr = torch.autograd.grad(loss, [xl, xr])
        xl = attack.stereo_rcnn_pgd_step(xl.detach(), gl.contiguous(), clean_l, alpha, eps255)
        xr = attack.stereo_rcnn_pgd_step(xr.detach(), gr.contiguous(), clean_r, alpha, eps255)
        losses.append(loss.detach())
    return xl, xr, torch.stack(losses)


def patch_attack_image(model, im_left, im_right, rois_left, rois_right, targets, patch, center_l, center_r, radius,
                       iters=2, alpha=1e3, eps=0.1, means=(102.9801, 115.9465, 122.7717), delta_hook=None):
    """Inner loop of the Stereo R-CNN universal-patch attack for one image,
    attack/Stereo-RCNN/patch_attack.py:219-281 (defaults :43-46, alpha :102): blend the patch into both 600x1987
    frames (:225-230), forward + backward, crop both gradients at the patch boxes (:260-266), clipped DESCENT
    step (:268-270), per-channel clamp of the patch to the valid mean-subtracted 0-255 range (:272-281).
    The image's only ground-truth box is the patch's own square (``attack.stereo_rcnn_fake_gt``, :187-207):
    RoI 0 of both views is that square and its class target is 1.  ``patch`` [1,3,dim,dim] is updated in
    place; ``delta_hook`` all-reduces the clipped step over the ranks.  Returns the per-iteration losses."""
    from . import attack
    lo = [0 - m for m in means]
    hi = [255 - m for m in means]
    gl, gr, _, _ = attack.stereo_rcnn_fake_gt(center_l, center_r, radius)
    rois_left, rois_right = rois_left.clone(), rois_right.clone()
    rois_left[0, 1:] = gl[0, 0, :4].to(rois_left)
    rois_right[0, 1:] = gr[0, 0, :4].to(rois_right)
    targets = dict(targets)
    targets['cls'] = targets['cls'].clone()
    targets['cls'][0] = 1
    losses = []
    for _ in range(iters):
        attack.patch_apply(im_left, patch, center_l, radius)
        attack.patch_apply(im_right, patch, center_r, radius)
        xl, xr = im_left.detach().requires_grad_(True), im_right.detach().requires_grad_(True)
        loss = model(xl, xr, rois_left, rois_right, targets)
        g_l, g_r = torch.autograd.grad(loss, [xl, xr])
        if delta_hook is None:
            attack.patch_update(patch, g_l.contiguous(), g_r.contiguous(), center_l, center_r, radius, alpha, eps, lo, hi)
        else:
            delta = torch.empty_like(patch)
            attack.patch_update(patch, g_l.contiguous(), g_r.contiguous(), center_l, center_r, radius, alpha, eps,
                                delta_out=delta)
            attack.patch_axpy(patch, delta_hook(delta), lo, hi)
        losses.append(loss.detach())
    return torch.stack(losses)
