"""CPU oracle (TEST INFRASTRUCTURE, not product code) for the pixel-update
arithmetic of the attack loops.  Parity: PINNED by reference source text.

Every function is a functional restatement (no in-place aliasing, every sample
of the batch treated like the reference treats sample 0 -- SURVEY Appendix A)
of the lines cited in its docstring, with the same operation order so that the
fp32 result is bit-identical to executing those lines.
"""
import math
import random

import numpy as np
import torch

# attack/DSGN/pgd_attack.py:153-154
IMAGENET_MEAN = [0.485, 0.456, 0.406]
IMAGENET_STD = [0.229, 0.224, 0.225]
# attack/Stereo-RCNN/pgd_attack.py:189-207 (cfg.PIXEL_MEANS, BGR)
STEREO_RCNN_MEANS = [102.9801, 115.9465, 122.7717]


def denormalize(im, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """attack/DSGN/pgd_attack.py:196-200: ``im[c] = im[c] * std[c] + mean[c]``
    (applied to every sample, not just batch element 0)."""
    out = im.clone()
    for c in range(len(mean)):
        out[:, c] = im[:, c] * std[c] + mean[c]
    return out


def normalize(im, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """attack/DSGN/pgd_attack.py:203-207: ``im[c] = (im[c] - mean[c]) / std[c]``."""
    out = im.clone()
    for c in range(len(mean)):
        out[:, c] = (im[:, c] - mean[c]) / std[c]
    return out


def pgd_step_linf(x_norm, grad, clean01, alpha, eps, mean=IMAGENET_MEAN, std=IMAGENET_STD,
                  lo=0.0, hi=1.0):
    """One L-inf PGD update, attack/DSGN/pgd_attack.py:339-354.

    x01 = denorm(x); adv = x01 + alpha*sign(g); eta = clamp(adv-clean, -eps, eps);
    x01' = clamp(ori + eta, 0, 1) (ori == clean, :254 vs :297); return norm(x01').
    """
    x01 = denormalize(x_norm, mean, std)                       # :339-340
    adv = x01 + alpha * grad.sign()                            # :343-344
    eta = torch.clamp(adv - clean01, min=-eps, max=eps)        # :346-347
    x01n = torch.clamp(clean01 + eta, min=lo, max=hi)          # :349-350
    return normalize(x01n, mean, std)                          # :353-354


def pgd_step_l2(x_norm, grad, clean01, alpha, eps, mean=IMAGENET_MEAN, std=IMAGENET_STD,
                lo=0.0, hi=1.0):
    """L2 variant named by north_star; NOT in the reference (only the L-inf clamp
    exists, attack/DSGN/pgd_attack.py:346-347).  Standard Madry form per image:
    adv = x01 + alpha*g/||g||_2 ; eta = (adv-clean)*min(1, eps/||adv-clean||_2)."""
    x01 = denormalize(x_norm, mean, std)
    n = x_norm.shape[0]
    gnorm = grad.reshape(n, -1).double().pow(2).sum(1).sqrt().float().clamp_min(1e-12)
    adv = x01 + (alpha / gnorm).view(n, 1, 1, 1) * grad
    eta = adv - clean01
    enorm = eta.reshape(n, -1).double().pow(2).sum(1).sqrt().float().clamp_min(1e-12)
    factor = torch.clamp(eps / enorm, max=1.0).view(n, 1, 1, 1)
    eta = eta * factor
    x01n = torch.clamp(clean01 + eta, min=lo, max=hi)
    return normalize(x01n, mean, std)


def stereo_rcnn_pgd_step(x, grad, clean, alpha, eps255, means=STEREO_RCNN_MEANS):
    """attack/Stereo-RCNN/pgd_attack.py:177-217 (mean-subtracted 0-255 BGR space):
    adv = x + alpha*sign(g); eta = clamp(adv-clean, +-eps); per-channel clamp of
    clean+eta to [0-m_c, 255-m_c].  ``eps255`` is already 255*args.eps (:57)."""
    adv = x + alpha * grad.sign()                               # :177-179
    eta = torch.clamp(adv - clean, min=-eps255, max=eps255)     # :181-184
    holder = clean + eta                                        # :186-187
    chans = []
    for c in range(3):                                          # :189-207
        chans.append(torch.clamp(holder[:, c], min=(0 - means[c]), max=(255 - means[c])))
    return torch.stack(chans, 1)                                # :209-217


def patch_dim_radius(short_side, ratio):
    """attack/DSGN/patch_attack.py:213-218 / attack/Stereo-RCNN/patch_attack.py:58-65."""
    patch_dim = int(short_side * ratio)
    if patch_dim % 2 == 0:
        patch_dim += 1
    return patch_dim, int(patch_dim / 2)


def generate_round_mask(radius, rng=None, height=384, width=1248):
    """attack/DSGN/patch_attack.py:237-256 (height=384,width=1248) and
    attack/Stereo-RCNN/patch_attack.py:79-97 (600,1987).  ``rng`` replaces the
    reference's unseeded module-level ``random``."""
    rng = rng or random
    center_row = rng.randint(int(height * 0.4), int(height - radius - 1))
    center_col = rng.randint(int(width * 0.2), int(width * 0.8))
    center_l = [center_row, center_col]
    center_r = [center_row, int(center_col - (40 * 1.6))]
    Y, X = np.ogrid[:height, :width]
    masks = []
    for c in (center_l, center_r):
        dist = np.sqrt((Y - c[0]) ** 2 + (X - c[1]) ** 2)
        m = (dist <= radius).astype('float32')
        masks.append(torch.from_numpy(np.array([[m, m, m]])))
    return center_l, center_r, masks[0], masks[1]


def round_mask(center, radius, height, width):
    Y, X = np.ogrid[:height, :width]
    dist = np.sqrt((Y - center[0]) ** 2 + (X - center[1]) ** 2)
    m = (dist <= radius).astype('float32')
    return torch.from_numpy(np.array([[m, m, m]]))


def patch_apply(img, patch, center, radius):
    """attack/DSGN/patch_attack.py:326-333, 369-376: zero-pad the patch to the
    frame at (row-r, col-r) and blend ``(1-m)*img + m*pad``.  img: [1,3,H,W]."""
    h, w = img.shape[-2:]
    mask = round_mask(center, radius, h, w).to(img.dtype)
    pad = torch.nn.ConstantPad2d((center[1] - radius, (w - 1) - (center[1] + radius),
                                  center[0] - radius, (h - 1) - (center[0] + radius)), 0.0)
    return torch.mul((1 - mask), img) + torch.mul(mask, pad(patch))


def patch_update(patch, grad_l, grad_r, center_l, center_r, radius, alpha, eps, lo=None, hi=None):
    """attack/DSGN/patch_attack.py:416-430: crop both image gradients at the patch
    boxes, ``patch -= clamp(0.5*alpha*(gL+gR), -eps, eps)``.  With ``lo/hi`` (per
    channel) also the Stereo R-CNN range clamp, attack/Stereo-RCNN/patch_attack.py:272-281."""
    gl = grad_l[:, :, (center_l[0] - radius):(center_l[0] + radius + 1),
                (center_l[1] - radius):(center_l[1] + radius + 1)]
    gr = grad_r[:, :, (center_r[0] - radius):(center_r[0] + radius + 1),
                (center_r[1] - radius):(center_r[1] + radius + 1)]
    out = patch - torch.clamp(0.5 * alpha * (gl + gr), min=-eps, max=eps)
    if lo is not None:
        chans = [torch.clamp(out[:, c], min=lo[c], max=hi[c]) for c in range(out.shape[1])]
        out = torch.stack(chans, 1)
    return out


def roi_levels(rois):
    """attack/Stereo-RCNN/stereo_rcnn.py:113-119: FPN level per RoI.  Natural log
    (not log2) is the reference's behaviour and is reproduced."""
    h = rois[:, 4] - rois[:, 2] + 1
    w = rois[:, 3] - rois[:, 1] + 1
    lvl = torch.round(torch.log(torch.sqrt(h * w) / 224.0) + 4)
    lvl[lvl < 2] = 2
    lvl[lvl > 5] = 5
    return lvl


def pyramid_roi_feat(feat_maps, rois, im_h, pooled):
    """attack/Stereo-RCNN/stereo_rcnn.py:110-141 with the upstream ROIAlign
    (maskrcnn-benchmark, sampling_ratio=0, legacy/unaligned) replaced by the
    stock ``torchvision.ops.roi_align(aligned=False)`` which implements the same
    published algorithm."""
    from torchvision.ops import roi_align
    lvl = roi_levels(rois)
    feats, idxs = [], []
    for i, l in enumerate(range(2, 6)):
        sel = (lvl == l).nonzero().reshape(-1)
        if sel.numel() == 0:
            continue
        idxs.append(sel)
        scale = feat_maps[i].size(2) / im_h
        feats.append(roi_align(feat_maps[i], rois[sel], (pooled, pooled), scale, 0, False))
    feat = torch.cat(feats, 0)
    order = torch.sort(torch.cat(idxs, 0))[1]
    return feat[order]
