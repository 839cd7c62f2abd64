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
    """``_StereoRCNN.PyramidRoI_Feat``: per level l in 2..5, RoIAlign(feat_l,
    rois[level==l], scale = feat_l.H / im_h); concatenate; restore RoI order.
    ``pooled`` = 7 (cfg.POOLING_SIZE) or 14 for the keypoint branch (:44-45)."""
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
