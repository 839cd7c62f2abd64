"""Attack-step functions with the semantics of the reference's inline loops.

The reference exposes no function API for the step (it is script code in
``main()``); these functions carry exactly those lines' semantics so that
attack/DSGN/*.py and attack/Stereo-RCNN/*.py can call them in place of the ~26
ATen launches per iteration (INTEGRATION.md shows the two-line patch):

  pgd_step / pgd_step_pair  == attack/DSGN/pgd_attack.py:339-354
  stereo_rcnn_pgd_step      == attack/Stereo-RCNN/pgd_attack.py:177-217
  patch_apply               == attack/DSGN/patch_attack.py:326-333, 369-376
  patch_update              == attack/DSGN/patch_attack.py:416-430
                               (+ attack/Stereo-RCNN/patch_attack.py:272-281 with lo/hi)
  generate_round_mask       == attack/DSGN/patch_attack.py:237-243 (centres only;
                               the mask itself is evaluated inside the kernel)
  pgd_attack / patch_attack == the ``for iteration in range(...)`` loops
                               (pgd_attack.py:300-354, patch_attack.py:367-430)
"""
import ctypes
import random

import torch

from . import _lib
from ._lib import check, c_vp, f32_array
from . import ops
from .ops import _need_cuda, _p, _stream

# attack/DSGN/pgd_attack.py:153-154
IMAGENET_MEAN = [0.485, 0.456, 0.406]
IMAGENET_STD = [0.229, 0.224, 0.225]
# attack/Stereo-RCNN/pgd_attack.py:189-207
STEREO_RCNN_MEANS = [102.9801, 115.9465, 122.7717]


def _chan(v, c, default):
    if v is None:
        v = default
    if isinstance(v, (int, float)):
        v = [v] * c
    assert len(v) == c
    return [float(t) for t in v]


def _pgd_update_sets(xs, gs, cs, outs, alpha, eps, mean, std, lo, hi):
    lib = _lib.load()
    n, c, h, w = xs[0].shape
    for t in list(xs) + list(gs) + list(cs) + list(outs):
        _need_cuda(t)
        if tuple(t.shape) != (n, c, h, w) or not t.is_contiguous():
            raise RuntimeError("pgd_step: tensors must be contiguous NCHW of identical shape")
    denorm = mean is not None
    lo_a, hi_a = f32_array(_chan(lo, c, 0.0)), f32_array(_chan(hi, c, 1.0))
    mean_a = f32_array(_chan(mean, c, 0.0)) if denorm else None
    std_a = f32_array(_chan(std, c, 1.0)) if denorm else None
    keep = [_lib.ptr_array(ts) for ts in (xs, gs, cs, outs)]
    ptrs = [ctypes.cast(k, ctypes.POINTER(ctypes.c_void_p)) for k in keep]
    with ops._op("pgd_update", 1, 16 * len(xs) * n * c * h * w):      # 3 reads + 1 write per element
        check(lib.b2_pgd_update(ptrs[0], ptrs[1], ptrs[2], ptrs[3], len(xs), n, c, h * w, float(alpha), float(eps),
                                int(denorm), mean_a, std_a, lo_a, hi_a, _stream()), "pgd_update")
    return outs


def pgd_step(x_norm, grad, clean01, alpha, eps, mean=IMAGENET_MEAN, std=IMAGENET_STD, norm='linf',
             lo=0.0, hi=1.0, out=None):
    """One PGD update of a batch of normalised images [N,C,H,W] (every sample is
    treated like the reference treats sample 0).  ``mean=None`` skips the
    (de)normalisation (Stereo R-CNN space).  Returns a new tensor (or ``out``)."""
    out = torch.empty_like(x_norm) if out is None else out
    if norm == 'linf':
        _pgd_update_sets([x_norm], [grad], [clean01], [out], alpha, eps, mean, std, lo, hi)
        return out
    if norm != 'l2':
        raise ValueError("norm must be 'linf' or 'l2'")
    lib = _lib.load()
    _need_cuda(x_norm, grad, clean01, out)
    n, c, h, w = x_norm.shape
    denorm = mean is not None
    ws = torch.empty(lib.b2_pgd_update_l2_workspace_bytes(n), device=x_norm.device, dtype=torch.uint8)
    check(lib.b2_pgd_update_l2(_p(x_norm.contiguous()), _p(grad.contiguous()), _p(clean01.contiguous()), _p(out),
                               n, c, h * w, float(alpha), float(eps), int(denorm),
                               f32_array(_chan(mean, c, 0.0)) if denorm else None,
                               f32_array(_chan(std, c, 1.0)) if denorm else None,
                               f32_array(_chan(lo, c, 0.0)), f32_array(_chan(hi, c, 1.0)), _p(ws), _stream()),
          "pgd_update_l2")
    return out


def pgd_step_pair(xL, gL, cL, xR, gR, cR, alpha, eps, mean=IMAGENET_MEAN, std=IMAGENET_STD, lo=0.0, hi=1.0,
                  inplace=False):
    """Left and right image batches updated by ONE kernel launch."""
    oL = xL if inplace else torch.empty_like(xL)
    oR = xR if inplace else torch.empty_like(xR)
    _pgd_update_sets([xL, xR], [gL, gR], [cL, cR], [oL, oR], alpha, eps, mean, std, lo, hi)
    return oL, oR


def stereo_rcnn_pgd_step(x, grad, clean, alpha, eps255, means=STEREO_RCNN_MEANS):
    """attack/Stereo-RCNN/pgd_attack.py:177-217; ``eps255`` = 255*args.eps (:57)."""
    lo = [0 - m for m in means]
    hi = [255 - m for m in means]
    return pgd_step(x, grad, clean, alpha, eps255, mean=None, std=None, lo=lo, hi=hi)


def patch_dim_radius(short_side, ratio):
    """attack/DSGN/patch_attack.py:213-218."""
    patch_dim = int(short_side * ratio)
    if patch_dim % 2 == 0:
        patch_dim += 1
    return patch_dim, int(patch_dim / 2)


def generate_round_mask(radius, rng=None, height=384, width=1248):
    """Patch centres of attack/DSGN/patch_attack.py:239-243 (seedable ``rng``
    instead of the module-level unseeded ``random``).  Returns (center_l, center_r)."""
    rng = rng or random
    center_row = rng.randint(int(height * 0.4), int(height - radius - 1))
    center_col = rng.randint(int(width * 0.2), int(width * 0.8))
    return [center_row, center_col], [center_row, int(center_col - (40 * 1.6))]


def patch_apply(img, patch, center, radius):
    """In-place blend of the circular patch into ``img`` [N,C,H,W] at ``center``
    (one (row, col), a list of N of them, or an int32 CUDA tensor [N,2] -- read at run time, so a captured
    CUDA graph can be replayed with other positions)."""
    lib = _lib.load()
    _need_cuda(img, patch)
    n, c, h, w = img.shape
    if not img.is_contiguous():
        raise RuntimeError("patch_apply: img must be contiguous NCHW")
    if isinstance(center, torch.Tensor):
        if not (center.is_cuda and center.dtype == torch.int32 and center.numel() == 2 * n and center.is_contiguous()):
            raise RuntimeError("patch_apply: device centres must be a contiguous int32 CUDA tensor [N,2]")
        patch = patch.reshape(c, 2 * radius + 1, 2 * radius + 1).contiguous()
        check(lib.b2_patch_apply_dev(_p(img), _p(patch), n, c, h, w, c_vp(center.data_ptr()), int(radius), _stream()),
              "patch_apply_dev")
        return img
    centers = [center] * n if isinstance(center[0], int) else list(center)
    flat = (ctypes.c_int * (2 * n))(*[int(v) for cc in centers for v in cc])
    patch = patch.reshape(c, 2 * radius + 1, 2 * radius + 1).contiguous()
    check(lib.b2_patch_apply(_p(img), _p(patch), n, c, h, w, flat, int(radius), _stream()), "patch_apply")
    return img


def patch_update(patch, grad_l, grad_r, center_l, center_r, radius, alpha, eps, lo=None, hi=None,
                 delta_out=None):
    """In-place patch step from the two image gradients [1,C,H,W].  ``center_l`` may be an int32 CUDA tensor
    (cyL, cxL, cyR, cxR) read at run time (``center_r`` is then ignored): graph-replayable."""
    lib = _lib.load()
    _need_cuda(patch, grad_l, grad_r)
    _, c, h, w = grad_l.shape
    if not (patch.is_contiguous() and grad_l.is_contiguous() and grad_r.is_contiguous()):
        raise RuntimeError("patch_update: tensors must be contiguous")
    if isinstance(center_l, torch.Tensor):
        if not (center_l.is_cuda and center_l.dtype == torch.int32 and center_l.numel() == 4 and center_l.is_contiguous()):
            raise RuntimeError("patch_update: device centres must be a contiguous int32 CUDA tensor (cyL,cxL,cyR,cxR)")
        check(lib.b2_patch_update_dev(_p(patch), _p(grad_l), _p(grad_r), c, h, w, c_vp(center_l.data_ptr()), int(radius),
                                      float(alpha), float(eps), f32_array(lo), f32_array(hi), _p(delta_out), _stream()),
              "patch_update_dev")
        return patch
    check(lib.b2_patch_update(_p(patch), _p(grad_l), _p(grad_r), c, h, w, int(center_l[0]), int(center_l[1]),
                              int(center_r[0]), int(center_r[1]), int(radius), float(alpha), float(eps),
                              f32_array(lo), f32_array(hi), _p(delta_out), _stream()), "patch_update")
    return patch


def patch_axpy(patch, delta, lo=None, hi=None):
    """patch = clamp(patch - delta, lo_c, hi_c) in place: second half of the split update of the multi-GPU
    universal patch (``patch_update(..., delta_out=)`` -> all_reduce -> this)."""
    lib = _lib.load()
    _need_cuda(patch, delta)
    c, dim = patch.shape[-3], patch.shape[-1]
    if not (patch.is_contiguous() and delta.is_contiguous()) or delta.numel() != patch.numel():
        raise RuntimeError("patch_axpy: patch and delta must be contiguous tensors of the same size")
    check(lib.b2_patch_axpy(_p(patch), _p(delta), c, dim, f32_array(lo), f32_array(hi), _stream()), "patch_axpy")
    return patch


# ---------------------------------------------------------------------------
# Targets of the universal patch attack and patch initialisation (host side)
# ---------------------------------------------------------------------------
# attack/DSGN/patch_attack.py:342-354 -- the fake ground truth every image is given: a car at 29 m
FAKE_GT_BBOX = (569.33, 180.88, 613.91, 225.02)                       # (x1, y1, x2, y2), :343-346
FAKE_GT_BOX3D = (1.65, 1.67, 3.64, -0.78, 1.98, 29.11, -1.60)         # (h, w, l, x, y, z, theta), :349-355


def inject_fake_gt(bbox, box3d):
    """attack/DSGN/patch_attack.py:336-354: zero every ground-truth box of the image and make box 0 the fake
    car.  ``bbox`` [K,4], ``box3d`` [K,7] (the ``.data`` of ``targets[0].bbox`` / ``.box3d``), modified in place
    and returned."""
    for i in range(len(bbox)):
        bbox[i] = torch.zeros(size=bbox[i].shape)
        box3d[i] = torch.zeros(size=box3d[i].shape)
    for j, v in enumerate(FAKE_GT_BBOX):
        bbox[0, j] = v
    for j, v in enumerate(FAKE_GT_BOX3D):
        box3d[0, j] = v
    return bbox, box3d


def stereo_rcnn_fake_gt(center_l, center_r, radius, max_boxes=30):
    """attack/Stereo-RCNN/patch_attack.py:187-207: the only ground-truth box of the image is the patch's own
    bounding square (left, right and merged = left), num_boxes = 1.  Returns (gt_left, gt_right, gt_merge)
    [1,max_boxes,5] and num_boxes."""
    gl, gr, gm = (torch.zeros(1, max_boxes, 5) for _ in range(3))
    for t, c in ((gl, center_l), (gr, center_r), (gm, center_l)):
        t[0, 0, 0] = c[1] - radius
        t[0, 0, 1] = c[0] - radius
        t[0, 0, 2] = c[1] + radius
        t[0, 0, 3] = c[0] + radius
    return gl, gr, gm, torch.tensor(1)


def resize_patch(patch, dim):
    """attack/DSGN/patch_attack.py:222-227: a patch trained at another size (e.g. on Stereo R-CNN) is resized
    with ``cv2.resize(..., INTER_LINEAR)``; bilinear with half-pixel centres and edge clamping is exactly
    ``F.interpolate(mode='bilinear', align_corners=False)`` (no OpenCV dependency in the product; pinned against
    cv2 in tests/test_host_logic.py)."""
    patch = torch.as_tensor(patch, dtype=torch.float32)
    if patch.shape[-1] == dim and patch.shape[-2] == dim:
        return patch.clone()
    return torch.nn.functional.interpolate(patch, size=(dim, dim), mode='bilinear', align_corners=False)


def init_patch(patch_ratio, save_dir, short_side=384, resize=True):
    """attack/DSGN/patch_attack.py:211-234 (``short_side`` 384, resume + resize) and
    attack/Stereo-RCNN/patch_attack.py:58-76 (600, ``resize=False``: loaded as is).  Resumes from
    ``<save_dir>/epoch0/patch.npy`` when that directory exists, else creates it with a zero patch.
    Returns (patch_dim, radius, patch [1,3,dim,dim] float32 numpy array)."""
    import os
    import numpy as np
    d0 = os.path.join(save_dir, 'epoch0')
    patch_dim, radius = patch_dim_radius(short_side, patch_ratio)
    if os.path.isdir(d0):
        patch = np.load(os.path.join(d0, 'patch.npy'))
        if resize:
            patch = resize_patch(patch, patch_dim).numpy()
    else:
        os.makedirs(d0)
        patch = np.zeros((1, 3, patch_dim, patch_dim), dtype=np.float32)
        np.save(os.path.join(d0, 'patch.npy'), patch)
    return patch_dim, radius, patch


# ---------------------------------------------------------------------------
# Loops
# ---------------------------------------------------------------------------
def pgd_attack(model, loss_fn, imgL, imgR, calib, iters, alpha, eps, norm='linf', on_iter=None):
    """The reference's PGD loop (attack/DSGN/pgd_attack.py:300-354) for a batch of
    pairs.  ``calib`` = (fu, baseline, Proj, Proj_R); ``loss_fn(outputs) ->
    scalar`` is the quantity ascended.  Returns (advL, advR, losses[iters])."""
    mean, std = IMAGENET_MEAN, IMAGENET_STD
    from .ops import _need_cuda as need
    need(imgL, imgR)
    m = torch.tensor(mean, device=imgL.device).view(1, 3, 1, 1)
    s = torch.tensor(std, device=imgL.device).view(1, 3, 1, 1)
    cleanL, cleanR = imgL * s + m, imgR * s + m          # denormalised clean copies (:297-298)
    xL, xR = imgL.clone(), imgR.clone()
    losses = []
    for it in range(iters):
        xL.requires_grad_(True)
        xR.requires_grad_(True)
        outputs = model(xL, xR, calib[0], calib[1], calib[2], calibs_Proj_R=calib[3])
        loss = loss_fn(outputs)
        gL, gR = torch.autograd.grad(loss, [xL, xR])
        xL, xR = xL.detach(), xR.detach()
        if norm == 'linf':
            xL, xR = pgd_step_pair(xL, gL.contiguous(), cleanL, xR, gR.contiguous(), cleanR, alpha, eps,
                                   inplace=True)
        else:
            xL = pgd_step(xL, gL.contiguous(), cleanL, alpha, eps, norm=norm, out=xL)
            xR = pgd_step(xR, gR.contiguous(), cleanR, alpha, eps, norm=norm, out=xR)
        losses.append(loss.detach())
        if on_iter is not None:
            on_iter(it, xL, xR)
    return xL, xR, torch.stack(losses)


def patch_attack_step(model, loss_fn, imgL, imgR, calib, patch, center_l, center_r, radius, iters=2,
                      alpha=1e3, eps=8.0 / 255, delta_hook=None, lo=None, hi=None):
    """Inner loop of the universal patch attack for one pair
    (attack/DSGN/patch_attack.py:367-430): blend, forward/backward, crop, clipped
    descent.  ``delta_hook(delta) -> delta`` lets the multi-GPU driver all-reduce
    the clipped step before it is applied."""
    losses = []
    for _ in range(iters):
        patch_apply(imgL, patch, center_l, radius)
        patch_apply(imgR, patch, center_r, radius)
        xL, xR = imgL.detach().requires_grad_(True), imgR.detach().requires_grad_(True)
        outputs = model(xL, xR, calib[0], calib[1], calib[2], calibs_Proj_R=calib[3])
        loss = loss_fn(outputs)
        gL, gR = torch.autograd.grad(loss, [xL, xR])
        if delta_hook is None:
            patch_update(patch, gL.contiguous(), gR.contiguous(), center_l, center_r, radius, alpha, eps, lo, hi)
        else:
            delta = torch.empty_like(patch)
            patch_update(patch, gL.contiguous(), gR.contiguous(), center_l, center_r, radius, alpha, eps,
                         delta_out=delta)
            patch_axpy(patch, delta_hook(delta), lo, hi)
        losses.append(loss.detach())
    return patch, torch.stack(losses)
