"""B200 execution path of the DSGN-shaped detector the attack differentiates
through: same call signature and output dict as the reference's model call
(attack/DSGN/pgd_attack.py:308-323) and the same parameter names/shapes as the
stock-op restatement (so a state_dict loads into either, ``module.`` prefix of
the reference's DataParallel checkpoints accepted, pgd_attack.py:138-144).

What runs where:
  * plane-sweep cost volume, 3-D conv / deconv (+GroupNorm/ReLU/residual) of both
    hourglass stacks, frustum->voxel lifting: hand-written sm_100a kernels via
    ``ops`` (channels-last volumes end to end, no layout conversions);
  * 2-D feature extractor, depth soft-argmin head and BEV 2-D head: stock torch
    CUDA ops (outside the four subsystems of the north star; 8f "next" rows).
"""
import os
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


def default_cfg(**over):
    cfg = SimpleNamespace(
        min_depth=2.0, max_depth=40.4, RPN3D_ENABLE=True, loss_disp=True,
        PlaneSweepVolume=True, GN=True, debug=False,
        maxdisp=192, downsample=4, depth_interval=0.2,
        x_range=(-30.4, 30.4), y_range=(-1.0, 3.0), z_range=(2.0, 40.4), voxel=0.2,
        feat_ch=32, psv_ch=64, rpn_ch=32, gv_ch=64, bev_ch=128,
        backbone_blocks=(3, 16, 3, 3), spp_pools=(64, 32, 16, 8),
        y_pool=4, num_anchors=4, reg_dim=7, gn_groups=32,
    )
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def tiny_cfg(**over):
    base = dict(maxdisp=32, min_depth=2.0, max_depth=8.4, depth_interval=0.2,
                x_range=(-3.2, 3.2), y_range=(-0.8, 0.8), z_range=(2.0, 8.4), voxel=0.4,
                backbone_blocks=(1, 1, 1, 1), spp_pools=(4, 2), y_pool=2)
    base.update(over)
    return default_cfg(**base)


def _gn(cfg, c):
    return nn.GroupNorm(min(cfg.gn_groups, c), c)


def _convbn(cfg, cin, cout, k, stride, pad, dilation=1):
    return nn.Sequential(
        nn.Conv2d(cin, cout, k, stride, dilation if dilation > 1 else pad, dilation, bias=False),
        _gn(cfg, cout))


def _convbn_3d(cfg, cin, cout, k=3, stride=1, pad=1):
    return nn.Sequential(nn.Conv3d(cin, cout, k, stride, pad, bias=False), _gn(cfg, cout))


def run_convbn_3d(seq, x, relu=False, res=None, fork=False):
    """conv3d/deconv3d -> GroupNorm (+res) (+ReLU) through the sm_100a kernels.
    ``seq`` = Sequential(Conv3d | ConvTranspose3d, GroupNorm) used as parameter holder.
    ``fork``: x has other consumers too; returns (y, x2) and those consumers must read x2 -- their
    gradient is then added inside this conv's data-gradient kernel (``ops.conv3d_fork``)."""
    conv, norm = seq[0], seq[1]
    transposed = isinstance(conv, nn.ConvTranspose3d)
    # the conv epilogue adds up the GroupNorm statistics of its own output where the kernel supports it
    if fork:
        y, part, x2 = ops.conv3d_fork(x, conv.weight, stride=conv.stride[0], transposed=transposed)
    else:
        y, part = ops.conv3d_with_stats(x, conv.weight, stride=conv.stride[0], transposed=transposed)
    y = ops.groupnorm_act(y, norm.weight, norm.bias, norm.num_groups, norm.eps, relu=relu, res=res, partial=part)
    return (y, x2) if fork else y


# 2-D convolutions.  'b2' (default): our tcgen05 kernels with the error-compensated 3xTF32 split done
# in-kernel (fp32-class accuracy, ops.conv2d).  'cudnn' / 'cudnn_tf32x3': the stock cuDNN kernels of
# round 1, kept for A/B measurements only (TF32 per torch.backends flags / split stacked on the host).
BACKBONE_IMPL = os.environ.get("B2_BACKBONE", "b2")
# concat / split glue of the extractor and the heads through the one-launch ops (0: stock torch.cat / slicing, A/B only)
GLUE_OPS = os.environ.get("B2_GLUE_OPS", "1") != "0"


def set_backbone_impl(mode):
    global BACKBONE_IMPL
    assert mode in ("b2", "cudnn", "cudnn_tf32x3")
    BACKBONE_IMPL = mode


_W3_CACHE = {}


def _w3(w, dim):
    """[w_hi, w_lo, w_hi] stacked along ``dim`` (cuDNN A/B mode only)."""
    key = (id(w), dim)
    hit = _W3_CACHE.get(key)
    if hit is not None and hit[0] is w and hit[1] == w._version:
        return hit[2]
    wh, wl = ops.tf32_split(w.detach())
    w3 = torch.cat([wh, wl, wh], dim).contiguous(memory_format=torch.channels_last)
    _W3_CACHE[key] = (w, w._version, w3)
    return w3


class Conv2dTf32x3Fn(torch.autograd.Function):
    """A/B mode 'cudnn_tf32x3': three TF32 cuDNN convolutions fused into one call by stacking the split
    operands along the input channels (round 1's high-fidelity mode)."""

    @staticmethod
    def forward(ctx, x, w, bias, stride, padding, dilation):
        xh, xl = ops.tf32_split(x)
        x3 = torch.cat([xh, xh, xl], 1).contiguous(memory_format=torch.channels_last)
        y = F.conv2d(x3, _w3(w, 1), bias, stride, padding, dilation)
        ctx.save_for_backward(w)
        ctx.cfg = (tuple(x.shape), stride, padding, dilation)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (w,) = ctx.saved_tensors
        shape, stride, padding, dilation = ctx.cfg
        gh, gl = ops.tf32_split(g)
        g3 = torch.cat([gh, gh, gl], 1).contiguous(memory_format=torch.channels_last)
        gx = torch.nn.grad.conv2d_input(shape, _w3(w, 0), g3, stride, padding, dilation)
        return gx, None, None, None, None, None


def conv2d_any(conv, x, fork=False, stats=False):
    """``conv`` = nn.Conv2d used as parameter holder.  Returns y, or (y, x2) with ``fork`` (x2 = x for its
    other consumers; with our kernels their gradient is added inside the data-gradient launch).
    ``stats``: (y, partial[, x2]) -- the conv epilogue's GroupNorm partial sums of y (None where not available)."""
    if BACKBONE_IMPL == "b2":
        k, pad, dil = conv.kernel_size[0], conv.padding[0], conv.dilation[0]
        if conv.kernel_size[0] != conv.kernel_size[1] or pad != dil * (k // 2) or conv.groups != 1 \
                or conv.stride[0] != conv.stride[1]:
            raise RuntimeError("conv2d: unsupported layer %r" % (conv,))
        if stats:
            return ops.conv2d_with_stats(x, conv.weight, conv.bias, conv.stride[0], dil, fork=fork)
        if fork:
            return ops.conv2d_fork(x, conv.weight, conv.bias, conv.stride[0], dil)
        return ops.conv2d(x, conv.weight, conv.bias, conv.stride[0], dil)
    x = x.contiguous(memory_format=torch.channels_last)
    if BACKBONE_IMPL == "cudnn_tf32x3" and conv.in_channels >= 16:
        y = Conv2dTf32x3Fn.apply(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation)
    else:
        y = conv(x)
    if stats:
        return (y, None, x) if fork else (y, None)
    return (y, x) if fork else y


def run_convbn_2d(seq, x, relu=False, res=None, fork=False):
    """conv2d -> GroupNorm (+res) (+ReLU), all on the sm_100a kernels.  ``seq`` = Sequential(Conv2d, GroupNorm)."""
    conv, norm = seq[0], seq[1]
    # the conv epilogue adds up the GroupNorm statistics of its own output (no statistics pass over y), and the
    # norm's output carries a GnLink: the data-gradient launch of its consumer adds up the backward sums as well
    out = conv2d_any(conv, x, fork, stats=True)
    y, part = out[0], out[1]
    y = ops.groupnorm_act(y, norm.weight, norm.bias, norm.num_groups, norm.eps, relu=relu, res=res, partial=part)
    return (y, out[2]) if fork else y


class BasicBlock(nn.Module):
    def __init__(self, cfg, cin, cout, stride, dilation):
        super().__init__()
        self.conv1 = nn.Sequential(_convbn(cfg, cin, cout, 3, stride, 1, dilation), nn.ReLU(inplace=True))
        self.conv2 = _convbn(cfg, cout, cout, 3, 1, 1, dilation)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), _gn(cfg, cout))

    def forward(self, x):
        # x feeds conv1 AND the shortcut: conv1 forks it, so the shortcut's gradient is added inside conv1's
        # data-gradient kernel (no autograd accumulation pass)
        out, x2 = run_convbn_2d(self.conv1[0], x, relu=True, fork=True)
        short = x2 if self.downsample is None else run_convbn_2d(self.downsample, x2)
        return run_convbn_2d(self.conv2, out, relu=False, res=short)      # GN(conv2) + shortcut, one pass


_INTERP_CACHE = {}


def _interp_matrix(n_in, n_out, device):
    """[n_out, n_in] bilinear (align_corners=False) interpolation matrix, ATen's
    area_pixel_compute_source_index rule: src = (dst + .5) * in/out - .5, clamped at 0."""
    key = (n_in, n_out, str(device))
    m = _INTERP_CACHE.get(key)
    if m is None:
        dst = torch.arange(n_out, dtype=torch.float32)
        src = ((dst + 0.5) * (n_in / n_out) - 0.5).clamp_min(0.0)
        i0 = src.floor().long().clamp_max(n_in - 1)
        i1 = (i0 + 1).clamp_max(n_in - 1)
        w1 = src - i0.float()
        m = torch.zeros(n_out, n_in)
        m.scatter_add_(1, i0.view(-1, 1), (1 - w1).view(-1, 1))
        m.scatter_add_(1, i1.view(-1, 1), w1.view(-1, 1))
        m = m.to(device)
        _INTERP_CACHE[key] = m
    return m


def upsample_bilinear_matmul(x, size):
    """F.interpolate(x, size, 'bilinear', align_corners=False) as two small GEMMs: the SPP maps
    are tiny (1x4 .. 12x39), and ATen's upsample backward serialises on atomics there (7.6 ms per
    iteration measured); the GEMM form is deterministic and ~100x faster in backward.
    Both GEMMs run on the channels-last memory as it is ([N*h, w, C] then [N, h, X*C] are contiguous views), so
    neither direction makes a layout copy and the result is a channels-last map."""
    n, c, h, w = x.shape
    ah = _interp_matrix(h, size[0], x.device)
    aw = _interp_matrix(w, size[1], x.device)
    xc = x.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)        # [N, h, w, C], contiguous
    t = torch.matmul(aw, xc.reshape(n * h, w, c))                                   # [N*h, X, C]
    u = torch.matmul(ah, t.view(n, h, size[1] * c))                                 # [N, Y, X*C]
    return u.view(n, size[0], size[1], c).permute(0, 3, 1, 2)


class FeatureExtraction(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        b = cfg.backbone_blocks
        self.firstconv = nn.Sequential(
            _convbn(cfg, 3, 32, 3, 2, 1), nn.ReLU(inplace=True),
            _convbn(cfg, 32, 32, 3, 1, 1), nn.ReLU(inplace=True),
            _convbn(cfg, 32, 32, 3, 1, 1), nn.ReLU(inplace=True))
        self.layer1 = self._make(cfg, 32, 32, b[0], 1, 1)
        self.layer2 = self._make(cfg, 32, 64, b[1], 2, 1)
        self.layer3 = self._make(cfg, 64, 128, b[2], 1, 1)
        self.layer4 = self._make(cfg, 128, 128, b[3], 1, 2)
        self.branches = nn.ModuleList([
            nn.Sequential(nn.AvgPool2d(p, p), _convbn(cfg, 128, 32, 1, 1, 0), nn.ReLU(inplace=True))
            for p in cfg.spp_pools])
        cat = 64 + 128 + 32 * len(cfg.spp_pools)
        self.lastconv = nn.Sequential(_convbn(cfg, cat, 128, 3, 1, 1), nn.ReLU(inplace=True),
                                      nn.Conv2d(128, cfg.feat_ch, 1, bias=False))
        self.rpnconv = nn.Sequential(_convbn(cfg, cat, 128, 3, 1, 1), nn.ReLU(inplace=True),
                                     nn.Conv2d(128, cfg.rpn_ch, 1, bias=False))

    @staticmethod
    def _make(cfg, cin, cout, n, stride, dilation):
        layers = [BasicBlock(cfg, cin, cout, stride, dilation)]
        layers += [BasicBlock(cfg, cout, cout, 1, dilation) for _ in range(n - 1)]
        return nn.Sequential(*layers)

    def forward(self, x, rpn_samples=None):
        """``rpn_samples``: compute the image-feature head only for the first that many samples
        (the model runs left and right images as one batch; only the left needs that head)."""
        out = x                                   # NCHW image batch: the first layer's kernel reads it as is
        for i in (0, 2, 4):
            out = run_convbn_2d(self.firstconv[i], out, relu=True)
        out = self.layer1(out)
        raw = self.layer2(out)
        skip = self.layer4(self.layer3(raw))
        size = skip.shape[-2:]
        # SPP: the pools are nested (every window size is a multiple of the smallest), so pool
        # once by the smallest and re-pool that -- same windows, no 64x64-wide serial loops
        pools = [br[0].kernel_size if isinstance(br[0].kernel_size, int) else br[0].kernel_size[0]
                 for br in self.branches]
        base = min(pools)
        cat = [raw, skip]
        if all(p % base == 0 for p in pools):
            pooled0 = F.avg_pool2d(skip, base, base)
            for br, p in zip(self.branches, pools):
                y = pooled0 if p == base else F.avg_pool2d(pooled0, p // base, p // base)
                y = run_convbn_2d(br[1], y, relu=True)
                cat.append(upsample_bilinear_matmul(y, size))
        else:
            for br in self.branches:
                y = run_convbn_2d(br[1], br[0](skip), relu=True)
                cat.append(upsample_bilinear_matmul(y, size))
        n_rpn = cat[0].shape[0] if rpn_samples is None else rpn_samples
        if BACKBONE_IMPL == "b2" and GLUE_OPS:
            # one launch for the concatenation and one for its backward (contiguous gradient slices); the gradient
            # of the first n_rpn samples coming from the second head is merged in one launch as well
            cat = ops.cat_channels(cat)
            cat, cat_rpn = ops.prefix_fork(cat, n_rpn) if n_rpn < cat.shape[0] else (cat, cat)
        else:
            cat = torch.cat(cat, 1).contiguous(memory_format=torch.channels_last)
            cat_rpn = cat[:n_rpn]
        f = conv2d_any(self.lastconv[2], run_convbn_2d(self.lastconv[0], cat, relu=True))
        r = conv2d_any(self.rpnconv[2], run_convbn_2d(self.rpnconv[0], cat_rpn, relu=True))
        return f, r


class Hourglass3d(nn.Module):
    def __init__(self, cfg, c):
        super().__init__()
        self.conv1 = nn.Sequential(_convbn_3d(cfg, c, 2 * c, 3, 2, 1), nn.ReLU(inplace=True))
        self.conv2 = _convbn_3d(cfg, 2 * c, 2 * c, 3, 1, 1)
        self.conv3 = nn.Sequential(_convbn_3d(cfg, 2 * c, 2 * c, 3, 2, 1), nn.ReLU(inplace=True))
        self.conv4 = nn.Sequential(_convbn_3d(cfg, 2 * c, 2 * c, 3, 1, 1), nn.ReLU(inplace=True))
        self.conv5 = nn.Sequential(
            nn.ConvTranspose3d(2 * c, 2 * c, 3, 2, 1, output_padding=1, bias=False), _gn(cfg, 2 * c))
        self.conv6 = nn.Sequential(
            nn.ConvTranspose3d(2 * c, c, 3, 2, 1, output_padding=1, bias=False), _gn(cfg, c))

    def forward(self, x, res):
        """returns conv6(...) + res (the caller's residual fused into the last norm).
        x and pre each feed a conv AND a residual: the conv forks its input so that the residual's
        gradient is added inside the conv's data-gradient kernel (no separate accumulation pass)."""
        out, x2 = run_convbn_3d(self.conv1[0], x, relu=True, fork=True)
        if res is x:
            res = x2
        pre = run_convbn_3d(self.conv2, out, relu=True)
        out, pre2 = run_convbn_3d(self.conv3[0], pre, relu=True, fork=True)
        out = run_convbn_3d(self.conv4[0], out, relu=True)
        post = run_convbn_3d(self.conv5, out, relu=True, res=pre2)
        return run_convbn_3d(self.conv6, post, relu=False, res=res)


def psv_depths(cfg, device=None):
    step = cfg.depth_interval * cfg.downsample
    d = cfg.maxdisp // cfg.downsample
    return cfg.min_depth + (torch.arange(d, dtype=torch.float32, device=device) + 0.5) * step


def full_depths(cfg, device=None):
    return cfg.min_depth + (torch.arange(cfg.maxdisp, dtype=torch.float32, device=device) + 0.5) \
        * cfg.depth_interval


def plane_shifts(cfg, fu, baseline):
    z = psv_depths(cfg)
    return (fu.float().view(-1, 1) * baseline.float().view(-1, 1) / (z.view(1, -1) * cfg.downsample))


def voxel_grid(cfg):
    def centres(lo, hi):
        n = int(round((hi - lo) / cfg.voxel))
        return lo + (torch.arange(n, dtype=torch.float32) + 0.5) * cfg.voxel
    xs, ys, zs = centres(*cfg.x_range), centres(*cfg.y_range), centres(*cfg.z_range)
    z, y, x = torch.meshgrid(zs, ys, xs, indexing='ij')
    return torch.stack([x, y, z], -1)


def lifting_grid(cfg, proj, feat_hw):
    """[N,Z,Y,X,3] normalised (u, v, plane) grid; computed on the host in fp32 (a
    function of the calibration only) and cached per calibration by the model."""
    pts = voxel_grid(cfg)
    hom = torch.cat([pts, torch.ones_like(pts[..., :1])], -1)
    cam = torch.einsum('nij,zyxj->nzyxi', proj.float(), hom)
    u = cam[..., 0] / cam[..., 2]
    v = cam[..., 1] / cam[..., 2]
    hf, wf = feat_hw
    zp = psv_depths(cfg)
    gu = 2.0 * (u / cfg.downsample) / (wf - 1) - 1.0
    gv = 2.0 * (v / cfg.downsample) / (hf - 1) - 1.0
    gz = 2.0 * (pts[..., 2].unsqueeze(0) - zp[0]) / (zp[-1] - zp[0]) - 1.0
    return torch.stack([gu, gv, gz.expand_as(gu)], -1)


class StereoNet(nn.Module):
    """``StereoNet(cfg)(imgL, imgR, calibs_fu, calibs_baseline, calibs_Proj,
    calibs_Proj_R=)`` -> dict(depth_preds, bbox_cls, bbox_reg, bbox_centerness)
    -- attack/DSGN/pgd_attack.py:136, 308-323.

    ``depth_preds`` follows upstream's EVAL-mode convention (the attack scripts call ``model.eval()``,
    pgd_attack.py:140): ONE [N,H,W] tensor, not the training-mode list of per-scale maps.  The
    reference's loss lines (:311-317) iterate it, i.e. they walk the batch dimension -- correct only
    for the batch size of 1 the scripts use; ``attack_loss`` documents how N > 1 is treated here."""

    def __init__(self, cfg=None):
        super().__init__()
        cfg = cfg or default_cfg()
        self.cfg = cfg
        c = cfg.psv_ch
        self.feature_extraction = FeatureExtraction(cfg)
        self.dres0 = nn.Sequential(_convbn_3d(cfg, 2 * cfg.feat_ch, c), nn.ReLU(inplace=True),
                                   _convbn_3d(cfg, c, c), nn.ReLU(inplace=True))
        self.dres1 = nn.Sequential(_convbn_3d(cfg, c, c), nn.ReLU(inplace=True), _convbn_3d(cfg, c, c))
        self.hg = Hourglass3d(cfg, c)
        self.classif1 = nn.Sequential(_convbn_3d(cfg, c, c), nn.ReLU(inplace=True),
                                      nn.Conv3d(c, 1, 3, 1, 1, bias=False))
        g = cfg.gv_ch
        self.rpn3d_conv = nn.Sequential(_convbn_3d(cfg, c + cfg.rpn_ch, g), nn.ReLU(inplace=True))
        self.rpn3d_hg = Hourglass3d(cfg, g)
        ny = int(round((cfg.y_range[1] - cfg.y_range[0]) / cfg.voxel)) // cfg.y_pool
        b = cfg.bev_ch
        self.bev_conv = nn.Sequential(_convbn(cfg, g * ny, b, 3, 1, 1), nn.ReLU(inplace=True),
                                      _convbn(cfg, b, b, 3, 1, 1), nn.ReLU(inplace=True))
        self.bbox_cls = nn.Conv2d(b, cfg.num_anchors, 3, 1, 1)
        self.bbox_reg = nn.Conv2d(b, cfg.num_anchors * cfg.reg_dim, 3, 1, 1)
        self.bbox_centerness = nn.Conv2d(b, cfg.num_anchors, 3, 1, 1)
        self._lift_cache = {}
        self._shift_cache = {}
        self._head_cache = None

    # the attack never needs parameter gradients
    def freeze(self):
        for p in self.parameters():
            p.requires_grad_(False)
        # (cuDNN A/B modes only: channels-last weights for its NHWC kernels)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)
        return self.eval()

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = {(k[7:] if k.startswith('module.') else k): v for k, v in state_dict.items()}
        return super().load_state_dict(sd, strict=strict, **kw)

    # -- cached calibration-only quantities (a few calibrations stay resident: KITTI has one per recording
    # day; building an entry syncs with the host, so it must happen outside any graph capture -- the engine
    # warms every new calibration up eagerly before it captures) ---------------------------------
    CALIB_CACHE = 4

    def _shifts(self, fu, baseline, device):
        key = (tuple(fu.flatten().tolist()), tuple(baseline.flatten().tolist()))
        hit = self._shift_cache.get(key)
        if hit is None:
            hit = plane_shifts(self.cfg, fu.cpu(), baseline.cpu()).to(device)
            if len(self._shift_cache) >= self.CALIB_CACHE:
                self._shift_cache.pop(next(iter(self._shift_cache)))
            self._shift_cache[key] = hit
        return hit

    def _lift_plan(self, proj, psv_spatial, img_spatial, device):
        key = (proj.cpu().double().numpy().tobytes(), tuple(psv_spatial), tuple(img_spatial))
        hit = self._lift_cache.get(key)
        if hit is None:
            grid3 = lifting_grid(self.cfg, proj.cpu(), psv_spatial[1:]).to(device).contiguous()
            n, z, y, x, _ = grid3.shape
            grid2 = grid3[..., :2].contiguous().view(n, z * y, x, 2)
            hit = (grid3, ops.GridPlan(grid3, psv_spatial, True), ops.GridPlan(grid2, img_spatial, True))
            if len(self._lift_cache) >= self.CALIB_CACHE:
                self._lift_cache.pop(next(iter(self._lift_cache)))
            self._lift_cache[key] = hit
        return hit

    # -- stages --------------------------------------------------------------
    def psv_stage(self, featL, featR, fu, baseline):
        shifts = self._shifts(fu, baseline, featL.device)
        cost = ops.build_cost_volume(featL, featR, shifts, channels_last=True)
        x = run_convbn_3d(self.dres0[0], cost, relu=True)
        cost0 = run_convbn_3d(self.dres0[2], x, relu=True)
        x, cost0 = run_convbn_3d(self.dres1[0], cost0, relu=True, fork=True)     # cost0 also feeds the residual
        cost0 = run_convbn_3d(self.dres1[2], x, relu=False, res=cost0)
        out = self.hg(cost0, res=cost0)
        x, out = run_convbn_3d(self.classif1[0], out, relu=True, fork=True)      # out also feeds the lifting
        cost1 = ops.conv3d_c1(x, self.classif1[2].weight)
        return cost, out, cost1

    def depth_head(self, cost1, img_hw):
        cfg = self.cfg
        return ops.depth_head(cost1, (cfg.maxdisp, img_hw[0], img_hw[1]), cfg.min_depth, cfg.depth_interval)

    def lift(self, out, rpn_feat, proj):
        grid3, plan3, plan2 = self._lift_plan(proj, out.shape[2:], rpn_feat.shape[2:], out.device)
        return ops.lift(out, rpn_feat, grid3, plan3, plan2, align_corners=True)

    def bev_stage(self, vox):
        cfg = self.cfg
        v = run_convbn_3d(self.rpn3d_conv[0], vox, relu=True)
        v = self.rpn3d_hg(v, res=v)
        bev = ops.bev_pool(v, cfg.y_pool)          # AvgPool3d over Y + (C, Y/p) -> channels, one pass
        bev = run_convbn_2d(self.bev_conv[0], bev, relu=True)
        bev = run_convbn_2d(self.bev_conv[2], bev, relu=True)
        return self.heads(bev)

    def heads(self, bev):
        """bbox_cls / bbox_reg / bbox_centerness (three 3x3 convs with bias on the same map) as ONE
        128 -> 64 conv (4 + 28 + 4 channels, zero-padded to the tensor core's N granularity)."""
        convs = (self.bbox_cls, self.bbox_reg, self.bbox_centerness)
        if BACKBONE_IMPL != "b2":
            return tuple(conv2d_any(c, bev) for c in convs)
        key = tuple((c.weight._version, c.weight.data_ptr(), c.bias._version) for c in convs)
        if self._head_cache is None or self._head_cache[0] != key:
            widths = [c.out_channels for c in convs]
            tot = sum(widths)
            pad = (-tot) % 32
            w = torch.cat([c.weight.detach() for c in convs] +
                          [convs[0].weight.new_zeros(pad, *convs[0].weight.shape[1:])], 0).contiguous()
            b = torch.cat([c.bias.detach() for c in convs] + [convs[0].bias.new_zeros(pad)], 0).contiguous()
            self._head_cache = (key, w, b, widths)
        _, w, b, widths = self._head_cache
        y = ops.conv2d(bev, w, b, 1, 1)
        if GLUE_OPS:
            return ops.split_channels(y, widths)   # views; their gradients are assembled in one launch
        outs, o = [], 0
        for n in widths:
            outs.append(y[:, o:o + n])
            o += n
        return tuple(outs)

    def forward(self, imgL, imgR, calibs_fu, calibs_baseline, calibs_Proj, calibs_Proj_R=None):
        if not imgL.is_cuda:
            raise RuntimeError("StereoNet (B200 path) needs CUDA inputs; there is no CPU fallback")
        # left and right images share the extractor weights: one batched pass (every kernel of the
        # launch-bound 2-D section does twice the work; GroupNorm statistics are per sample, so the
        # result equals two separate passes)
        n = imgL.shape[0]
        feat, rpnL = self.feature_extraction(torch.cat([imgL, imgR], 0), rpn_samples=n)
        featL, featR = ops.split_batch(feat, n) if GLUE_OPS else (feat[:n], feat[n:])
        cost, out, cost1 = self.psv_stage(featL, featR, calibs_fu, calibs_baseline)
        outputs = {'depth_preds': self.depth_head(cost1, imgL.shape[-2:])}
        if self.cfg.RPN3D_ENABLE:
            vox = self.lift(out, rpnL, calibs_Proj)
            cls, reg, ctr = self.bev_stage(vox)
            outputs.update(bbox_cls=cls, bbox_reg=reg, bbox_centerness=ctr)
        return outputs


def attack_loss(cfg, outputs, disp_true, labels):
    """Scalar ascended by the attack: the reference's depth term
    (attack/DSGN/pgd_attack.py:269, 310-319) plus the fixed differentiable
    stand-in for the unavailable upstream RPN3DLoss.

    ``outputs['depth_preds']`` is the eval-mode [N,H,W] tensor (see ``StereoNet.forward``).  The reference's
    lines iterate it (``for o in outputs['depth_preds']``), which for its batch size of 1 yields the single
    [H,W] map with weight ``[0.5, 0.7, 1.0][2]`` = 1.0 and ``mask[0]``; for N > 1 those lines would weight
    the SAMPLES 0.5/0.7/1.0 and mask every sample with sample 0's mask.  Here every pair
    gets what the reference gives its one pair: weight 1.0, its own mask, mean over its valid pixels;
    the per-pair terms are summed.  tests/test_gpu_e2e.py executes the reference's lines on this dict."""
    loss = 0.
    if cfg.loss_disp:
        pred = outputs['depth_preds']
        mask = ((disp_true > cfg.min_depth) & (disp_true <= cfg.max_depth)).to(pred.dtype)
        # mean over the valid pixels, written without boolean indexing (no host sync -> the
        # iteration stays CUDA-graph capturable); same value as smooth_l1(pred[mask], gt[mask]).mean()
        per = F.smooth_l1_loss(pred, disp_true, reduction='none')
        num = (per * mask).flatten(1).sum(1)
        den = mask.flatten(1).sum(1).clamp_min(1.0)
        loss = loss + (num / den).sum()
    if cfg.RPN3D_ENABLE:
        cls, reg, ctr = outputs['bbox_cls'], outputs['bbox_reg'], outputs['bbox_centerness']
        tgt = labels['cls']
        p = torch.sigmoid(cls)
        bce = F.binary_cross_entropy_with_logits(cls, tgt, reduction='none')
        focal = (tgt * (1 - p) ** 2 * 0.25 + (1 - tgt) * p ** 2 * 0.75) * bce
        npos = tgt.sum().clamp_min(1.0)
        loss = loss + focal.sum() / npos
        pos = tgt.repeat_interleave(cfg.reg_dim, 1)
        loss = loss + (F.smooth_l1_loss(reg, labels['reg'], reduction='none') * pos).sum() / npos
        loss = loss + (F.binary_cross_entropy_with_logits(ctr, labels['ctr'], reduction='none') * tgt).sum() / npos
    return loss


def decode_detections(cfg, outputs, proj, score_thresh=0.5, topk=20, cls_id=2):
    """Stand-in for the upstream post-processor the reference calls before ``kitti_output``
    (attack/DSGN/predict_and_save_pgd.py:220-283; un-vendored): the inverse of the stand-in target assignment
    ``synthetic.labels_from_box3d`` -- BEV cells whose best anchor's sigmoid score passes ``score_thresh`` (at most
    ``topk``, highest first) become boxes (h, w, l, centre, ry) from that anchor's 7 regression channels; the 2-D box is
    the projection of the 8 corners with ``proj`` [3,4].  Returns a list of dicts in the format
    ``kitti_io.write_detections`` takes (cls 2 = 'Car', the class of the reference's fake ground truth).
    One host synchronisation: meant for the hand-off after an attack, not for the iteration."""
    import math
    cls = torch.sigmoid(outputs['bbox_cls'][0].float())                 # [A, Z, X]
    reg = outputs['bbox_reg'][0].float()
    a, zz, xx = cls.shape
    score, anchor = cls.max(0)
    flat = score.flatten()
    k = min(int(topk), flat.numel())
    top, idx = torch.topk(flat, k)
    keep = top >= score_thresh
    top, idx, anchor = top[keep].cpu(), idx[keep].cpu(), anchor.flatten()[idx[keep]].cpu()
    reg = reg.cpu()
    P = torch.as_tensor(proj, dtype=torch.float64).reshape(3, 4)
    dets = []
    for s, i, an in zip(top.tolist(), idx.tolist(), anchor.tolist()):
        iz, ix = divmod(i, xx)
        zc = cfg.z_range[0] + (iz + 0.5) * cfg.voxel
        xc = cfg.x_range[0] + (ix + 0.5) * cfg.voxel
        r = reg[an * cfg.reg_dim:(an + 1) * cfg.reg_dim, iz, ix].tolist()
        x, y, z = xc + r[0], r[1], zc + r[2]
        h, w, l = (math.exp(min(max(v, -4.0), 4.0)) for v in r[3:6])
        ry = an * math.pi / 2 + r[6]
        # 8 corners (KITTI camera frame: x right, y down, z forward; the box stands on y + h/2)
        cs, sn = math.cos(ry), math.sin(ry)
        corners = []
        for dx in (-l / 2, l / 2):
            for dz in (-w / 2, w / 2):
                for dy in (-h / 2, h / 2):
                    corners.append((x + cs * dx + sn * dz, y + dy, z - sn * dx + cs * dz))
        pts = torch.tensor(corners, dtype=torch.float64)
        uvw = torch.cat([pts, torch.ones(8, 1, dtype=torch.float64)], 1) @ P.t()
        zc_ = uvw[:, 2].clamp_min(1e-3)
        u, v = uvw[:, 0] / zc_, uvw[:, 1] / zc_
        dets.append(dict(cls=cls_id, bbox=[u.min().item(), v.min().item(), u.max().item(), v.max().item()],
                         hwl=(h, w, l), center3d=(x, y, z), ry=ry, score=s))
    return dets


def build_model(cfg=None, seed=1, device='cuda'):
    """Seeded default-PyTorch init (reference default seed 1, pgd_attack.py:41, 86),
    frozen, eval() (:140), on ``device``."""
    torch.manual_seed(seed)
    return StereoNet(cfg).freeze().to(device)
