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


# Stock 2-D convolutions (cuDNN).  'tf32' = PyTorch's default on this hardware; 'tf32x3' = the same
# TF32 tensor-core kernels run on an error-compensated split (x = x_hi + x_lo, w = w_hi + w_lo,
# y ~ x_hi*w_hi + x_hi*w_lo + x_lo*w_hi, every factor exactly representable in TF32) -> fp32-class
# accuracy at 3x the (small) 2-D conv cost.  cuDNN's true-fp32 channels-last kernels are ~10x slower.
BACKBONE_PRECISION = os.environ.get("B2_BACKBONE", "tf32")


def set_backbone_precision(mode):
    global BACKBONE_PRECISION
    assert mode in ("tf32", "tf32x3")
    BACKBONE_PRECISION = mode


def _tf32_split(t):
    hi = (t.view(torch.int32) & ~0x1FFF).view(torch.float32)      # exactly TF32-representable
    return hi, t - hi                                              # lo has <= 13 significant bits


_W3_CACHE = {}


def _w3(w, dim):
    """[w_hi, w_lo, w_hi] stacked along ``dim`` (1: input channels for the forward, 0: output channels
    for the data gradient); cached per frozen weight."""
    key = (id(w), dim)
    hit = _W3_CACHE.get(key)
    if hit is not None and hit[0] is w and hit[1] == w._version:
        return hit[2]
    wh, wl = _tf32_split(w.detach())
    w3 = torch.cat([wh, wl, wh], dim).contiguous(memory_format=torch.channels_last)
    _W3_CACHE[key] = (w, w._version, w3)
    return w3


class Conv2dTf32x3Fn(torch.autograd.Function):
    """conv2d(x, w) with frozen w: three TF32 convolutions fused into one call by stacking the split
    operands along the input channels; the data gradient is the same trick on the output gradient."""

    @staticmethod
    def forward(ctx, x, w, bias, stride, padding, dilation):
        xh, xl = _tf32_split(x)
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
        gh, gl = _tf32_split(g)
        g3 = torch.cat([gh, gh, gl], 1).contiguous(memory_format=torch.channels_last)
        gx = torch.nn.grad.conv2d_input(shape, _w3(w, 0), g3, stride, padding, dilation)   # stacked along Cout
        return gx, None, None, None, None, None


def conv2d_stock(conv, x):
    if BACKBONE_PRECISION == "tf32x3" and x.is_cuda and conv.in_channels >= 16:
        return Conv2dTf32x3Fn.apply(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation)
    return conv(x)


def run_convbn_2d(seq, x, relu=False, res=None):
    """cuDNN conv2d (channels-last) -> our GroupNorm (+res) (+ReLU) kernel.
    ``seq`` = Sequential(Conv2d, GroupNorm)."""
    conv, norm = seq[0], seq[1]
    y = conv2d_stock(conv, x)
    return ops.groupnorm_act(y, norm.weight, norm.bias, norm.num_groups, norm.eps, relu=relu, res=res)


class BasicBlock(nn.Module):
    def __init__(self, cfg, cin, cout, stride, dilation):
        super().__init__()
        self.conv1 = nn.Sequential(_convbn(cfg, cin, cout, 3, stride, 1, dilation), nn.ReLU(inplace=True))
        self.conv2 = _convbn(cfg, cout, cout, 3, 1, 1, dilation)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), _gn(cfg, cout))

    def forward(self, x):
        out = run_convbn_2d(self.conv1[0], x, relu=True)
        short = x if self.downsample is None else run_convbn_2d(self.downsample, x)
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
    iteration measured); the GEMM form is deterministic and ~100x faster in backward."""
    ah = _interp_matrix(x.shape[-2], size[0], x.device)
    aw = _interp_matrix(x.shape[-1], size[1], x.device)
    return torch.matmul(torch.matmul(ah, x), aw.t())


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
        x = x.contiguous(memory_format=torch.channels_last)
        out = x
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
        cat = torch.cat(cat, 1).contiguous(memory_format=torch.channels_last)
        f = conv2d_stock(self.lastconv[2], run_convbn_2d(self.lastconv[0], cat, relu=True))
        n_rpn = cat.shape[0] if rpn_samples is None else rpn_samples
        r = conv2d_stock(self.rpnconv[2], run_convbn_2d(self.rpnconv[0], cat[:n_rpn], relu=True))
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
    calibs_Proj_R=)`` -> dict(depth_preds [N,H,W], bbox_cls, bbox_reg,
    bbox_centerness) -- attack/DSGN/pgd_attack.py:136, 308-323."""

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

    # the attack never needs parameter gradients
    def freeze(self):
        for p in self.parameters():
            p.requires_grad_(False)
        # 2-D parts run channels-last end to end (cuDNN NHWC kernels, our GroupNorm kernel)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)
        return self.eval()

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = {(k[7:] if k.startswith('module.') else k): v for k, v in state_dict.items()}
        return super().load_state_dict(sd, strict=strict, **kw)

    # -- cached calibration-only quantities ---------------------------------
    def _shifts(self, fu, baseline, device):
        key = (tuple(fu.flatten().tolist()), tuple(baseline.flatten().tolist()))
        hit = self._shift_cache.get(key)
        if hit is None:
            hit = plane_shifts(self.cfg, fu.cpu(), baseline.cpu()).to(device)
            self._shift_cache = {key: hit}
        return hit

    def _lift_plan(self, proj, psv_spatial, img_spatial, device):
        key = (proj.cpu().double().numpy().tobytes(), tuple(psv_spatial), tuple(img_spatial))
        hit = self._lift_cache.get(key)
        if hit is None:
            grid3 = lifting_grid(self.cfg, proj.cpu(), psv_spatial[1:]).to(device).contiguous()
            n, z, y, x, _ = grid3.shape
            grid2 = grid3[..., :2].contiguous().view(n, z * y, x, 2)
            hit = (grid3, ops.GridPlan(grid3, psv_spatial, True), ops.GridPlan(grid2, img_spatial, True))
            self._lift_cache = {key: hit}
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
        return conv2d_stock(self.bbox_cls, bev), conv2d_stock(self.bbox_reg, bev), conv2d_stock(self.bbox_centerness, bev)

    def forward(self, imgL, imgR, calibs_fu, calibs_baseline, calibs_Proj, calibs_Proj_R=None):
        if not imgL.is_cuda:
            raise RuntimeError("StereoNet (B200 path) needs CUDA inputs; there is no CPU fallback")
        # left and right images share the extractor weights: one batched pass (every kernel of the
        # launch-bound 2-D section does twice the work; GroupNorm statistics are per sample, so the
        # result equals two separate passes)
        n = imgL.shape[0]
        feat, rpnL = self.feature_extraction(torch.cat([imgL, imgR], 0), rpn_samples=n)
        featL, featR = feat[:n], feat[n:]
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
    stand-in for the unavailable upstream RPN3DLoss."""
    loss = 0.
    if cfg.loss_disp:
        pred = outputs['depth_preds']
        mask = ((disp_true > cfg.min_depth) & (disp_true <= cfg.max_depth)).to(pred.dtype)
        # mean over the valid pixels, written without boolean indexing (no host sync -> the
        # iteration stays CUDA-graph capturable); same value as smooth_l1(pred[mask], gt[mask]).mean()
        per = F.smooth_l1_loss(pred, disp_true, reduction='none')
        loss = loss + (per * mask).sum() / mask.sum().clamp_min(1.0)
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


def build_model(cfg=None, seed=1, device='cuda'):
    """Seeded default-PyTorch init (reference default seed 1, pgd_attack.py:41, 86),
    frozen, eval() (:140), on ``device``."""
    torch.manual_seed(seed)
    return StereoNet(cfg).freeze().to(device)
