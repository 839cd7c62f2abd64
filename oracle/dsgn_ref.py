"""CPU oracle (TEST INFRASTRUCTURE, not product code): a pure-PyTorch fp32
restatement of the DSGN-shaped detector the attack loop differentiates through.

PARITY UNPINNED.  The reference imports the model from the un-vendored,
un-pinned package Jia-Research-Lab/DSGN ("latest",
attack/DSGN/README.md:12; imports attack/DSGN/pgd_attack.py:27-33; constructor
:136; call :308; consumed keys :311, :322).  Nothing under /root/reference pins
its arithmetic, so this file restates the *published* architecture (PSMNet-style
2-D extractor, plane-sweep volume, 3-D hourglass, depth soft-argmin, frustum ->
voxel lifting by grid_sample, 3-D hourglass, BEV heads) using only stock torch
ops (Conv3d/ConvTranspose3d/GroupNorm/F.grid_sample/slice+lerp), and pins every
free choice here:

  * normalisation = GroupNorm(32, C) everywhere (cfg.GN); eval() mode.
  * depth planes (full res): z_j = min_depth + (j + 0.5) * depth_interval,
    j < maxdisp; PSV planes at 1/4: z_d = min_depth + (d + 0.5) * 4*interval.
  * cost volume: cost[:, :C, d] = L ; cost[:, C:, d, h, w] = R[h, w - s_d] with
    s_d = f_u * baseline / (z_d * downsample) feature px, linear interpolation
    for the fractional part, zero where the (integer part of the) source column
    would be < 0.  Left half is zeroed under the same validity mask.
  * depth head: trilinear x4 upsample (align_corners=False), softmax over
    depth, expectation over z_j.
  * lifting: voxel centres (x, y, z) -> P @ [x, y, z, 1] -> (u, v); the grid is
    normalised with align_corners=True against the 1/4-res feature map
    (u/4, v/4) and the PSV plane centres; zeros padding.
  * world grid X in [-30.4, 30.4), Y in [-1, 3), Z in [2, 40.4) at 0.2 m ->
    192 x 20 x 304 voxels (the survey's figure).
"""
import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F


def default_cfg(**over):
    cfg = SimpleNamespace(
        # attributes the reference scripts read from cfg
        # (attack/DSGN/pgd_attack.py:269, 310, 321)
        min_depth=2.0, max_depth=40.4, RPN3D_ENABLE=True, loss_disp=True,
        PlaneSweepVolume=True, GN=True, debug=False,
        # geometry
        maxdisp=192, downsample=4, depth_interval=0.2,
        x_range=(-30.4, 30.4), y_range=(-1.0, 3.0), z_range=(2.0, 40.4), voxel=0.2,
        # widths
        feat_ch=32, psv_ch=64, rpn_ch=32, gv_ch=64, bev_ch=128,
        backbone_blocks=(3, 16, 3, 3), spp_pools=(64, 32, 16, 8),
        y_pool=4, num_anchors=4, reg_dim=7, gn_groups=32,
    )
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def tiny_cfg(**over):
    """Spatially shrunk configuration with the real channel widths; image 32x64."""
    base = dict(maxdisp=32, min_depth=2.0, max_depth=8.4, depth_interval=0.2,
                x_range=(-3.2, 3.2), y_range=(-0.8, 0.8), z_range=(2.0, 8.4), voxel=0.4,
                backbone_blocks=(1, 1, 1, 1), spp_pools=(4, 2), y_pool=2)
    base.update(over)
    return default_cfg(**base)


def gn(cfg, c):
    return nn.GroupNorm(min(cfg.gn_groups, c), c)


def convbn(cfg, cin, cout, k, stride, pad, dilation=1):
    return nn.Sequential(
        nn.Conv2d(cin, cout, k, stride, dilation if dilation > 1 else pad, dilation, bias=False),
        gn(cfg, cout))


def convbn_3d(cfg, cin, cout, k=3, stride=1, pad=1):
    return nn.Sequential(nn.Conv3d(cin, cout, k, stride, pad, bias=False), gn(cfg, cout))


class BasicBlock(nn.Module):
    def __init__(self, cfg, cin, cout, stride, dilation):
        super().__init__()
        self.conv1 = nn.Sequential(convbn(cfg, cin, cout, 3, stride, 1, dilation), nn.ReLU(inplace=True))
        self.conv2 = convbn(cfg, cout, cout, 3, 1, 1, dilation)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), gn(cfg, cout))

    def forward(self, x):
        out = self.conv2(self.conv1(x))
        if self.downsample is not None:
            x = self.downsample(x)
        return out + x


class FeatureExtraction(nn.Module):
    """PSMNet-style 2-D extractor with SPP; two 32-ch heads at 1/4 resolution
    (matching features for the PSV, image features for the voxel lifting)."""

    def __init__(self, cfg):
        super().__init__()
        b = cfg.backbone_blocks
        self.firstconv = nn.Sequential(
            convbn(cfg, 3, 32, 3, 2, 1), nn.ReLU(inplace=True),
            convbn(cfg, 32, 32, 3, 1, 1), nn.ReLU(inplace=True),
            convbn(cfg, 32, 32, 3, 1, 1), nn.ReLU(inplace=True))
        self.layer1 = self._make(cfg, 32, 32, b[0], 1, 1)
        self.layer2 = self._make(cfg, 32, 64, b[1], 2, 1)
        self.layer3 = self._make(cfg, 64, 128, b[2], 1, 1)
        self.layer4 = self._make(cfg, 128, 128, b[3], 1, 2)
        self.branches = nn.ModuleList([
            nn.Sequential(nn.AvgPool2d(p, p), convbn(cfg, 128, 32, 1, 1, 0), nn.ReLU(inplace=True))
            for p in cfg.spp_pools])
        cat = 64 + 128 + 32 * len(cfg.spp_pools)
        self.lastconv = nn.Sequential(convbn(cfg, cat, 128, 3, 1, 1), nn.ReLU(inplace=True),
                                      nn.Conv2d(128, cfg.feat_ch, 1, bias=False))
        self.rpnconv = nn.Sequential(convbn(cfg, cat, 128, 3, 1, 1), nn.ReLU(inplace=True),
                                     nn.Conv2d(128, cfg.rpn_ch, 1, bias=False))

    @staticmethod
    def _make(cfg, cin, cout, n, stride, dilation):
        layers = [BasicBlock(cfg, cin, cout, stride, dilation)]
        layers += [BasicBlock(cfg, cout, cout, 1, dilation) for _ in range(n - 1)]
        return nn.Sequential(*layers)

    def forward(self, x):
        out = self.layer1(self.firstconv(x))
        raw = self.layer2(out)
        skip = self.layer4(self.layer3(raw))
        size = skip.shape[-2:]
        cat = [raw, skip] + [F.interpolate(br(skip), size, mode='bilinear', align_corners=False)
                             for br in self.branches]
        cat = torch.cat(cat, 1)
        return self.lastconv(cat), self.rpnconv(cat)


class Hourglass3d(nn.Module):
    """PSMNet hourglass: two stride-2 stages down, two transposed convs up."""

    def __init__(self, cfg, c):
        super().__init__()
        self.conv1 = nn.Sequential(convbn_3d(cfg, c, 2 * c, 3, 2, 1), nn.ReLU(inplace=True))
        self.conv2 = convbn_3d(cfg, 2 * c, 2 * c, 3, 1, 1)
        self.conv3 = nn.Sequential(convbn_3d(cfg, 2 * c, 2 * c, 3, 2, 1), nn.ReLU(inplace=True))
        self.conv4 = nn.Sequential(convbn_3d(cfg, 2 * c, 2 * c, 3, 1, 1), nn.ReLU(inplace=True))
        self.conv5 = nn.Sequential(
            nn.ConvTranspose3d(2 * c, 2 * c, 3, 2, 1, output_padding=1, bias=False), gn(cfg, 2 * c))
        self.conv6 = nn.Sequential(
            nn.ConvTranspose3d(2 * c, c, 3, 2, 1, output_padding=1, bias=False), gn(cfg, c))

    def forward(self, x):
        out = self.conv1(x)
        pre = F.relu(self.conv2(out))
        out = self.conv4(self.conv3(pre))
        post = F.relu(self.conv5(out) + pre)
        return self.conv6(post)


def psv_depths(cfg, device=None):
    step = cfg.depth_interval * cfg.downsample
    d = cfg.maxdisp // cfg.downsample
    return cfg.min_depth + (torch.arange(d, dtype=torch.float32, device=device) + 0.5) * step


def full_depths(cfg, device=None):
    return cfg.min_depth + (torch.arange(cfg.maxdisp, dtype=torch.float32, device=device) + 0.5) \
        * cfg.depth_interval


def plane_shifts(cfg, fu, baseline):
    """Disparity of each PSV plane in feature pixels: [N, D] fp32."""
    z = psv_depths(cfg)
    return (fu.float().view(-1, 1) * baseline.float().view(-1, 1) / (z.view(1, -1) * cfg.downsample))


def build_cost_volume(left, right, shifts):
    """Plane-sweep concat volume, stock ops only.  left/right [N,C,H,W], shifts
    [N,D] (feature px, >= 0) -> [N,2C,D,H,W]."""
    n, c, h, w = left.shape
    d = shifts.shape[1]
    cols = torch.arange(w, dtype=left.dtype, device=left.device)
    planes = []
    for i in range(d):
        s = shifts[:, i]
        s0 = torch.floor(s)
        frac = (s - s0).view(n, 1, 1, 1)
        x0 = cols.view(1, w) - s0.view(n, 1)            # source column of the integer part
        valid = (x0 >= 0).to(left.dtype).view(n, 1, 1, w)
        i0 = x0.clamp(0, w - 1).long()
        i1 = (x0 - 1).clamp(0, w - 1).long()
        v1 = ((x0 - 1) >= 0).to(left.dtype).view(n, 1, 1, w)
        r0 = torch.gather(right, 3, i0.view(n, 1, 1, w).expand(n, c, h, w))
        r1 = torch.gather(right, 3, i1.view(n, 1, 1, w).expand(n, c, h, w)) * v1
        r = ((1 - frac) * r0 + frac * r1) * valid
        planes.append(torch.cat([left * valid, r], 1))
    return torch.stack(planes, 2)


def voxel_grid(cfg, device=None):
    def centres(lo, hi):
        n = int(round((hi - lo) / cfg.voxel))
        return lo + (torch.arange(n, dtype=torch.float32, device=device) + 0.5) * cfg.voxel
    xs, ys, zs = centres(*cfg.x_range), centres(*cfg.y_range), centres(*cfg.z_range)
    z, y, x = torch.meshgrid(zs, ys, xs, indexing='ij')
    return torch.stack([x, y, z], -1)                   # [Z,Y,X,3]


def lifting_grid(cfg, proj, feat_hw):
    """Normalised sampling grid [N,Z,Y,X,3] (x=u, y=v, z=depth plane), align_corners=True."""
    pts = voxel_grid(cfg)                                # [Z,Y,X,3]
    ones = torch.ones_like(pts[..., :1])
    hom = torch.cat([pts, ones], -1)                     # [Z,Y,X,4]
    p = proj.float()                                     # [N,3,4]
    cam = torch.einsum('nij,zyxj->nzyxi', p, hom)        # [N,Z,Y,X,3]
    u = cam[..., 0] / cam[..., 2]
    v = cam[..., 1] / cam[..., 2]
    hf, wf = feat_hw
    zp = psv_depths(cfg)
    gu = 2.0 * (u / cfg.downsample) / (wf - 1) - 1.0
    gv = 2.0 * (v / cfg.downsample) / (hf - 1) - 1.0
    gz = 2.0 * (pts[..., 2].unsqueeze(0) - zp[0]) / (zp[-1] - zp[0]) - 1.0
    gz = gz.expand_as(gu)
    return torch.stack([gu, gv, gz], -1)


class StereoNetRef(nn.Module):
    """``model(imgL, imgR, calibs_fu, calibs_baseline, calibs_Proj, calibs_Proj_R=)``
    -> dict(depth_preds, bbox_cls, bbox_reg, bbox_centerness), the call shape of
    attack/DSGN/pgd_attack.py:308-323."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        c = cfg.psv_ch
        self.feature_extraction = FeatureExtraction(cfg)
        self.dres0 = nn.Sequential(convbn_3d(cfg, 2 * cfg.feat_ch, c), nn.ReLU(inplace=True),
                                   convbn_3d(cfg, c, c), nn.ReLU(inplace=True))
        self.dres1 = nn.Sequential(convbn_3d(cfg, c, c), nn.ReLU(inplace=True), convbn_3d(cfg, c, c))
        self.hg = Hourglass3d(cfg, c)
        self.classif1 = nn.Sequential(convbn_3d(cfg, c, c), nn.ReLU(inplace=True),
                                      nn.Conv3d(c, 1, 3, 1, 1, bias=False))
        g = cfg.gv_ch
        self.rpn3d_conv = nn.Sequential(convbn_3d(cfg, c + cfg.rpn_ch, g), nn.ReLU(inplace=True))
        self.rpn3d_hg = Hourglass3d(cfg, g)
        ny = int(round((cfg.y_range[1] - cfg.y_range[0]) / cfg.voxel)) // cfg.y_pool
        b = cfg.bev_ch
        self.bev_conv = nn.Sequential(convbn(cfg, g * ny, b, 3, 1, 1), nn.ReLU(inplace=True),
                                      convbn(cfg, b, b, 3, 1, 1), nn.ReLU(inplace=True))
        self.bbox_cls = nn.Conv2d(b, cfg.num_anchors, 3, 1, 1)
        self.bbox_reg = nn.Conv2d(b, cfg.num_anchors * cfg.reg_dim, 3, 1, 1)
        self.bbox_centerness = nn.Conv2d(b, cfg.num_anchors, 3, 1, 1)

    # -- stages (split so tests can compare intermediates) ------------------
    def psv_stage(self, featL, featR, fu, baseline):
        shifts = plane_shifts(self.cfg, fu, baseline).to(featL.dtype)
        cost = build_cost_volume(featL, featR, shifts)
        cost0 = self.dres0(cost)
        cost0 = self.dres1(cost0) + cost0
        out = self.hg(cost0) + cost0
        cost1 = self.classif1(out)
        return cost, out, cost1

    def depth_head(self, cost1, img_hw):
        cfg = self.cfg
        up = F.interpolate(cost1, [cfg.maxdisp, img_hw[0], img_hw[1]], mode='trilinear',
                           align_corners=False)
        prob = F.softmax(up.squeeze(1), 1)
        z = full_depths(cfg).view(1, -1, 1, 1).to(cost1.dtype)
        return (prob * z).sum(1)

    def lift(self, out, rpn_feat, proj):
        grid = lifting_grid(self.cfg, proj, out.shape[-2:]).to(out.dtype)
        vox = F.grid_sample(out, grid, mode='bilinear', padding_mode='zeros', align_corners=True)
        n, zz, yy, xx, _ = grid.shape
        g2 = grid[..., :2].reshape(n, zz * yy, xx, 2)
        vox2 = F.grid_sample(rpn_feat, g2, mode='bilinear', padding_mode='zeros', align_corners=True)
        vox2 = vox2.view(n, -1, zz, yy, xx)
        return torch.cat([vox, vox2], 1)

    def bev_stage(self, vox):
        cfg = self.cfg
        v = self.rpn3d_conv(vox)
        v = self.rpn3d_hg(v) + v
        v = F.avg_pool3d(v, (1, cfg.y_pool, 1))                    # [N,C,Z,Y/p,X]
        n, c, zz, yy, xx = v.shape
        bev = v.permute(0, 1, 3, 2, 4).reshape(n, c * yy, zz, xx)
        bev = self.bev_conv(bev)
        return self.bbox_cls(bev), self.bbox_reg(bev), self.bbox_centerness(bev)

    def forward(self, imgL, imgR, calibs_fu, calibs_baseline, calibs_Proj, calibs_Proj_R=None):
        featL, rpnL = self.feature_extraction(imgL)
        featR, _ = self.feature_extraction(imgR)
        cost, out, cost1 = self.psv_stage(featL, featR, calibs_fu, calibs_baseline)
        outputs = {'depth_preds': self.depth_head(cost1, imgL.shape[-2:])}
        if self.cfg.RPN3D_ENABLE:
            vox = self.lift(out, rpnL, calibs_Proj)
            cls, reg, ctr = self.bev_stage(vox)
            outputs.update(bbox_cls=cls, bbox_reg=reg, bbox_centerness=ctr)
        return outputs


def attack_loss(cfg, outputs, disp_true, labels):
    """Scalar the attack ascends (attack/DSGN/pgd_attack.py:310-331).  The depth
    term is the reference's: mask = (gt > min_depth) & (gt <= max_depth) (:269),
    smooth-L1 mean over valid pixels, weight 1.0 (:314-317, len(depth_preds)==1
    per pair).  The upstream RPN3DLoss is unavailable; the detection term is a
    fixed differentiable stand-in: focal-weighted BCE on bbox_cls
    against a seeded label map, smooth-L1 on bbox_reg and BCE on centerness at
    the positives."""
    loss = 0.
    if cfg.loss_disp:
        pred = outputs['depth_preds']
        mask = (disp_true > cfg.min_depth) & (disp_true <= cfg.max_depth)
        loss = loss + F.smooth_l1_loss(pred[mask], disp_true[mask], reduction='mean')
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
        loss = loss + (F.binary_cross_entropy_with_logits(ctr, labels['ctr'], reduction='none')
                       * tgt).sum() / npos
    return loss


def make_labels(cfg, n, seed, device='cpu'):
    """Seeded synthetic detection targets on the BEV grid [Z, X]."""
    g = torch.Generator().manual_seed(seed)
    zz = int(round((cfg.z_range[1] - cfg.z_range[0]) / cfg.voxel))
    xx = int(round((cfg.x_range[1] - cfg.x_range[0]) / cfg.voxel))
    cls = (torch.rand(n, cfg.num_anchors, zz, xx, generator=g) < 0.01).float()
    reg = torch.randn(n, cfg.num_anchors * cfg.reg_dim, zz, xx, generator=g)
    ctr = torch.rand(n, cfg.num_anchors, zz, xx, generator=g)
    return {'cls': cls.to(device), 'reg': reg.to(device), 'ctr': ctr.to(device)}


def build_model(cfg, seed=1):
    """Seeded default-PyTorch init (reference default seed 1,
    attack/DSGN/pgd_attack.py:41, 86); eval() mode like :140."""
    torch.manual_seed(seed)
    m = StereoNetRef(cfg)
    m.eval()
    return m
