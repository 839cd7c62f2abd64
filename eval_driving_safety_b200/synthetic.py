"""Synthetic KITTI-shaped stereo pairs: there is no dataset access.

Pair i, seed 1000+i: imgL01 ~ U[0,1) [3,H,W] smoothed with a 5x5 box filter,
imgR01 = roll(imgL01, -d, W) with d in {8..64}, ImageNet-normalised; depth GT
~ U(2, 40.4) with 30 % zeros (invalid); KITTI nominal calibration.
"""
import torch
import torch.nn.functional as F

MEAN = [0.485, 0.456, 0.406]   # attack/DSGN/pgd_attack.py:153
STD = [0.229, 0.224, 0.225]    # attack/DSGN/pgd_attack.py:154
KITTI_F, KITTI_CU, KITTI_CV, KITTI_B = 721.5377, 609.5593, 172.854, 0.54


def make_pair(i, height=384, width=1248, max_depth=40.4, min_depth=2.0):
    g = torch.Generator().manual_seed(1000 + i)
    img = torch.rand(1, 3, height, width, generator=g)
    img = F.avg_pool2d(F.pad(img, (2, 2, 2, 2), mode='replicate'), 5, 1)
    d = int(torch.randint(8, 65, (1,), generator=g).item())
    d = min(d, max(width // 8, 1))
    imgR = torch.roll(img, -d, 3)
    mean = torch.tensor(MEAN).view(1, 3, 1, 1)
    std = torch.tensor(STD).view(1, 3, 1, 1)
    depth = min_depth + (max_depth - min_depth) * torch.rand(1, height, width, generator=g)
    depth = depth * (torch.rand(1, height, width, generator=g) >= 0.3)
    return {'imgL': (img - mean) / std, 'imgR': (imgR - mean) / std, 'disp_L': depth}


def make_calib(n=1, scale=1.0, cu=KITTI_CU, cv=KITTI_CV):
    """(calibs_fu [N], calibs_baseline [N], calibs_Proj [N,3,4], calibs_Proj_R [N,3,4]) as CPU
    float64 tensors, the types the reference builds at attack/DSGN/pgd_attack.py:261-266."""
    f = KITTI_F * scale
    P = torch.tensor([[f, 0, cu, 0], [0, f, cv, 0], [0, 0, 1, 0]], dtype=torch.float64)
    PR = P.clone()
    PR[0, 3] = -f * KITTI_B
    fu = torch.full((n,), f, dtype=torch.float64)
    base = torch.abs((P[0, 3] - PR[0, 3]) / P[0, 0]).repeat(n)
    return fu, base, P.repeat(n, 1, 1), PR.repeat(n, 1, 1)


def make_labels(cfg, n, seed, device='cpu'):
    g = torch.Generator().manual_seed(seed)
    zz = int(round((cfg.z_range[1] - cfg.z_range[0]) / cfg.voxel))
    xx = int(round((cfg.x_range[1] - cfg.x_range[0]) / cfg.voxel))
    cls = (torch.rand(n, cfg.num_anchors, zz, xx, generator=g) < 0.01).float()
    reg = torch.randn(n, cfg.num_anchors * cfg.reg_dim, zz, xx, generator=g)
    ctr = torch.rand(n, cfg.num_anchors, zz, xx, generator=g)
    return {'cls': cls.to(device), 'reg': reg.to(device), 'ctr': ctr.to(device)}


def make_targets(n_boxes, seed):
    """Seeded synthetic ground truth of one image in the reference's container layout: ``bbox`` [K,4]
    (x1,y1,x2,y2) and ``box3d`` [K,7] (h,w,l,x,y,z,theta) -- the tensors behind ``targets[0].bbox.data`` /
    ``targets[0].box3d.data`` that attack/DSGN/patch_attack.py:336-354 overwrites."""
    g = torch.Generator().manual_seed(seed)
    x1 = torch.rand(n_boxes, generator=g) * 1100
    y1 = 150 + torch.rand(n_boxes, generator=g) * 100
    bbox = torch.stack([x1, y1, x1 + 40 + torch.rand(n_boxes, generator=g) * 80,
                        y1 + 30 + torch.rand(n_boxes, generator=g) * 60], 1)
    box3d = torch.stack([1.4 + 0.4 * torch.rand(n_boxes, generator=g), 1.5 + 0.3 * torch.rand(n_boxes, generator=g),
                         3.4 + 1.0 * torch.rand(n_boxes, generator=g), -20 + 40 * torch.rand(n_boxes, generator=g),
                         1.5 + 0.5 * torch.rand(n_boxes, generator=g), 5 + 33 * torch.rand(n_boxes, generator=g),
                         -3.14 + 6.28 * torch.rand(n_boxes, generator=g)], 1)
    return bbox, box3d


def labels_from_box3d(cfg, box3d, device='cpu'):
    """BEV label maps {'cls','reg','ctr'} for the stand-in detection loss from ground-truth 3-D boxes
    (h,w,l,x,y,z,theta; all-zero rows are ignored, as the reference leaves them after zeroing the real ground
    truth).  The upstream RPN3DLoss target assignment is unavailable (un-vendored DSGN); this is the fixed
    stand-in: the 4 anchors are yaw bins of 90 degrees; cells whose centre lies inside a box's BEV footprint
    are positives of the bin nearest to theta; regression targets are (dx, dy, dz, log h, log w, log l,
    dtheta) relative to the cell / the bin; centerness is the usual sqrt(min/max . min/max) of the distances
    to the footprint's edges."""
    import math
    zz = int(round((cfg.z_range[1] - cfg.z_range[0]) / cfg.voxel))
    xx = int(round((cfg.x_range[1] - cfg.x_range[0]) / cfg.voxel))
    zc = cfg.z_range[0] + (torch.arange(zz, dtype=torch.float32) + 0.5) * cfg.voxel
    xc = cfg.x_range[0] + (torch.arange(xx, dtype=torch.float32) + 0.5) * cfg.voxel
    Z, X = torch.meshgrid(zc, xc, indexing='ij')
    a, r = cfg.num_anchors, cfg.reg_dim
    cls = torch.zeros(1, a, zz, xx)
    reg = torch.zeros(1, a * r, zz, xx)
    ctr = torch.zeros(1, a, zz, xx)
    for b in torch.as_tensor(box3d, dtype=torch.float32).reshape(-1, 7):
        h, w, l, x, y, z, th = (float(v) for v in b)
        if l <= 0 or w <= 0:
            continue
        dx, dz = X - x, Z - z
        u = math.cos(th) * dx - math.sin(th) * dz            # along the box length
        v = math.sin(th) * dx + math.cos(th) * dz            # along the box width
        inside = (u.abs() <= l / 2) & (v.abs() <= w / 2)
        k = int(round(th / (math.pi / 2))) % a
        cls[0, k][inside] = 1.0
        vals = (x - X, torch.full_like(X, y), z - Z, torch.full_like(X, math.log(h)), torch.full_like(X, math.log(w)),
                torch.full_like(X, math.log(l)), torch.full_like(X, th - k * math.pi / 2))
        for j, t in enumerate(vals[:r]):
            reg[0, k * r + j][inside] = t[inside]
        lu, ru, lv, rv = l / 2 + u, l / 2 - u, w / 2 + v, w / 2 - v
        c = torch.sqrt((torch.minimum(lu, ru) / torch.maximum(lu, ru).clamp_min(1e-6)).clamp_min(0) *
                       (torch.minimum(lv, rv) / torch.maximum(lv, rv).clamp_min(1e-6)).clamp_min(0))
        ctr[0, k][inside] = c[inside]
    return {'cls': cls.to(device), 'reg': reg.to(device), 'ctr': ctr.to(device)}
