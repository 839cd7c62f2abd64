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
