"""Size-independent properties of the volume kernels at BASELINE's FULL sizes (KITTI 384x1248:
PSV 64x48x96x312, voxel grid 192x20x304), where the CPU oracle is too slow to be the checker:
adjointness of every forward / data-gradient pair (<A x, y> == <x, A^T y>), GroupNorm invariants,
closed-form answers of the depth head, linearity, and the empty-input edge cases of the C ABI."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(built_lib):
    from eval_driving_safety_b200 import ops
    return ops


def _cl3(shape, g, scale=1.0):
    """random channels-last volume with logical NCDHW ``shape``"""
    n, c, d, h, w = shape
    return (torch.randn(n, d, h, w, c, generator=g) * scale).cuda().permute(0, 4, 1, 2, 3)


def _dot(a, b):
    return (a.double() * b.double()).sum().item()


# TF32 rounds both operands of both directions, so <conv(x), y> and <x, dgrad(y)> agree to TF32
# rounding noise averaged over ~1e8 products, not to fp32 epsilon.
CONV_CASES = [
    # (Cin, Cout, stride, transposed, input spatial)           -- the layer shapes of the two hourglasses
    (64, 64, 1, False, (48, 96, 312)),       # dres0/dres1/classif1 (PSV, full resolution)
    (96, 64, 1, False, (192, 20, 304)),      # rpn3d_conv on the voxel grid (role-swapped tiles)
    (64, 128, 2, False, (48, 96, 312)),      # hourglass conv1
    (128, 128, 2, False, (24, 48, 156)),     # hourglass conv3
    (128, 128, 2, True, (12, 24, 78)),       # hourglass conv5 (transposed)
    (128, 64, 2, True, (96, 10, 152)),       # hourglass conv6 on the voxel grid (transposed)
]


@pytest.mark.parametrize("cin,cout,stride,transposed,sp", CONV_CASES)
def test_conv3d_adjoint_at_kitti_sizes(ops, cin, cout, stride, transposed, sp):
    g = torch.Generator().manual_seed(cin + cout + stride + sp[0])
    x = _cl3((1, cin) + sp, g).requires_grad_(True)
    wshape = (cin, cout, 3, 3, 3) if transposed else (cout, cin, 3, 3, 3)
    w = (torch.randn(wshape, generator=g) * (27 * cin) ** -0.5).cuda()
    y = ops.conv3d(x, w, stride=stride, transposed=transposed)
    exp = tuple(s * stride for s in sp) if transposed else tuple((s - 1) // stride + 1 for s in sp)
    assert y.shape == (1, cout) + exp
    gy = _cl3(tuple(y.shape), g)
    (gx,) = torch.autograd.grad(y, x, gy)
    lhs, rhs = _dot(y, gy), _dot(x, gx)
    scale = (_dot(y, y) * _dot(gy, gy)) ** 0.5
    assert abs(lhs - rhs) < 2e-5 * scale, (lhs, rhs, scale)
    # linearity of the forward (TF32 rounding commutes with a power-of-two scale): bit-exact
    assert torch.equal(ops.conv3d(2 * x.detach(), w, stride=stride, transposed=transposed), 2 * y.detach())


STAT_CASES = [
    # (N, Cin, Cout, stride, transposed, input spatial, statistics expected)
    (1, 64, 64, 1, False, (48, 96, 312), True),       # N-stacked stride-1 kernel, KITTI PSV size
    (1, 96, 64, 1, False, (192, 20, 304), True),      # role-swapped tiles
    (1, 128, 128, 1, False, (12, 24, 78), True),      # two N tiles
    (1, 32, 32, 1, False, (5, 7, 9), True),           # ragged: partial tiles / planes must not be counted
    (1, 128, 64, 2, True, (24, 48, 156), True),       # class-stacked transposed kernel
    (1, 128, 128, 2, True, (6, 5, 9), True),          # transposed, two N tiles, ragged
    (1, 64, 128, 2, False, (12, 24, 40), False),      # stride-2 conv: generic kernel, GroupNorm's own pass
    (2, 64, 64, 1, False, (6, 16, 16), False),        # N > 1: a CTA row would mix samples
]


@pytest.mark.parametrize("n,cin,cout,stride,transposed,sp,expect", STAT_CASES)
def test_conv_epilogue_groupnorm_statistics(ops, n, cin, cout, stride, transposed, sp, expect):
    """conv3d_with_stats: the epilogue's per-CTA partial sums reproduce the statistics of the written
    output (fp64 reference), GroupNorm fed with them equals GroupNorm with its own pass, forward and
    backward, and the whole thing is bitwise reproducible."""
    g = torch.Generator().manual_seed(cin * 3 + cout + sp[0])
    x = _cl3((n, cin) + sp, g).requires_grad_(True)
    wshape = (cin, cout, 3, 3, 3) if transposed else (cout, cin, 3, 3, 3)
    w = (torch.randn(wshape, generator=g) * (27 * cin) ** -0.5).cuda()
    y, part = ops.conv3d_with_stats(x, w, stride=stride, transposed=transposed)
    y0 = ops.conv3d(x, w, stride=stride, transposed=transposed)
    assert torch.equal(y, y0)                                        # same kernel, same output
    assert (part is not None) == expect
    gamma = (torch.rand(cout, generator=g) + 0.5).cuda()
    beta = torch.randn(cout, generator=g).cuda()
    ref = ops.groupnorm_act(y0, gamma, beta, 32, 1e-5, relu=True)
    if part is None:
        return
    assert part.shape[1:] == (2, cout)
    yd = y.detach().double()
    s_ref, q_ref = yd.sum((0, 2, 3, 4)), (yd * yd).sum((0, 2, 3, 4))
    s_got, q_got = part[:, 0].double().sum(0), part[:, 1].double().sum(0)
    scale = q_ref.sqrt() * yd[0, 0].numel() ** 0.5 + 1e-6
    assert ((s_got - s_ref).abs() / scale).max().item() < 1e-5
    assert ((q_got - q_ref).abs() / q_ref).max().item() < 1e-5
    out = ops.groupnorm_act(y, gamma, beta, 32, 1e-5, relu=True, partial=part)
    assert (out - ref).abs().max().item() < 2e-5
    gy = _cl3(tuple(out.shape), g)
    (gx,) = torch.autograd.grad(out, x, gy)
    (gx_ref,) = torch.autograd.grad(ref, x, gy)
    assert (gx - gx_ref).abs().max().item() < 1e-4 * gx_ref.abs().max().item() + 1e-6
    y2, part2 = ops.conv3d_with_stats(x, w, stride=stride, transposed=transposed)
    assert torch.equal(part, part2)                                  # fixed summation order


@pytest.mark.parametrize("cin,cout,stride,transposed,sp", [
    (64, 64, 1, False, (12, 24, 40)),        # dgrad = stride-1 kernel, fused add
    (64, 96, 1, False, (5, 7, 9)),           # dgrad has Nt = 32 from 64 channels; ragged tiles
    (64, 128, 2, False, (12, 24, 40)),       # dgrad = class-stacked transposed kernel, fused add
    (128, 64, 2, True, (6, 12, 20)),         # dgrad = stride-2 conv (generic kernel): separate add
])
def test_conv3d_fork_adds_the_other_gradient_in_the_epilogue(ops, cin, cout, stride, transposed, sp):
    """x feeds a conv and a second consumer: with conv3d_fork the second consumer's gradient is added by
    the data-gradient kernel (out = acc + addend), bit-identical to autograd's separate accumulation."""
    g = torch.Generator().manual_seed(cin + cout + sp[2])
    x = _cl3((1, cin) + sp, g).requires_grad_(True)
    wshape = (cin, cout, 3, 3, 3) if transposed else (cout, cin, 3, 3, 3)
    w = (torch.randn(wshape, generator=g) * (27 * cin) ** -0.5).cuda()
    y0 = ops.conv3d(x, w, stride=stride, transposed=transposed)
    gy, gr = _cl3(tuple(y0.shape), g), _cl3(tuple(x.shape), g)
    (ref,) = torch.autograd.grad((y0 * gy).sum() + (x * gr).sum(), x)
    y, part, x2 = ops.conv3d_fork(x, w, stride=stride, transposed=transposed)
    assert torch.equal(y, y0) and x2.data_ptr() == x.data_ptr()
    (got,) = torch.autograd.grad((y * gy).sum() + (x2 * gr).sum(), x)
    assert torch.equal(got, ref)
    # either output alone
    (g1,) = torch.autograd.grad((ops.conv3d_fork(x, w, stride=stride, transposed=transposed)[2] * gr).sum(), x)
    assert torch.equal(g1, gr)
    (g2,) = torch.autograd.grad((ops.conv3d_fork(x, w, stride=stride, transposed=transposed)[0] * gy).sum(), x)
    (g2r,) = torch.autograd.grad((ops.conv3d(x, w, stride=stride, transposed=transposed) * gy).sum(), x)
    assert torch.equal(g2, g2r)


@pytest.mark.parametrize("relu,with_res,second_consumer,sp,c", [
    (True, False, False, (12, 24, 40), 64),     # norm+ReLU -> conv: mask recomputed in the conv epilogue
    (False, True, False, (12, 24, 40), 64),     # norm + residual, no activation
    (True, False, True, (5, 7, 9), 64),         # the norm's output also feeds a second consumer through the fork
    (True, False, False, (6, 10, 12), 128),     # two N tiles
    (True, True, False, (6, 10, 12), 64),       # ReLU + residual needs the saved output: not fused, still right
    (True, False, "plain", (6, 10, 12), 64),    # second consumer WITHOUT the fork: autograd's own (in-place) accumulation
])
def test_groupnorm_backward_sums_from_the_consuming_convs_dgrad(ops, relu, with_res, second_consumer, sp, c):
    """conv_a -> GroupNorm -> conv_b: conv_b's data-gradient launch adds up the norm's backward sums
    (sum gz*x, sum gz) in its epilogue; the input gradient equals the unfused path."""
    g = torch.Generator().manual_seed(sp[0] * 7 + c)
    x = _cl3((1, c) + sp, g).requires_grad_(True)
    wa = (torch.randn(c, c, 3, 3, 3, generator=g) * (27 * c) ** -0.5).cuda()
    wb = (torch.randn(c, c, 3, 3, 3, generator=g) * (27 * c) ** -0.5).cuda()
    gamma = (torch.rand(c, generator=g) + 0.5).cuda()
    beta = (torch.randn(c, generator=g) * 0.3).cuda()
    res = _cl3((1, c) + sp, g) if with_res else None
    gy, g2 = _cl3((1, c) + sp, g), _cl3((1, c) + sp, g)

    def run(fuse):
        old = ops.FUSE_GN_BWD
        ops.FUSE_GN_BWD = fuse
        try:
            n0 = ops.LAUNCH_COUNT
            ya = ops.conv3d(x, wa)
            h = ops.groupnorm_act(ya, gamma, beta, 32, 1e-5, relu=relu, res=res)
            if second_consumer == "plain":
                loss = (ops.conv3d(h, wb) * gy).sum() + (h * g2).sum()
            elif second_consumer:
                yb, _, h2 = ops.conv3d_fork(h, wb)
                loss = (yb * gy).sum() + (h2 * g2).sum()
            else:
                loss = (ops.conv3d(h, wb) * gy).sum()
            (gx,) = torch.autograd.grad(loss, x)
            return gx, ops.LAUNCH_COUNT - n0
        finally:
            ops.FUSE_GN_BWD = old

    ref, n_ref = run(False)
    got, n_got = run(True)
    fusable = not (relu and with_res) and second_consumer != "plain"
    assert n_got == n_ref - (1 if fusable else 0)          # the norm's backward statistics launch is gone
    assert (got - ref).abs().max().item() < 2e-5 * ref.abs().max().item() + 1e-7
    got2, _ = run(True)
    assert torch.equal(got, got2)                          # reproducible


def test_conv3d_c1_adjoint_at_kitti_size(ops):
    g = torch.Generator().manual_seed(5)
    x = _cl3((1, 64, 48, 96, 312), g).requires_grad_(True)
    w = (torch.randn(1, 64, 3, 3, 3, generator=g) * 0.02).cuda()
    y = ops.conv3d_c1(x, w)
    assert y.shape == (1, 1, 48, 96, 312)
    gy = torch.randn(y.shape, generator=g).cuda()
    (gx,) = torch.autograd.grad(y, x, gy)
    lhs, rhs = _dot(y, gy), _dot(x, gx)
    assert abs(lhs - rhs) < 1e-6 * (_dot(y, y) * _dot(gy, gy)) ** 0.5      # fp32 SIMT both ways


def test_lift_adjoint_at_kitti_size(ops):
    """grid_sample 3-D (PSV -> voxels) and 2-D (image features -> voxels) with the real KITTI frustum grid:
    forward gather vs the CSR-plan backward."""
    from eval_driving_safety_b200 import dsgn, synthetic
    cfg = dsgn.default_cfg()
    _, _, P, _ = synthetic.make_calib(1)
    g = torch.Generator().manual_seed(7)
    psv = _cl3((1, 64, 48, 96, 312), g).requires_grad_(True)
    grid3 = dsgn.lifting_grid(cfg, P, (96, 312)).cuda().contiguous()
    plan3 = ops.GridPlan(grid3, (48, 96, 312), True)
    out = ops.grid_sample(psv, grid3, True, plan3)
    assert out.shape == (1, 64, 192, 20, 304)
    gy = _cl3(tuple(out.shape), g)
    (gp,) = torch.autograd.grad(out, psv, gy)
    lhs, rhs = _dot(out, gy), _dot(psv, gp)
    assert abs(lhs - rhs) < 1e-6 * (_dot(out, out) * _dot(gy, gy)) ** 0.5
    # a constant volume samples to the constant wherever all eight corners are inside, 0 where none is
    ones = torch.ones_like(psv.detach())
    o1 = ops.grid_sample(ones, grid3, True, plan3)
    assert o1.max().item() <= 1.0 + 1e-5 and o1.min().item() >= 0.0
    inside = (grid3.abs() < 0.98).all(-1)[0]                                # [Z, Y, X]
    assert inside.any() and (o1[0, :, inside] - 1.0).abs().max().item() < 1e-5
    # deterministic backward at full size
    (gp2,) = torch.autograd.grad(ops.grid_sample(psv, grid3, True, plan3), psv, gy)
    assert torch.equal(gp, gp2)


def test_cost_volume_adjoint_at_kitti_size(ops):
    from eval_driving_safety_b200 import dsgn, synthetic
    cfg = dsgn.default_cfg()
    fu, b, _, _ = synthetic.make_calib(1)
    shifts = dsgn.plane_shifts(cfg, fu, b).cuda()
    g = torch.Generator().manual_seed(9)
    l = torch.randn(1, 32, 96, 312, generator=g).cuda().requires_grad_(True)
    r = torch.randn(1, 32, 96, 312, generator=g).cuda().requires_grad_(True)
    c = ops.build_cost_volume(l, r, shifts, channels_last=True)
    gy = _cl3(tuple(c.shape), g)
    gl, gr = torch.autograd.grad(c, [l, r], gy)
    lhs, rhs = _dot(c, gy), _dot(l, gl) + _dot(r, gr)
    assert abs(lhs - rhs) < 1e-6 * (_dot(c, c) * _dot(gy, gy)) ** 0.5


@pytest.mark.parametrize("shape", [(1, 64, 48, 96, 312), (1, 128, 96, 10, 152), (2, 32, 192, 624)])
def test_groupnorm_invariants_at_kitti_sizes(ops, shape):
    g = torch.Generator().manual_seed(shape[1])
    n, c = shape[:2]
    if len(shape) == 5:
        x = _cl3(shape, g)
    else:
        x = torch.randn(n, shape[2], shape[3], c, generator=g).cuda().permute(0, 3, 1, 2)
    x = x * 3.0 + 1.5
    gamma, beta = torch.ones(c).cuda(), torch.zeros(c).cuda()
    y = ops.groupnorm_act(x, gamma, beta, 32, 1e-5, relu=False)
    yg = y.reshape(n, 32, c // 32, -1).double()
    assert yg.mean((2, 3)).abs().max().item() < 1e-4                       # zero mean per (sample, group)
    assert (yg.var((2, 3), unbiased=False) - 1.0).abs().max().item() < 1e-3  # unit variance
    # invariance to an affine change of the input (up to rounding): GN(a x + b) == GN(x)
    y2 = ops.groupnorm_act(x * 4.0 + 8.0, gamma, beta, 32, 1e-5, relu=False)
    assert (y - y2).abs().max().item() < 2e-4
    # backward: the data gradient of GN is orthogonal to constants and to x_hat within each group
    xr = x.detach().clone().requires_grad_(True)
    yr = ops.groupnorm_act(xr, gamma, beta, 32, 1e-5, relu=False)
    gy = torch.randn(yr.shape, generator=torch.Generator().manual_seed(1)).cuda().contiguous(
        memory_format=torch.channels_last_3d if len(shape) == 5 else torch.channels_last)
    (gx,) = torch.autograd.grad(yr, xr, gy)
    gxg = gx.reshape(n, 32, c // 32, -1).double()
    num = gxg.abs().mean().item()
    assert gxg.mean((2, 3)).abs().max().item() < 1e-4 * max(num, 1e-3) + 1e-6
    assert (gxg * yg).mean((2, 3)).abs().max().item() < 1e-4 * max(num, 1e-3) + 1e-6
    # ReLU + bitwise reproducibility at full size
    a = ops.groupnorm_act(x, gamma, beta, 32, 1e-5, relu=True)
    assert torch.equal(a, torch.relu(y)) and torch.equal(a, ops.groupnorm_act(x, gamma, beta, 32, 1e-5, relu=True))


def test_depth_head_closed_forms_at_kitti_size(ops):
    from eval_driving_safety_b200 import dsgn
    cfg = dsgn.default_cfg()
    j, h, w = cfg.maxdisp, 384, 1248
    z = cfg.min_depth + (torch.arange(j, dtype=torch.float64) + 0.5) * cfg.depth_interval
    # constant cost -> uniform softmax -> depth = mean of the plane depths, gradient wrt cost sums to 0
    cost = torch.full((1, 1, 48, 96, 312), 0.7).cuda().requires_grad_(True)
    d = ops.depth_head(cost, (j, h, w), cfg.min_depth, cfg.depth_interval)
    assert d.shape == (1, h, w)
    assert (d - z.mean().item()).abs().max().item() < 1e-4
    (gc,) = torch.autograd.grad(d, cost, torch.ones_like(d))
    assert abs(gc.double().sum().item()) < 1e-2 * gc.double().abs().sum().item() + 1e-6
    # softmax is shift invariant
    g = torch.Generator().manual_seed(3)
    c2 = torch.randn(1, 1, 48, 96, 312, generator=g).cuda()
    d0 = ops.depth_head(c2, (j, h, w), cfg.min_depth, cfg.depth_interval)
    d1 = ops.depth_head(c2 + 5.0, (j, h, w), cfg.min_depth, cfg.depth_interval)
    assert (d0 - d1).abs().max().item() < 1e-3
    assert d0.min().item() >= z[0].item() - 1e-3 and d0.max().item() <= z[-1].item() + 1e-3
    # a cost that is +inf-like on one coarse plane puts the depth inside that plane's span
    c3 = torch.zeros(1, 1, 48, 96, 312)
    c3[:, :, 20] = 80.0
    d3 = ops.depth_head(c3.cuda(), (j, h, w), cfg.min_depth, cfg.depth_interval)
    lo, hi = z[4 * 20 - 2].item(), z[4 * 20 + 5].item()
    assert d3.min().item() >= lo and d3.max().item() <= hi


def test_pgd_update_properties_at_kitti_size(ops):
    """Idempotence / box properties of the pixel update on a full 384x1248 pair (bit-level)."""
    from eval_driving_safety_b200 import attack
    from oracle import attack_ref as A
    g = torch.Generator().manual_seed(11)
    clean = torch.rand(1, 3, 384, 1248, generator=g)
    mean, std = torch.tensor(A.IMAGENET_MEAN).view(1, 3, 1, 1), torch.tensor(A.IMAGENET_STD).view(1, 3, 1, 1)
    x = ((clean - mean) / std).cuda()
    grad = torch.randn(1, 3, 384, 1248, generator=g).cuda()
    eps, alpha = 0.03, 0.0075
    x1 = attack.pgd_step(x, grad, clean.cuda(), alpha, eps)
    x01 = x1.cpu() * std + mean
    assert (x01 - clean).abs().max().item() <= eps + 1e-6                    # inside the eps-ball
    assert x01.min().item() >= -1e-6 and x01.max().item() <= 1 + 1e-6        # inside the pixel box
    # zero gradient: sign(0) = 0 -> the step only re-projects (a fixed point after one application)
    z = torch.zeros_like(grad)
    x2 = attack.pgd_step(x1, z, clean.cuda(), alpha, eps)
    x3 = attack.pgd_step(x2, z, clean.cuda(), alpha, eps)
    assert torch.equal(x2, x3)
    # many steps with a constant gradient saturate at clean +- eps (clipped to the box)
    xs = x
    for _ in range(6):
        xs = attack.pgd_step(xs, grad, clean.cuda(), alpha, eps)
    want = (clean + eps * torch.sign(grad.cpu())).clamp(0, 1)
    assert ((xs.cpu() * std + mean) - want).abs().max().item() < 1e-5


def test_cabi_empty_inputs_are_noops(built_lib):
    """N == 0 / zero voxels: every entry point returns 0 without launching (reference scripts can hit
    this with an empty shard or an image without RoIs)."""
    from eval_driving_safety_b200 import _lib
    lib = _lib.load()
    buf = torch.zeros(1024, device="cuda")
    p = ctypes.c_void_p(buf.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.b2_cost_volume_fwd(p, p, p, p, 0, 32, 48, 96, 312, 1, st) == 0
    assert lib.b2_grid_sample3d_fwd(p, p, p, 0, 64, 48, 96, 312, 0, 64, 0, 1, st) == 0
    assert lib.b2_groupnorm_fwd(p, None, p, p, p, p, 0, 64, 100, 32, 1e-5, 0, p, st) == 0
    assert lib.b2_roi_align_fwd(p, None, p, 0, 256, 38, 125, 7, 0.0625, st) == 0
    torch.cuda.synchronize()
    assert buf.abs().max().item() == 0.0
