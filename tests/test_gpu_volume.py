"""GPU parity of the volume kernels through the C ABI against the oracle's stock
torch ops on CPU: cost volume (subsystem 1), grid_sample lifting (2), conv3d /
deconv3d / GroupNorm (3), RoIAlign (config 5)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import max_err, rel_err
from oracle import attack_ref as A
from oracle import dsgn_ref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(built_lib):
    from eval_driving_safety_b200 import ops
    return ops


# ---------------------------------------------------------------- cost volume
@pytest.mark.parametrize("channels_last", [True, False])
@pytest.mark.parametrize("dims", [(1, 4, 6, 8, 16), (2, 32, 12, 24, 78), (1, 8, 5, 3, 7)])
def test_cost_volume_fwd_bwd(ops, dims, channels_last):
    n, c, d, h, w = dims
    g = torch.Generator().manual_seed(sum(dims))
    l = torch.randn(n, c, h, w, generator=g, requires_grad=True)
    r = torch.randn(n, c, h, w, generator=g, requires_grad=True)
    shifts = torch.rand(n, d, generator=g) * (w + 2)
    shifts[:, 0] = 0.0
    shifts[:, 1] = 3.0                                   # integer shift
    ref = R.build_cost_volume(l, r, shifts)
    gy = torch.randn(ref.shape, generator=g)
    gl_ref, gr_ref = torch.autograd.grad(ref, [l, r], gy)
    lc, rc = l.detach().cuda().requires_grad_(True), r.detach().cuda().requires_grad_(True)
    out = ops.build_cost_volume(lc, rc, shifts.cuda(), channels_last)
    assert out.shape == ref.shape
    assert torch.equal(out.cpu(), ref)                   # fp32 copy/lerp with the oracle's op order: bit-exact
    gl, gr = torch.autograd.grad(out, [lc, rc], gy.cuda())
    assert max_err(gl.cpu(), gl_ref) < 1e-5 and max_err(gr.cpu(), gr_ref) < 1e-5   # sum over planes, order differs
    gl2, gr2 = torch.autograd.grad(ops.build_cost_volume(lc, rc, shifts.cuda(), channels_last), [lc, rc], gy.cuda())
    assert torch.equal(gl, gl2) and torch.equal(gr, gr2)                            # deterministic


def test_cost_volume_kitti_size_properties(ops):
    """Full size (368 MB): linearity + PSMNet integer-shift identity, size-independent checks."""
    from eval_driving_safety_b200 import dsgn, synthetic
    cfg = dsgn.default_cfg()
    fu, b, _, _ = synthetic.make_calib(1)
    shifts = dsgn.plane_shifts(cfg, fu, b).cuda()
    g = torch.Generator().manual_seed(1)
    l, r = torch.randn(1, 32, 96, 312, generator=g).cuda(), torch.randn(1, 32, 96, 312, generator=g).cuda()
    c1 = ops.build_cost_volume(l, r, shifts)
    assert c1.shape == (1, 64, 48, 96, 312)
    c2 = ops.build_cost_volume(2 * l, 2 * r, shifts)
    assert torch.equal(c2, 2 * c1)
    ish = torch.floor(shifts)
    ci = ops.build_cost_volume(l, r, ish)
    d = 17
    s = int(ish[0, d])
    assert torch.equal(ci[0, :32, d, :, s:], l[0, :, :, s:]) and torch.equal(ci[0, 32:, d, :, s:], r[0, :, :, :312 - s])
    assert ci[0, :, d, :, :s].abs().max() == 0


# ---------------------------------------------------------------- grid sample
@pytest.mark.parametrize("align", [True, False])
def test_grid_sample3d_fwd_bwd(ops, align):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 64, 6, 9, 11, generator=g, requires_grad=True)
    grid = (torch.rand(2, 7, 5, 13, 3, generator=g) * 2.6 - 1.3)        # some samples out of range
    grid[0, 0, 0, 0] = torch.tensor([-1.0, 1.0, 0.0])
    ref = F.grid_sample(x, grid, mode='bilinear', padding_mode='zeros', align_corners=align)
    gy = torch.randn(ref.shape, generator=g)
    (gx_ref,) = torch.autograd.grad(ref, x, gy)
    xc = x.detach().cuda().requires_grad_(True)
    out = ops.grid_sample(xc, grid.cuda(), align)
    ok = ~torch.isnan(ref)
    assert max_err(out.cpu()[ok], ref[ok]) < 1e-5
    (gx,) = torch.autograd.grad(out, xc, gy.cuda())
    assert max_err(gx.cpu(), torch.nan_to_num(gx_ref)) < 1e-4
    (gx2,) = torch.autograd.grad(ops.grid_sample(xc, grid.cuda(), align), xc, gy.cuda())
    assert torch.equal(gx, gx2)                                          # deterministic backward


def test_grid_sample2d_and_lift(ops):
    g = torch.Generator().manual_seed(4)
    img = torch.randn(1, 32, 10, 14, generator=g, requires_grad=True)
    psv = torch.randn(1, 64, 5, 10, 14, generator=g, requires_grad=True)
    grid3 = torch.rand(1, 6, 4, 9, 3, generator=g) * 2.4 - 1.2
    grid2 = grid3[..., :2].reshape(1, 24, 9, 2)
    ref2 = F.grid_sample(img, grid2, mode='bilinear', padding_mode='zeros', align_corners=True)
    out2 = ops.grid_sample(img.detach().cuda(), grid2.cuda(), True)
    assert max_err(out2.cpu(), ref2) < 1e-5
    ref = torch.cat([F.grid_sample(psv, grid3, mode='bilinear', padding_mode='zeros', align_corners=True),
                     ref2.view(1, 32, 6, 4, 9)], 1)
    gy = torch.randn(ref.shape, generator=g)
    gp_ref, gi_ref = torch.autograd.grad(ref, [psv, img], gy)
    pc, ic = psv.detach().cuda().requires_grad_(True), img.detach().cuda().requires_grad_(True)
    out = ops.lift(pc, ic, grid3.cuda())
    assert out.shape == ref.shape and max_err(out.cpu(), ref) < 1e-5
    gp, gi = torch.autograd.grad(out, [pc, ic], gy.cuda())
    assert max_err(gp.cpu(), gp_ref) < 1e-4 and max_err(gi.cpu(), gi_ref) < 1e-4


# ---------------------------------------------------------------- conv3d
CONV_CASES = [  # (cin, cout, stride, transposed, spatial)
    (64, 64, 1, False, (4, 6, 8)), (96, 64, 1, False, (4, 4, 8)), (128, 128, 1, False, (2, 4, 6)),
    (64, 128, 2, False, (4, 6, 8)), (128, 128, 2, False, (4, 4, 4)),
    (128, 128, 2, True, (2, 3, 4)), (128, 64, 2, True, (2, 2, 6)), (8, 12, 1, False, (3, 5, 7))]


def _conv_ref(x, w, stride, transposed):
    if transposed:
        return F.conv_transpose3d(x, w, None, 2, 1, output_padding=1)
    return F.conv3d(x, w, None, stride, 1)


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv3d_simt_fwd_dgrad(ops, case):
    cin, cout, stride, transposed, sp = case
    g = torch.Generator().manual_seed(cin + cout + stride)
    x = torch.randn(2, cin, *sp, generator=g, requires_grad=True)
    w = torch.randn((cin, cout, 3, 3, 3) if transposed else (cout, cin, 3, 3, 3), generator=g) * 0.05
    ref = _conv_ref(x, w, stride, transposed)
    gy = torch.randn(ref.shape, generator=g)
    (gx_ref,) = torch.autograd.grad(ref, x, gy)
    xc = x.detach().cuda().requires_grad_(True)
    out = ops.conv3d(xc, w.cuda(), stride, transposed, impl=1)
    assert out.shape == ref.shape
    assert rel_err(out.cpu(), ref) < 1e-4        # fp32 both sides, K up to 3456, different summation order
    (gx,) = torch.autograd.grad(out, xc, gy.cuda())
    assert rel_err(gx.cpu(), gx_ref) < 1e-4


def test_conv3d_c1_and_groupnorm(ops):
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 64, 4, 6, 10, generator=g, requires_grad=True)
    w = torch.randn(1, 64, 3, 3, 3, generator=g) * 0.1
    ref = F.conv3d(x, w, None, 1, 1)
    gy = torch.randn(ref.shape, generator=g)
    (gx_ref,) = torch.autograd.grad(ref, x, gy)
    xc = x.detach().cuda().requires_grad_(True)
    out = ops.conv3d_c1(xc, w.cuda())
    assert rel_err(out.cpu(), ref) < 1e-5
    (gx,) = torch.autograd.grad(out, xc, gy.cuda())
    assert rel_err(gx.cpu(), gx_ref) < 1e-5


@pytest.mark.parametrize("c,relu,has_res", [(64, True, False), (64, False, True), (128, True, True), (64, False, False)])
def test_groupnorm_act(ops, c, relu, has_res):
    g = torch.Generator().manual_seed(c + relu)
    x = (torch.randn(2, c, 3, 5, 7, generator=g) * 2 + 0.5).requires_grad_(True)
    res = torch.randn(2, c, 3, 5, 7, generator=g).requires_grad_(True) if has_res else None
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    y = F.group_norm(x, 32, gamma, beta, 1e-5)
    if has_res:
        y = y + res
    if relu:
        y = F.relu(y)
    gy = torch.randn(y.shape, generator=g)
    grads_ref = torch.autograd.grad(y, [x] + ([res] if has_res else []), gy)
    xc = x.detach().cuda().requires_grad_(True)
    rc = res.detach().cuda().requires_grad_(True) if has_res else None
    out = ops.groupnorm_act(xc, gamma.cuda(), beta.cuda(), 32, 1e-5, relu, rc)
    assert max_err(out.cpu(), y) < 2e-5
    grads = torch.autograd.grad(out, [xc] + ([rc] if has_res else []), gy.cuda())
    for a, b in zip(grads, grads_ref):
        assert max_err(a.cpu(), b) < 5e-5
    grads2 = torch.autograd.grad(ops.groupnorm_act(xc, gamma.cuda(), beta.cuda(), 32, 1e-5, relu, rc),
                                 [xc] + ([rc] if has_res else []), gy.cuda())
    assert all(torch.equal(a, b) for a, b in zip(grads, grads2))


# ---------------------------------------------------------------- RoIAlign
@pytest.mark.parametrize("pooled", [7, 14])
def test_roi_align_vs_torchvision(ops, pooled):
    from torchvision.ops import roi_align
    g = torch.Generator().manual_seed(21)
    feat = torch.randn(1, 16, 38, 125, generator=g, requires_grad=True)
    x1, y1 = torch.rand(24, generator=g) * 1700, torch.rand(24, generator=g) * 500
    rois = torch.stack([torch.zeros(24), x1, y1, x1 + torch.rand(24, generator=g) * 400 + 1,
                        y1 + torch.rand(24, generator=g) * 200 + 1], 1)
    rois[0] = torch.tensor([0, -30.0, -20.0, 50.0, 40.0])            # partially outside
    rois[1] = torch.tensor([0, 1900.0, 560.0, 2100.0, 650.0])        # beyond the far edges
    scale = 38 / 600
    ref = roi_align(feat, rois, (pooled, pooled), scale, 0, False)
    gy = torch.randn(ref.shape, generator=g)
    (gf_ref,) = torch.autograd.grad(ref, feat, gy)
    fc = feat.detach().cuda().requires_grad_(True)
    out = ops.roi_align(fc, rois.cuda(), pooled, scale)
    assert max_err(out.cpu(), ref) < 1e-4
    (gf,) = torch.autograd.grad(out, fc, gy.cuda())
    assert max_err(gf.cpu(), gf_ref) < 1e-4
    (gf2,) = torch.autograd.grad(ops.roi_align(fc, rois.cuda(), pooled, scale), fc, gy.cuda())
    assert torch.equal(gf, gf2)


def test_pyramid_roi_dispatch_matches_oracle(ops):
    """FPN level dispatch of attack/Stereo-RCNN/stereo_rcnn.py:110-141 through our RoIAlign."""
    from eval_driving_safety_b200 import stereo_rcnn
    g = torch.Generator().manual_seed(22)
    sizes = [(150, 497), (75, 249), (38, 125), (19, 63)]
    feats = [torch.randn(1, 8, h, w, generator=g) for h, w in sizes]
    x1, y1 = torch.rand(40, generator=g) * 1500, torch.rand(40, generator=g) * 400
    w = torch.exp(torch.rand(40, generator=g) * 6.0) + 2
    h = torch.exp(torch.rand(40, generator=g) * 5.0) + 2
    rois = torch.stack([torch.zeros(40), x1, y1, x1 + w, y1 + h], 1)
    ref = A.pyramid_roi_feat(feats, rois, 600.0, 7)
    out = stereo_rcnn.pyramid_roi_feat([f.cuda() for f in feats], rois.cuda(), 600.0, 7)
    assert max_err(out.cpu(), ref) < 1e-4


# ---------------------------------------------------------------- conv3d, tensor-core path
TC_CASES = [c for c in CONV_CASES if c[0] % 32 == 0 and c[1] % 32 == 0] + [
    (64, 64, 1, False, (3, 20, 19)), (64, 96, 1, False, (2, 5, 9)), (64, 128, 2, False, (6, 36, 20)),
    (128, 64, 2, True, (3, 18, 10)),
    # D >> H: the kernels put the 16-row tile dimension along D instead of H ("role swap")
    (64, 64, 1, False, (20, 5, 9)), (96, 64, 1, False, (33, 3, 8)), (64, 128, 2, False, (32, 6, 8)),
    (128, 64, 2, True, (16, 3, 5))]


@pytest.fixture(params=[0, 1], ids=["single-cta", "cta-pair"])
def dc_pair(request, built_lib):
    """Both variants of the transposed-conv and of the stride-2 (one tap per stage) kernel: single-CTA and the tcgen05
    cta_group::2 CTA pair (default)."""
    from eval_driving_safety_b200 import _lib
    _lib.set_flag("conv_dc_pair", request.param)
    _lib.set_flag("conv_s2_pair", request.param)
    yield request.param
    _lib.set_flag("conv_dc_pair", None)
    _lib.set_flag("conv_s2_pair", None)


@pytest.mark.parametrize("case", TC_CASES)
def test_conv3d_tcgen05_fwd_dgrad(ops, case, dc_pair):
    """tcgen05/TMEM/TMA implicit GEMM against torch CPU fp32.  Tolerance: TF32 operands (10-bit
    mantissa, truncation) with fp32 accumulation -> ~8e-4 relative (measured); bound 3e-3."""
    cin, cout, stride, transposed, sp = case
    if not dc_pair and not (transposed or stride == 2):
        pytest.skip("the kernel variant only concerns DECONV launches (transposed forward, stride-2 data gradient)")
    g = torch.Generator().manual_seed(cin + cout + stride + sp[1])
    x = torch.randn(2, cin, *sp, generator=g, requires_grad=True)
    w = torch.randn((cin, cout, 3, 3, 3) if transposed else (cout, cin, 3, 3, 3), generator=g) * 0.05
    ref = _conv_ref(x, w, stride, transposed)
    gy = torch.randn(ref.shape, generator=g)
    (gx_ref,) = torch.autograd.grad(ref, x, gy)
    xc = x.detach().cuda().requires_grad_(True)
    out = ops.conv3d(xc, w.cuda(), stride, transposed, impl=0)
    assert out.shape == ref.shape and rel_err(out.cpu(), ref) < 3e-3
    (gx,) = torch.autograd.grad(out, xc, gy.cuda())
    assert rel_err(gx.cpu(), gx_ref) < 3e-3
    out2 = ops.conv3d(xc, w.cuda(), stride, transposed, impl=0)
    assert torch.equal(out, out2)


def test_conv3d_tcgen05_rejects_unsupported_widths(ops):
    x, w = torch.randn(1, 8, 4, 4, 4).cuda(), torch.randn(12, 8, 3, 3, 3).cuda()
    with pytest.raises(RuntimeError, match="impl=1"):
        ops.conv3d(x, w, 1, False, impl=0)
    w.requires_grad_(True)
    with pytest.raises(RuntimeError, match="frozen"):
        ops.conv3d(x, w, 1, False, impl=1)


def test_conv3d_kitti_size_tcgen05_vs_fp32_kernel(ops):
    """Full PSV size [1,64,48,96,312] (318 GFLOP): tensor-core path against the fp32 SIMT kernel,
    plus linearity (size-independent property)."""
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, 48, 96, 312, 64, generator=g).cuda().permute(0, 4, 1, 2, 3)
    w = (torch.randn(64, 64, 3, 3, 3, generator=g) * 0.03).cuda()
    a = ops.conv3d(x, w, 1, False, impl=0)
    b = ops.conv3d(x, w, 1, False, impl=1)
    assert rel_err(a, b) < 3e-3
    assert torch.equal(ops.conv3d(2 * x, w, 1, False, impl=0), 2 * a)     # power-of-two scaling is exact
    xs = torch.randn(1, 24, 48, 156, 128, generator=g).cuda().permute(0, 4, 1, 2, 3)
    wt = (torch.randn(128, 64, 3, 3, 3, generator=g) * 0.03).cuda()
    assert rel_err(ops.conv3d(xs, wt, 2, True, impl=0), ops.conv3d(xs, wt, 2, True, impl=1)) < 3e-3
    w2 = (torch.randn(128, 64, 3, 3, 3, generator=g) * 0.03).cuda()
    assert rel_err(ops.conv3d(x, w2, 2, False, impl=0), ops.conv3d(x, w2, 2, False, impl=1)) < 3e-3


# ---------------------------------------------------------------- fused depth head
@pytest.mark.parametrize("dims", [(1, 8, 8, 16, 4), (2, 12, 5, 7, 4), (1, 6, 9, 11, 2)])
def test_depth_head_vs_torch(ops, dims):
    """trilinear upsample (align_corners=False) + softmax + expectation, fwd and deterministic bwd."""
    n, d, hc, wc, f = dims
    g = torch.Generator().manual_seed(sum(dims))
    cost = (torch.randn(n, 1, d, hc, wc, generator=g) * 3).requires_grad_(True)
    J, H, W = d * f, hc * f, wc * f
    z0, dz = 2.0, 0.2
    up = F.interpolate(cost, [J, H, W], mode='trilinear', align_corners=False)
    prob = F.softmax(up.squeeze(1), 1)
    z = (z0 + (torch.arange(J, dtype=torch.float32) + 0.5) * dz).view(1, -1, 1, 1)
    ref = (prob * z).sum(1)
    gy = torch.randn(ref.shape, generator=g)
    (gc_ref,) = torch.autograd.grad(ref, cost, gy)
    cc = cost.detach().cuda().requires_grad_(True)
    out = ops.depth_head(cc, (J, H, W), z0, dz)
    assert out.shape == ref.shape and max_err(out.cpu(), ref) < 2e-5
    (gc,) = torch.autograd.grad(out, cc, gy.cuda())
    assert max_err(gc.cpu(), gc_ref) < 2e-5
    (gc2,) = torch.autograd.grad(ops.depth_head(cc, (J, H, W), z0, dz), cc, gy.cuda())
    assert torch.equal(gc, gc2)


def test_depth_head_x4_kernels_against_the_generic_ones(ops):
    """The ratio-4 specialisation (KITTI: 48x96x312 -> 192x384x1248) against the generic kernels on the same input:
    same lerp and accumulation order; only the softmax shift differs (maximum of the coarse column instead of the
    fine one), i.e. last-ulp differences."""
    from eval_driving_safety_b200 import _lib
    g = torch.Generator().manual_seed(17)
    res = {}
    for dims in ((1, 12, 24, 40), (2, 5, 7, 9)):
        n, d, hc, wc = dims
        cost = (torch.randn(n, 1, d, hc, wc, generator=g) * 4).cuda().requires_grad_(True)
        gy = torch.randn(n, 4 * hc, 4 * wc, generator=g).cuda()
        for flag in (0, 1):
            _lib.set_flag("depth_head_x4", flag)
            try:
                out = ops.depth_head(cost, (4 * d, 4 * hc, 4 * wc), 2.0, 0.2)
                (gc,) = torch.autograd.grad(out, cost, gy)
            finally:
                _lib.set_flag("depth_head_x4", None)
            res[flag] = (out, gc)
        assert max_err(res[1][0], res[0][0]) < 3e-6 * res[0][0].abs().max().item()        # a few ulps
        assert max_err(res[1][1], res[0][1]) < 5e-6 * max(1.0, res[0][1].abs().max().item())


def test_bev_pool_vs_torch(ops):
    g = torch.Generator().manual_seed(31)
    for ydim, p in ((8, 2), (8, 4), (10, 4), (7, 3)):      # the last two leave rows past the last whole window
        v = torch.randn(2, 64, 6, ydim, 5, generator=g, requires_grad=True)
        t = F.avg_pool3d(v, (1, p, 1))
        n, c, zz, yy, xx = t.shape
        ref = t.permute(0, 1, 3, 2, 4).reshape(n, c * yy, zz, xx)
        gy = torch.randn(ref.shape, generator=g)
        (gv_ref,) = torch.autograd.grad(ref, v, gy)
        vc = v.detach().cuda().requires_grad_(True)
        out = ops.bev_pool(vc, p)
        assert out.shape == ref.shape and max_err(out.cpu(), ref) < 1e-6
        (gv,) = torch.autograd.grad(out, vc, gy.cuda())
        assert max_err(gv.cpu(), gv_ref) < 1e-6


@pytest.mark.parametrize("c", [32, 64])
def test_grid_sample_bwd_both_row_mappings(ops, c):
    """The CSR backward has two thread mappings (sub-warp per cell / whole warp per cell with split
    rows); both must give the reference gradient and each must be bitwise reproducible."""
    g = torch.Generator().manual_seed(40 + c)
    x = torch.randn(1, c, 6, 7, generator=g, requires_grad=True)
    grid = torch.rand(1, 40, 50, 2, generator=g) * 2.2 - 1.1          # ~48 samples per input pixel
    ref = F.grid_sample(x, grid, mode='bilinear', padding_mode='zeros', align_corners=True)
    gy = torch.randn(ref.shape, generator=g)
    (gx_ref,) = torch.autograd.grad(ref, x, gy)
    xc = x.detach().cuda().requires_grad_(True)
    for long_rows in (0, 1):
        plan = ops.GridPlan(grid.cuda(), (6, 7), True)
        plan.long_rows = long_rows
        res = []
        for _ in range(2):
            out = ops.grid_sample(xc, grid.cuda(), True, plan)
            res.append(torch.autograd.grad(out, xc, gy.cuda())[0])
        assert max_err(res[0].cpu(), gx_ref) < 2e-4 and torch.equal(res[0], res[1])


def test_fused_lift_is_bit_identical_to_the_two_plain_kernels(ops):
    """b2_lift_fwd (one launch for the trilinear PSV sample and the bilinear image sample, 4 lanes per voxel with 4
    float4 each) only changes WHICH thread computes what: the result must equal the plain per-op kernels bit for
    bit and F.grid_sample on the CPU to 1e-5; the wide-lane gather backward must match autograd to 1e-4 and be
    deterministic."""
    g = torch.Generator().manual_seed(41)
    img = torch.randn(2, 32, 12, 16, generator=g, requires_grad=True)
    psv = torch.randn(2, 64, 6, 12, 16, generator=g, requires_grad=True)
    grid3 = torch.rand(2, 21, 5, 19, 3, generator=g) * 2.4 - 1.2           # voxel count not a multiple of the block
    grid2 = grid3[..., :2].reshape(2, 21 * 5, 19, 2)
    ref = torch.cat([F.grid_sample(psv, grid3, mode='bilinear', padding_mode='zeros', align_corners=True),
                     F.grid_sample(img, grid2, mode='bilinear', padding_mode='zeros', align_corners=True).view(2, 32, 21, 5, 19)], 1)
    gy = torch.randn(ref.shape, generator=g)
    gp_ref, gi_ref = torch.autograd.grad(ref, [psv, img], gy)
    res = {}
    saved = ops.LIFT_FUSED
    try:
        for fused in (False, True):
            ops.LIFT_FUSED = fused
            pc, ic = psv.detach().cuda().requires_grad_(True), img.detach().cuda().requires_grad_(True)
            out = ops.lift(pc, ic, grid3.cuda())
            res[fused] = (out,) + torch.autograd.grad(out, [pc, ic], gy.cuda())
    finally:
        ops.LIFT_FUSED = saved
    for a, b in zip(res[False], res[True]):
        assert torch.equal(a, b)
    out, gp, gi = res[True]
    assert max_err(out.cpu(), ref) < 1e-5 and max_err(gp.cpu(), gp_ref) < 1e-4 and max_err(gi.cpu(), gi_ref) < 1e-4
