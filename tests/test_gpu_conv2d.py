"""GPU parity of the 2-D convolution kernels (feature extractor / BEV head layers behind
attack/DSGN/pgd_attack.py:308, :336) through the C ABI against stock torch ops on the CPU in fp32
(what the reference computes): forward and data gradient, every layer class of the extractor."""
import pytest
import torch
import torch.nn.functional as F

from helpers import max_err, rel_err

pytestmark = pytest.mark.gpu

# tolerances: 3xTF32 split = fp32-class (products exact, fp32 accumulation in another order);
# plain TF32 = 10-bit mantissa operands
TOL_SPLIT, TOL_TF32 = 2e-5, 3e-3


@pytest.fixture(scope="module")
def ops(built_lib):
    from eval_driving_safety_b200 import ops
    return ops


# (n, cin, cout, h, w, k, stride, dilation, bias)
CASES = [
    (2, 32, 32, 24, 40, 3, 1, 1, False),       # firstconv[2], layer1
    (1, 32, 64, 32, 48, 3, 2, 1, False),       # layer2 block 0 conv1 (stride 2)
    (1, 32, 64, 32, 48, 1, 2, 1, False),       # layer2 block 0 downsample (1x1 stride 2)
    (2, 64, 64, 19, 27, 3, 1, 1, False),       # layer2, ragged tile edges
    (1, 64, 128, 17, 9, 1, 1, 1, False),       # layer3 downsample (1x1)
    (1, 128, 128, 20, 24, 3, 1, 2, False),     # layer4 (dilation 2)
    (1, 320, 128, 18, 26, 3, 1, 1, False),     # lastconv / bev_conv[0]: dgrad has 320 = 2 x 160 output channels
    (1, 128, 32, 3, 9, 1, 1, 1, False),        # SPP branch on a tiny map
    (1, 128, 64, 16, 40, 3, 1, 1, True),       # fused detection heads (bias)
    (1, 128, 16, 16, 16, 3, 1, 1, True),       # narrowest N tile (forward only: the data gradient's K is Cout)
    (1, 128, 28, 12, 20, 3, 1, 1, True),       # a detection head as the reference builds it: padded to 32 internally
]


@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("case", CASES)
def test_conv2d_fwd_dgrad_vs_torch(ops, case, split):
    n, cin, cout, h, w, k, stride, dil, has_bias = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(n, cin, h, w, generator=g, requires_grad=True)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g) if has_bias else None
    pad = dil * (k // 2)
    ref = F.conv2d(x, wt, b, stride, pad, dil)
    gy = torch.randn(ref.shape, generator=g)
    (gx_ref,) = torch.autograd.grad(ref, x, gy)
    ops.set_conv2d_split(split)
    saved_bwd = ops.CONV2D_SPLIT_BWD
    ops.CONV2D_SPLIT_BWD = 1                  # this test is about the kernels: split (or not) in BOTH directions
    try:
        xc = x.detach().cuda().requires_grad_(True)
        y = ops.conv2d(xc, wt.cuda(), b.cuda() if has_bias else None, stride, dil)
        gx = torch.autograd.grad(y, xc, gy.cuda())[0] if cout != 16 else None   # the data gradient's K is Cout
        if split and gx is not None:
            ops.CONV2D_SPLIT_BWD = 0          # the benchmarked default: forward split, plain-TF32 data gradient
            y0 = ops.conv2d(xc, wt.cuda(), b.cuda() if has_bias else None, stride, dil)
            (gx0,) = torch.autograd.grad(y0, xc, gy.cuda())
            assert torch.equal(y0, y) and rel_err(gx0.cpu(), gx_ref) < TOL_TF32
    finally:
        ops.set_conv2d_split(True)
        ops.CONV2D_SPLIT_BWD = saved_bwd
    tol = TOL_SPLIT if split else TOL_TF32
    assert y.shape == ref.shape
    print("conv2d %s split=%s: rel. error fwd %.2e dgrad %s" % (case, split, rel_err(y.cpu(), ref),
                                                              "%.2e" % rel_err(gx.cpu(), gx_ref) if gx is not None else "-"))
    assert rel_err(y.cpu(), ref) < tol, rel_err(y.cpu(), ref)
    assert max_err(y.cpu(), ref) < tol * 50
    if gx is not None:
        assert gx.shape == gx_ref.shape
        assert rel_err(gx.cpu(), gx_ref) < tol, rel_err(gx.cpu(), gx_ref)
        assert max_err(gx.cpu(), gx_ref) < tol * 50


def test_conv2d_tensor_core_truncates_tf32_operands(ops):
    """The in-kernel split feeds the RAW fp32 activation tile as x_hi and relies on kind::tf32 ignoring the low
    13 mantissa bits.  split = 2 rewrites the tile as (x & ~0x1fff) first: the results must be bit-identical."""
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 64, 40, 48, generator=g).cuda()
    wt = (torch.randn(128, 64, 3, 3, generator=g) / 24.0).cuda()
    try:
        ops.set_conv2d_split(1)
        y1 = ops.conv2d(x, wt)
        ops.set_conv2d_split(2)
        y2 = ops.conv2d(x, wt)
    finally:
        ops.set_conv2d_split(1)
    assert torch.equal(y1, y2)


def test_conv2d_split_is_fp32_class_on_a_kitti_size_layer(ops):
    """64 -> 64 3x3 at 96x312 (the layer class that dominates the extractor), L+R batch: against fp64 the
    3xTF32 result must be as accurate as a stock fp32 convolution; plain TF32 is ~1000x worse."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 64, 96, 312, generator=g)
    wt = torch.randn(64, 64, 3, 3, generator=g) / 24.0
    ref64 = F.conv2d(x.double(), wt.double(), None, 1, 1)
    err32 = rel_err(F.conv2d(x, wt, None, 1, 1), ref64)
    y3 = ops.conv2d(x.cuda(), wt.cuda())
    ops.set_conv2d_split(False)
    try:
        y1 = ops.conv2d(x.cuda(), wt.cuda())
    finally:
        ops.set_conv2d_split(True)
    e3, e1 = rel_err(y3.cpu(), ref64), rel_err(y1.cpu(), ref64)
    print("rel. error vs fp64: fp32 CPU %.2e, 3xTF32 %.2e, TF32 %.2e" % (err32, e3, e1))
    # measured: 3.2e-6 (stacked accumulators, the default) / < 1e-6 (split = 3); at the model level both give the
    # same parity as cuDNN fp32 (tests/diag/diag_fullsize.py): 98.40 / 98.35 / 98.38 % identical FGSM pixels
    assert e3 < 5e-6
    assert e1 > 20 * e3                                   # the split is what buys the accuracy
    assert torch.equal(y3, ops.conv2d(x.cuda(), wt.cuda()))   # deterministic


def test_conv2d_fork_adds_the_other_gradient(ops):
    g = torch.Generator().manual_seed(9)
    for (cin, cout, stride, k) in [(64, 64, 1, 3), (32, 64, 2, 3), (32, 64, 2, 1)]:
        x = torch.randn(1, cin, 24, 32, generator=g).cuda().requires_grad_(True)
        wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
        y, x2 = ops.conv2d_fork(x, wt, None, stride, 1)
        other = torch.randn(x.shape, generator=g).cuda()
        gy = torch.randn(y.shape, generator=g).cuda()
        (gx,) = torch.autograd.grad([y, (x2 * other).sum()], x, [gy, torch.ones((), device='cuda')])
        (gplain,) = torch.autograd.grad(ops.conv2d(x, wt, None, stride, 1), x, gy)
        assert max_err(gx, gplain + other) < 1e-5


@pytest.mark.parametrize("dims", [(2, 32, 64), (1, 384, 1248), (1, 31, 45)])
def test_first_layer_fwd_dgrad_vs_torch(ops, dims):
    n, h, w = dims
    g = torch.Generator().manual_seed(h)
    x = torch.randn(n, 3, h, w, generator=g, requires_grad=True)
    wt = torch.randn(32, 3, 3, 3, generator=g) / 27 ** 0.5
    ref = F.conv2d(x, wt, None, 2, 1)
    gy = torch.randn(ref.shape, generator=g)
    (gx_ref,) = torch.autograd.grad(ref, x, gy)
    xc = x.detach().cuda().requires_grad_(True)
    y = ops.conv2d(xc, wt.cuda(), None, 2, 1)
    (gx,) = torch.autograd.grad(y, xc, gy.cuda())
    assert gx.is_contiguous() and gx.shape == x.shape          # NCHW, what b2_pgd_update reads
    assert max_err(y.cpu(), ref) < 2e-5 and max_err(gx.cpu(), gx_ref) < 2e-5


def test_conv2d_halo_kernel_equals_the_tap_per_stage_kernel(ops, built_lib):
    """Stride-1 3x3 layers have two kernels (halo reuse, the default; one tap per stage): the same products summed in
    another order (kw-major instead of tap-major) -> equal to fp32 rounding, dilation 1 and 2, with and without the split."""
    from eval_driving_safety_b200 import _lib
    g = torch.Generator().manual_seed(21)
    for (cin, cout, dil) in ((64, 64, 1), (128, 128, 2), (320, 128, 1)):
        x = torch.randn(2, cin, 37, 29, generator=g).cuda()
        wt = (torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).cuda()
        res = []
        try:
            for halo in (1, 0):
                _lib.set_flag("conv2d_halo", halo)
                for split in (1, 0):
                    ops.set_conv2d_split(split)
                    res.append(ops.conv2d(x, wt, None, 1, dil))
        finally:
            _lib.set_flag("conv2d_halo", None)
            ops.set_conv2d_split(1)
        assert rel_err(res[0], res[2]) < 1e-5 and rel_err(res[1], res[3]) < 1e-5


def test_conv2d_rejects_unsupported(ops):
    x = torch.randn(1, 24, 8, 8).cuda()
    with pytest.raises(RuntimeError):
        ops.conv2d(x, torch.randn(32, 24, 3, 3).cuda())          # Cin % 32
    with pytest.raises(RuntimeError):
        ops.conv2d(torch.randn(1, 32, 8, 8).cuda(), torch.randn(32, 32, 5, 5).cuda())
    with pytest.raises(RuntimeError):
        ops.conv2d(torch.randn(1, 32, 9, 8).cuda(), torch.randn(32, 32, 3, 3).cuda(), None, 2, 1)   # odd dims, stride 2
    with pytest.raises(RuntimeError):
        ops.conv2d(torch.randn(1, 32, 8, 8), torch.randn(32, 32, 3, 3))                            # CPU tensors


# ---------------------------------------------------------------------------------------------------------------
# GroupNorm sums in the conv2d epilogue (upstream convbn = Conv2d + GroupNorm behind attack/DSGN/pgd_attack.py:308/:336)
# ---------------------------------------------------------------------------------------------------------------
def _cl2(shape, g):
    return torch.randn(shape, generator=g).cuda().contiguous(memory_format=torch.channels_last)


# (n, cin, cout, h, w, k, stride, dilation, bias)
STAT_CASES = [
    (2, 32, 32, 24, 40, 3, 1, 1, False),       # halo kernel, two samples, fewer tiles than CTAs (zero rows)
    (2, 32, 64, 200, 328, 3, 1, 1, False),     # halo kernel, more tiles than CTAs: a CTA's sequence crosses the samples
    (2, 32, 64, 64, 96, 3, 2, 1, False),       # generic kernel, stride 2
    (3, 64, 128, 37, 29, 1, 1, 1, True),       # generic kernel 1x1, three samples, ragged tiles, bias
    (1, 128, 128, 20, 24, 3, 1, 2, False),     # dilation 2
    (2, 320, 128, 18, 26, 3, 1, 1, False),     # widest K
]


@pytest.mark.parametrize("split", [1, 0])
@pytest.mark.parametrize("case", STAT_CASES)
def test_conv2d_epilogue_adds_up_the_groupnorm_statistics(ops, case, split):
    """partial[n, rows, 2, C] summed over rows = per-channel (sum, sum of squares) of the conv output; GroupNorm fed
    with it equals GroupNorm with its own statistics pass; the table is bitwise reproducible."""
    n, cin, cout, h, w, k, stride, dil, has_bias = case
    g = torch.Generator().manual_seed(sum(case))
    x = _cl2((n, cin, h, w), g)
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
    b = torch.randn(cout, generator=g).cuda() if has_bias else None
    gamma, beta = (torch.rand(cout, generator=g) + 0.5).cuda(), (torch.randn(cout, generator=g) * 0.3).cuda()
    ops.set_conv2d_split(split)
    try:
        y0 = ops.conv2d(x, wt, b, stride, dil)
        n0 = ops.LAUNCH_COUNT
        y, part = ops.conv2d_with_stats(x, wt, b, stride, dil)
        assert ops.LAUNCH_COUNT - n0 == 1
        assert part is not None and part.shape[0] == n and tuple(part.shape[2:]) == (2, cout)
        assert torch.equal(y, y0)
        tot = part.double().sum(1).cpu()
        yd = y.double().cpu()
        ref = torch.stack([yd.sum((2, 3)), (yd * yd).sum((2, 3))], 1)
        assert (tot - ref).abs().max().item() < 2e-5 * ref.abs().max().item()
        y2, part2 = ops.conv2d_with_stats(x, wt, b, stride, dil)
        assert torch.equal(part, part2)
        n1 = ops.LAUNCH_COUNT
        a = ops.groupnorm_act(y, gamma, beta, 32, 1e-5, relu=True, partial=part)
        fused_launches = ops.LAUNCH_COUNT - n1
        r = ops.groupnorm_act(y, gamma, beta, 32, 1e-5, relu=True)
        assert fused_launches == 2 and ops.LAUNCH_COUNT - n1 == 5
        assert (a - r).abs().max().item() < 2e-5
    finally:
        ops.set_conv2d_split(1)


@pytest.mark.parametrize("relu,with_res,second_consumer,case", [
    (True, False, False, (2, 64, 64, 40, 56, 3, 1)),      # norm+ReLU -> 3x3 conv: mask recomputed in the halo kernel's epilogue
    (False, True, False, (2, 64, 64, 40, 56, 3, 1)),      # norm + shortcut, no activation (BasicBlock tail)
    (True, False, True, (2, 32, 32, 200, 328, 3, 1)),     # the norm's output is forked (BasicBlock input); many tiles per CTA
    (True, False, False, (3, 128, 32, 19, 27, 1, 1)),     # 1x1 consumer (generic kernel), ragged, three samples
    (True, False, False, (1, 128, 128, 24, 40, 3, 2)),    # dilation-2 consumer
    (True, True, False, (2, 64, 64, 24, 40, 3, 1)),       # ReLU + residual needs the saved output: not fused, still right
    # a second consumer WITHOUT the fork: autograd adds the two gradients itself (in place into the conv's gradient
    # buffer if it holds the last reference to it) -- the sums of the conv's epilogue must then be dropped
    (True, False, "plain_after", (2, 64, 64, 24, 40, 3, 1)),
    (True, False, "plain_before", (2, 64, 64, 24, 40, 3, 1)),
])
def test_groupnorm2d_backward_sums_from_the_consuming_convs_dgrad(ops, relu, with_res, second_consumer, case):
    """conv_a -> GroupNorm -> conv_b on 2-D maps: conv_b's data-gradient launch adds up the norm's backward sums in its
    epilogue and conv_a's epilogue the forward statistics; the input gradient equals the unfused path."""
    n, c, c2, h, w, k, dil = case
    g = torch.Generator().manual_seed(sum(case))
    x = _cl2((n, c, h, w), g).requires_grad_(True)
    wa = (torch.randn(c, c, 3, 3, generator=g) * (9 * c) ** -0.5).cuda()
    wb = (torch.randn(c2, c, k, k, generator=g) * (k * k * c) ** -0.5).cuda()
    gamma, beta = (torch.rand(c, generator=g) + 0.5).cuda(), (torch.randn(c, generator=g) * 0.3).cuda()
    res = _cl2((n, c, h, w), g) if with_res else None
    gy, g2 = _cl2((n, c2, h, w), g), _cl2((n, c, h, w), g)

    def run(fuse):
        old = ops.FUSE_GN_BWD, ops.FUSE_GN_STATS
        ops.FUSE_GN_BWD = ops.FUSE_GN_STATS = fuse
        try:
            n0 = ops.LAUNCH_COUNT
            ya, part = ops.conv2d_with_stats(x, wa)
            assert (part is not None) == fuse
            hmap = ops.groupnorm_act(ya, gamma, beta, 32, 1e-5, relu=relu, res=res, partial=part)
            if second_consumer == "plain_before":
                other = (hmap * g2).sum()
                loss = (ops.conv2d(hmap, wb, None, 1, dil) * gy).sum() + other
            elif second_consumer == "plain_after":
                loss = (ops.conv2d(hmap, wb, None, 1, dil) * gy).sum() + (hmap * g2).sum()
            elif second_consumer:
                yb, h2 = ops.conv2d_fork(hmap, wb, None, 1, dil)
                loss = (yb * gy).sum() + (h2 * g2).sum()
            else:
                loss = (ops.conv2d(hmap, wb, None, 1, dil) * gy).sum()
            (gx,) = torch.autograd.grad(loss, x)
            return gx, ops.LAUNCH_COUNT - n0
        finally:
            ops.FUSE_GN_BWD, ops.FUSE_GN_STATS = old

    ref, n_ref = run(False)
    got, n_got = run(True)
    fusable = not (relu and with_res) and second_consumer not in ("plain_after", "plain_before")
    assert n_got == n_ref - 1 - (1 if fusable else 0)      # both statistics launches of the norm are gone
    # fp32 partial sums in another order (the norm's backward cancels sum gz*x against mean * sum gz)
    assert (got - ref).abs().max().item() < 5e-5 * ref.abs().max().item() + 1e-7
    got2, _ = run(True)
    assert torch.equal(got, got2)                          # reproducible
