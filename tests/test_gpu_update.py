"""GPU parity of the fused perturbation / patch kernels (subsystem 4) through the
C ABI: bit-exact against the golden vectors (reference lines executed verbatim)
and against the oracle on seeded inputs, incl. KITTI full size and ragged sizes."""
import random

import pytest
import torch

from oracle import attack_ref as A

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def atk(built_lib):
    from eval_driving_safety_b200 import attack
    return attack


def dev(*ts):
    return [t.cuda() for t in ts]


def test_pgd_step_golden_bit_exact(atk, golden):
    for k in range(5):
        c = golden.case("pgd%d" % k)
        xL, gL, cL, xR, gR, cR = dev(c["xL"], c["gL"], c["cleanL"], c["xR"], c["gR"], c["cleanR"])
        oL, oR = atk.pgd_step_pair(xL, gL, cL, xR, gR, cR, c["alpha"], c["eps"])
        assert torch.equal(oL.cpu(), c["outL"]) and torch.equal(oR.cpu(), c["outR"]), k
        assert torch.equal(atk.pgd_step(xL, gL, cL, c["alpha"], c["eps"]).cpu(), c["outL"])


def test_stereo_rcnn_step_golden_bit_exact(atk, golden):
    for k in range(2):
        c = golden.case("srcnn%d" % k)
        x, g, cl = dev(c["x"], c["g"], c["clean"])
        assert torch.equal(atk.stereo_rcnn_pgd_step(x, g, cl, c["alpha"], c["eps"]).cpu(), c["out"])


@pytest.mark.parametrize("shape", [(8, 3, 384, 1248), (2, 3, 7, 13), (1, 3, 600, 1987), (3, 3, 1, 1)])
def test_pgd_step_vs_oracle_bit_exact(atk, shape):
    g = torch.Generator().manual_seed(sum(shape))
    clean = torch.rand(shape, generator=g)
    x = A.normalize((clean + 0.1 * (torch.rand(shape, generator=g) - 0.5)).clamp(0, 1))
    gr = torch.randn(shape, generator=g)
    gr[..., :1] = 0
    for alpha, eps in [(1 / 255, 0.3), (0.0075, 0.03)]:
        ref = A.pgd_step_linf(x, gr, clean, alpha, eps)
        out = atk.pgd_step(x.cuda(), gr.cuda(), clean.cuda(), alpha, eps)
        assert torch.equal(out.cpu(), ref)
    # in-place pair variant == out-of-place
    xL, xR = x.cuda().clone(), x.cuda().flip(0).contiguous()
    oL, oR = atk.pgd_step_pair(xL, gr.cuda(), clean.cuda(), xR, gr.cuda(), clean.cuda(), 0.01, 0.05, inplace=True)
    assert oL.data_ptr() == xL.data_ptr()
    assert torch.equal(oL.cpu(), A.pgd_step_linf(x, gr, clean, 0.01, 0.05))


def test_pgd_step_unaligned_scalar_path(atk):
    g = torch.Generator().manual_seed(5)
    base = torch.rand(1 + 2 * 3 * 5 * 7, generator=g).cuda()
    x = base[1:].view(2, 3, 5, 7)                      # 4-byte-aligned only -> scalar kernel
    gr, clean = torch.randn(2, 3, 5, 7, generator=g).cuda(), torch.rand(2, 3, 5, 7, generator=g).cuda()
    ref = A.pgd_step_linf(x.cpu(), gr.cpu(), clean.cpu(), 0.02, 0.1)
    assert torch.equal(atk.pgd_step(x, gr, clean, 0.02, 0.1).cpu(), ref)


def test_pgd_known_answers(atk):
    g = torch.Generator().manual_seed(11)
    clean = torch.rand(2, 3, 32, 48, generator=g)
    x = A.normalize(clean)
    zero = torch.zeros_like(x)
    out = atk.pgd_step(x.cuda(), zero.cuda(), clean.cuda(), 0.1, 0.3).cpu()      # grad = 0 -> sign = 0
    assert torch.allclose(A.denormalize(out), clean, atol=1e-6)
    gr = torch.randn(2, 3, 32, 48, generator=g)
    out = atk.pgd_step(x.cuda(), gr.cuda(), clean.cuda(), 0.1, 0.0).cpu()        # eps = 0 -> clean
    assert torch.allclose(A.denormalize(out), clean, atol=1e-6)
    out01 = A.denormalize(atk.pgd_step(x.cuda(), gr.cuda(), clean.cuda(), 1.0, 0.25).cpu())
    inside = (clean > 0.26) & (clean < 0.74)
    assert torch.allclose((out01 - clean).abs()[inside], torch.tensor(0.25), atol=1e-6)
    assert out01.min() >= -1e-6 and out01.max() <= 1 + 1e-6


def test_pgd_step_l2_vs_oracle(atk):
    g = torch.Generator().manual_seed(13)
    clean = torch.rand(3, 3, 40, 56, generator=g)
    x = A.normalize((clean + 0.05 * (torch.rand(3, 3, 40, 56, generator=g) - 0.5)).clamp(0, 1))
    gr = torch.randn(3, 3, 40, 56, generator=g)
    for alpha, eps in [(0.5, 1.0), (2.0, 0.5)]:
        ref = A.pgd_step_l2(x, gr, clean, alpha, eps)
        out = atk.pgd_step(x.cuda(), gr.cuda(), clean.cuda(), alpha, eps, norm='l2').cpu()
        assert torch.allclose(out, ref, atol=2e-5, rtol=1e-5)                    # fp32, different reduction order
        eta = (A.denormalize(out) - clean).reshape(3, -1).norm(dim=1)
        assert (eta <= eps * (1 + 1e-4)).all()
    a = atk.pgd_step(x.cuda(), gr.cuda(), clean.cuda(), 0.5, 1.0, norm='l2')
    b = atk.pgd_step(x.cuda(), gr.cuda(), clean.cuda(), 0.5, 1.0, norm='l2')
    assert torch.equal(a, b)                                                     # deterministic reductions


def test_patch_apply_and_update_golden(atk, golden):
    c = golden.case("patch")
    r = int(c["radius"])
    cl, cr = [int(v) for v in c["center_l"]], [int(v) for v in c["center_r"]]

    def embed(box, ce):
        full = torch.zeros(1, 3, 384, 1248)
        full[:, :, ce[0] - r - 2:ce[0] + r + 3, ce[1] - r - 2:ce[1] + r + 3] = box
        return full

    crop = lambda t, ce: t[:, :, ce[0] - r - 2:ce[0] + r + 3, ce[1] - r - 2:ce[1] + r + 3]
    for box, blend, ce in ((c["imgL_box"], c["blendL_box"], cl), (c["imgR_box"], c["blendR_box"], cr)):
        img = embed(box, ce).cuda()
        atk.patch_apply(img, c["patch"].cuda(), ce, r)
        assert torch.equal(crop(img.cpu(), ce), blend)
    patch = c["patch"].cuda().clone()
    atk.patch_update(patch, embed(c["gL_box"], cl).cuda(), embed(c["gR_box"], cr).cuda(), cl, cr, r,
                     c["alpha"], c["eps"])
    assert torch.equal(patch.cpu(), c["patch_out"])
    delta = torch.empty_like(patch)
    p2 = c["patch"].cuda().clone()
    atk.patch_update(p2, embed(c["gL_box"], cl).cuda(), embed(c["gR_box"], cr).cuda(), cl, cr, r,
                     c["alpha"], c["eps"], delta_out=delta)
    assert torch.equal(p2.cpu(), c["patch"]) and torch.equal((p2 - delta).cpu(), c["patch_out"])


def test_stereo_rcnn_patch_update_golden(atk, golden):
    c = golden.case("srcnn_patch")
    r = int(c["radius"])
    cl, cr = [300, 900], [300, 836]

    def embed(box, ce):
        full = torch.zeros(1, 3, 600, 1987)
        full[:, :, ce[0] - r:ce[0] + r + 1, ce[1] - r:ce[1] + r + 1] = box
        return full

    patch = c["patch"].cuda().clone()
    lo = [0 - m for m in A.STEREO_RCNN_MEANS]
    hi = [255 - m for m in A.STEREO_RCNN_MEANS]
    atk.patch_update(patch, embed(c["gL_box"], cl).cuda(), embed(c["gR_box"], cr).cuda(), cl, cr, r,
                     c["alpha"], c["eps"], lo, hi)
    assert torch.equal(patch.cpu(), c["patch_out"])


def test_patch_vs_oracle_random_centres(atk):
    rng = random.Random(1)
    g = torch.Generator().manual_seed(17)
    dim, r = A.patch_dim_radius(384, 0.2)
    patch = torch.randn(1, 3, dim, dim, generator=g)
    for _ in range(3):
        cl, cr = atk.generate_round_mask(r, rng)
        rng2 = random.Random(1)
        assert 153 <= cl[0] <= 345 and 249 <= cl[1] <= 998 and cr == [cl[0], cl[1] - 64]
        img = torch.randn(1, 3, 384, 1248, generator=g)
        out = atk.patch_apply(img.cuda().clone(), patch.cuda(), cl, r).cpu()
        assert torch.equal(out, A.patch_apply(img, patch, cl, r))
        gl, gr = torch.randn(1, 3, 384, 1248, generator=g) * 1e-5, torch.randn(1, 3, 384, 1248, generator=g) * 1e-5
        p = patch.cuda().clone()
        atk.patch_update(p, gl.cuda(), gr.cuda(), cl, cr, r, 1e3, 8 / 255)
        assert torch.equal(p.cpu(), A.patch_update(patch, gl, gr, cl, cr, r, 1e3, 8 / 255))
    with pytest.raises(RuntimeError, match="leaves"):
        atk.patch_update(patch.cuda(), gl.cuda(), gr.cuda(), [10, 10], [10, 10], r, 1e3, 8 / 255)
