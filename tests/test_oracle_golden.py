"""CPU: the oracle restatement against golden vectors produced by executing the
reference's own lines (tests/golden/make_golden.py).  Bit-exact."""
import numpy as np
import torch

from oracle import attack_ref as A


def test_dsgn_pgd_step_matches_reference_lines(golden):
    for k in range(5):
        c = golden.case("pgd%d" % k)
        outL = A.pgd_step_linf(c["xL"], c["gL"], c["cleanL"], c["alpha"], c["eps"])
        outR = A.pgd_step_linf(c["xR"], c["gR"], c["cleanR"], c["alpha"], c["eps"])
        assert torch.equal(outL, c["outL"]) and torch.equal(outR, c["outR"]), k


def test_known_answers(golden):
    c = golden.case("pgd2")            # zero gradient: x unchanged up to the [0,1]/eps projection
    x01 = A.denormalize(c["xL"])
    out01 = A.denormalize(A.pgd_step_linf(c["xL"], torch.zeros_like(c["gL"]), c["cleanL"], 0.1, 10.0))
    assert torch.allclose(out01, x01.clamp(0, 1), atol=1e-6)
    c = golden.case("pgd3")            # eps == 0: result is the clean image
    out01 = A.denormalize(c["outL"])
    assert torch.allclose(out01, c["cleanL"], atol=1e-6)
    c = golden.case("pgd4")            # alpha >= 2 eps: |eta| == eps wherever grad != 0 (before the [0,1] clamp)
    out01 = A.denormalize(c["outL"])
    nz = c["gL"] != 0
    inside = (c["cleanL"] > 0.26) & (c["cleanL"] < 0.74) & nz
    assert torch.allclose((out01 - c["cleanL"]).abs()[inside], torch.tensor(0.25), atol=1e-6)


def test_stereo_rcnn_pgd_step(golden):
    for k in range(2):
        c = golden.case("srcnn%d" % k)
        out = A.stereo_rcnn_pgd_step(c["x"], c["g"], c["clean"], c["alpha"], c["eps"])
        assert torch.equal(out, c["out"])
        outR = A.stereo_rcnn_pgd_step(c["x"].flip(-1), c["g"].flip(-1), c["clean"].flip(-1), c["alpha"], c["eps"])
        assert torch.equal(outR, c["outR"])


def _embed(box, center, radius, h, w, margin):
    full = torch.zeros(1, 3, h, w)
    full[:, :, center[0] - radius - margin:center[0] + radius + margin + 1,
         center[1] - radius - margin:center[1] + radius + margin + 1] = box
    return full


def test_dsgn_patch_blend_and_update(golden):
    c = golden.case("patch")
    r = int(c["radius"])
    cl, cr = [int(v) for v in c["center_l"]], [int(v) for v in c["center_r"]]
    assert cr == [cl[0], cl[1] - 64]                                    # patch_attack.py:243
    assert A.patch_dim_radius(384, 0.2) == (77, 38)                     # :213-218
    imgL = _embed(c["imgL_box"], cl, r, 384, 1248, 2)
    blend = A.patch_apply(imgL, c["patch"], cl, r)
    crop = lambda t, ce: t[:, :, ce[0] - r - 2:ce[0] + r + 3, ce[1] - r - 2:ce[1] + r + 3]
    assert torch.equal(crop(blend, cl), c["blendL_box"])
    assert A.round_mask(cl, r, 384, 1248).sum().item() == c["mask_l_sum"]
    gL = _embed(c["gL_box"], cl, r, 384, 1248, 2)
    gR = _embed(c["gR_box"], cr, r, 384, 1248, 2)
    out = A.patch_update(c["patch"], gL, gR, cl, cr, r, c["alpha"], c["eps"])
    assert torch.equal(out, c["patch_out"])


def test_stereo_rcnn_patch_update(golden):
    c = golden.case("srcnn_patch")
    r = int(c["radius"])
    cl, cr = [300, 900], [300, 836]
    gL = _embed(c["gL_box"], cl, r, 600, 1987, 0)
    gR = _embed(c["gR_box"], cr, r, 600, 1987, 0)
    lo = [0 - m for m in A.STEREO_RCNN_MEANS]
    hi = [255 - m for m in A.STEREO_RCNN_MEANS]
    out = A.patch_update(c["patch"], gL, gR, cl, cr, r, c["alpha"], c["eps"], lo, hi)
    assert torch.equal(out, c["patch_out"])
    assert A.patch_dim_radius(600, 0.1) == (61, 30)                     # Stereo-RCNN/patch_attack.py:60-65


def test_roi_levels(golden):
    c = golden.case("roi_levels")
    assert torch.equal(A.roi_levels(c["rois"]), c["levels"])


def test_round_mask_integer_form_is_exact():
    # sqrt(dy^2+dx^2) <= r  <=>  dy^2+dx^2 <= r^2 for integer coordinates
    for r in (1, 5, 30, 38):
        m = A.round_mask([50, 60], r, 100, 120)[0, 0]
        yy, xx = torch.meshgrid(torch.arange(100), torch.arange(120), indexing='ij')
        assert torch.equal(m.bool(), ((yy - 50) ** 2 + (xx - 60) ** 2) <= r * r)
