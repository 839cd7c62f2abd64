"""CPU tests of the host-side pieces added in round 2: fake ground truth of the patch attacks, patch
resume, loss assembly against the reference's own lines (goldens from tests/golden/make_golden.py, which
EXECUTES the cited reference lines), module swapping, the asynchronous image writer, the calibration key of
the graph engine, and a hand-computed geometry fixture that pins the oracle's and the product's
projection code independently of each other."""
import math
import os
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def g2():
    z = np.load(os.path.join(ROOT, "tests", "golden", "attack_round2.npz"))

    def case(name):
        return {k.split("/", 1)[1]: (torch.from_numpy(z[k]) if z[k].ndim > 0 else z[k].item())
                for k in z.files if k.startswith(name + "/")}
    return case


# ---------------------------------------------------------------- A8: fake ground truth
def test_fake_gt_matches_reference_lines(g2):
    """attack/DSGN/patch_attack.py:336-354."""
    from eval_driving_safety_b200 import attack
    c = g2("fake_gt")
    bbox, box3d = attack.inject_fake_gt(c["bbox_in"].clone(), c["box3d_in"].clone())
    assert torch.equal(bbox, c["bbox_out"]) and torch.equal(box3d, c["box3d_out"])
    assert attack.FAKE_GT_BOX3D[5] == 29.11 and attack.FAKE_GT_BBOX[0] == 569.33      # the literals of :343-355


def test_stereo_rcnn_fake_gt_matches_reference_lines(g2):
    """attack/Stereo-RCNN/patch_attack.py:188-207."""
    from eval_driving_safety_b200 import attack
    c = g2("srcnn_fake_gt")
    gl, gr, gm, nb = attack.stereo_rcnn_fake_gt(c["center_l"].tolist(), c["center_r"].tolist(), c["radius"])
    assert torch.equal(gl, c["gt_left"]) and torch.equal(gr, c["gt_right"]) and torch.equal(gm, c["gt_merge"])
    assert int(nb) == int(c["num_boxes"]) == 1


def test_fake_gt_labels_cover_the_fake_car():
    from eval_driving_safety_b200 import attack, dsgn, synthetic
    cfg = dsgn.default_cfg()
    bbox, box3d = synthetic.make_targets(6, 1)
    attack.inject_fake_gt(bbox, box3d)
    lab = synthetic.labels_from_box3d(cfg, box3d)
    h, w, l, x, y, z, th = attack.FAKE_GT_BOX3D
    k = int(round(th / (math.pi / 2))) % cfg.num_anchors
    assert lab["cls"].shape == (1, 4, 192, 304) and lab["reg"].shape == (1, 28, 192, 304)
    assert lab["cls"][0, k].sum() > 0 and lab["cls"].sum() == lab["cls"][0, k].sum()      # one box, one yaw bin
    iz, ix = int((z - cfg.z_range[0]) / cfg.voxel), int((x - cfg.x_range[0]) / cfg.voxel)
    assert lab["cls"][0, k, iz, ix] == 1                                                 # the cell holding the centre
    # positives ~ footprint area / cell area
    assert abs(lab["cls"].sum().item() - l * w / cfg.voxel ** 2) < 0.15 * l * w / cfg.voxel ** 2
    assert abs(lab["reg"][0, k * 7 + 5, iz, ix].item() - math.log(l)) < 1e-6
    assert 0 < lab["ctr"][0, k, iz, ix] <= 1
    # the zeroed real boxes contribute nothing
    lab0 = synthetic.labels_from_box3d(cfg, torch.zeros(3, 7))
    assert lab0["cls"].sum() == 0


# ---------------------------------------------------------------- patch resume
def test_init_patch_resume_matches_reference_cv2_resize(g2, tmp_path):
    """attack/DSGN/patch_attack.py:211-234: resume + cv2.INTER_LINEAR resize (61 -> 77) and the fresh patch."""
    from eval_driving_safety_b200 import attack
    c = g2("init_patch")
    d = tmp_path / "run"
    os.makedirs(d / "epoch0")
    np.save(d / "epoch0" / "patch.npy", c["src"].numpy())
    dim, radius, patch = attack.init_patch(0.2, str(d))
    assert (dim, radius) == (c["dim"], c["radius"]) == (77, 38)
    assert patch.shape == (1, 3, 77, 77) and patch.dtype == np.float32
    # same half-pixel-centre bilinear map; cv2 and torch round their fp32 interpolation weights differently (3e-6 measured)
    assert np.abs(patch - c["out"].numpy()).max() < 1e-5
    dim, radius, fresh = attack.init_patch(0.2, str(tmp_path / "new"))
    assert (dim, radius) == (c["fresh_dim"], c["fresh_radius"]) and fresh.shape == tuple(c["fresh_shape"].tolist())
    assert np.abs(fresh).sum() == c["fresh_sum"] == 0 and os.path.exists(tmp_path / "new" / "epoch0" / "patch.npy")
    # Stereo R-CNN: ratio 0.1 of 600 -> 61, loaded as is (attack/Stereo-RCNN/patch_attack.py:58-76)
    dim, radius, same = attack.init_patch(0.1, str(d), short_side=600, resize=False)
    assert (dim, radius) == (61, 30) and np.array_equal(same, c["src"].numpy())


# ---------------------------------------------------------------- A3: loss assembly
def test_depth_loss_assembly_matches_reference_lines(g2):
    """attack/DSGN/pgd_attack.py:269, 310-319 executed on the eval-mode output dict: the product's graph-safe
    masked-mean form and the oracle's ``pred[mask]`` form both reproduce the reference's value."""
    from eval_driving_safety_b200 import dsgn
    from oracle import dsgn_ref
    c = g2("loss_assembly")
    cfg = types.SimpleNamespace(min_depth=2.0, max_depth=40.4, loss_disp=True, RPN3D_ENABLE=False)
    out = {"depth_preds": c["pred"]}
    got_p = dsgn.attack_loss(cfg, out, c["disp_true"], None)
    got_o = dsgn_ref.attack_loss(cfg, out, c["disp_true"], None)
    assert abs(got_p.item() - float(c["loss"])) < 1e-6 and abs(got_o.item() - float(c["loss"])) < 1e-6
    mask = (c["disp_true"] > 2.0) & (c["disp_true"] <= 40.4)
    assert torch.equal(mask.float(), c["mask"]) and mask[0, 0, 0] and not mask[0, 0, 1]
    # batch of 2 pairs: every pair gets what the reference gives its one pair (weight 1.0, its own mask)
    pred2 = torch.cat([c["pred"], c["pred"].flip(-1)])
    disp2 = torch.cat([c["disp_true"], c["disp_true"].flip(-1)])
    assert abs(dsgn.attack_loss(cfg, {"depth_preds": pred2}, disp2, None).item() - 2 * float(c["loss"])) < 2e-6


def test_output_dict_follows_the_eval_mode_convention():
    """The reference iterates outputs['depth_preds'] (pgd_attack.py:311): with the eval-mode [1,H,W] tensor that
    is one [H,W] map and weight index 3 - 1 + 0 = 2 (1.0)."""
    pred = torch.zeros(1, 4, 6)
    items = [torch.squeeze(o, 1) for o in pred]
    assert len(items) == 1 and items[0].shape == (4, 6) and [0.5, 0.7, 1.0][3 - len(items) + 0] == 1.0


# ---------------------------------------------------------------- module swapping (ADVICE)
def _tiny_net(bias=False, padding=1):
    return nn.Sequential(nn.Conv3d(32, 64, 3, 2, padding, bias=bias), nn.GroupNorm(32, 64),
                         nn.ConvTranspose3d(64, 32, 3, 2, 1, output_padding=1, bias=False), nn.GroupNorm(32, 32),
                         nn.Sequential(nn.Conv2d(32, 64, 3, 1, 2, 2, bias=True), nn.Conv2d(3, 32, 3, 2, 1, bias=False)))


def test_swap_modules_keeps_parameters_and_hyperparameters():
    from eval_driving_safety_b200 import modules as M
    net = _tiny_net()
    before = {k: v.data_ptr() for k, v in net.state_dict().items()}
    M.swap_modules(net)
    after = net.state_dict()
    assert list(after.keys()) == list(before.keys()) and all(after[k].data_ptr() == before[k] for k in before)
    assert type(net[0]) is M.Conv3dSm100 and net[0].stride == (2, 2, 2) and net[0].bias is None
    assert type(net[1]) is M.GroupNormSm100 and net[1].num_groups == 32 and net[1].relu is False
    assert type(net[2]) is M.ConvTranspose3dSm100 and net[2].output_padding == (1, 1, 1)
    assert type(net[4][0]) is M.Conv2dSm100 and net[4][0].dilation == (2, 2) and net[4][0].bias is not None
    assert type(net[4][1]) is M.Conv2dSm100 and net[4][1].in_channels == 3
    assert not any(p.requires_grad for p in net.parameters())
    with pytest.raises(RuntimeError):                     # no CPU fallback behind the drop-ins
        net[0](torch.zeros(1, 32, 4, 4, 4))


@pytest.mark.parametrize("kw,what", [(dict(bias=True), "bias"), (dict(padding=0), "padding")])
def test_swap_modules_refuses_layers_it_cannot_reproduce(kw, what):
    from eval_driving_safety_b200 import modules as M
    with pytest.raises(ValueError, match=what):
        M.swap_modules(_tiny_net(**kw))
    net = M.swap_modules(_tiny_net(**kw), strict=False)   # non-strict: the odd layer stays stock, the rest is swapped
    assert type(net[0]) is nn.Conv3d and type(net[2]) is M.ConvTranspose3dSm100


def test_swap_modules_validates_every_hyperparameter():
    from eval_driving_safety_b200 import modules as M
    bad = [nn.Conv3d(32, 32, 3, 1, 1, dilation=2, bias=False), nn.Conv3d(32, 32, 3, 1, 1, groups=2, bias=False),
           nn.Conv3d(32, 32, 5, 1, 2, bias=False), nn.ConvTranspose3d(32, 32, 3, 2, 1, output_padding=0, bias=False),
           nn.ConvTranspose3d(32, 32, 3, 1, 1, bias=False), nn.Conv2d(32, 32, 3, 1, 0, bias=False),
           nn.Conv2d(24, 32, 3, 1, 1, bias=False), nn.Conv2d(32, 32, 3, 2, 2, 2, bias=False), nn.Conv2d(3, 32, 3, 1, 1, bias=False),
           nn.GroupNorm(4, 32, affine=False)]
    for layer in bad:
        with pytest.raises(ValueError):
            M.swap_modules(nn.Sequential(layer))


# ---------------------------------------------------------------- asynchronous image dump
def test_async_image_writer_writes_the_same_bytes(tmp_path):
    from eval_driving_safety_b200 import kitti_io
    g = torch.Generator().manual_seed(3)
    imgs = [torch.randn(3, 24, 40, generator=g) for _ in range(6)]
    with kitti_io.AsyncImageWriter(workers=3) as wr:
        for k, im in enumerate(imgs):
            wr.submit(im, str(tmp_path / "a" / ("%06d.png" % k)), 36, 20)
            im.add_(1.0)                                  # the loop goes on modifying the image: the snapshot must hold
    for k, im in enumerate(imgs):
        os.makedirs(tmp_path / "s", exist_ok=True)
        kitti_io.save_image(im - 1.0, str(tmp_path / "s" / ("%06d.png" % k)), 36, 20)
        assert (tmp_path / "a" / ("%06d.png" % k)).read_bytes() == (tmp_path / "s" / ("%06d.png" % k)).read_bytes()


# ---------------------------------------------------------------- engine: calibration identity
def test_calib_key_distinguishes_calibrations():
    from eval_driving_safety_b200 import engine, synthetic
    a, b = synthetic.make_calib(1), synthetic.make_calib(1)
    c = synthetic.make_calib(1, cu=600.0)
    assert engine.calib_key(a) == engine.calib_key(b) != engine.calib_key(c)
    d = tuple(t.clone() for t in a)
    d[2][0, 0, 3] += 1e-9                                 # any change of P is another calibration
    assert engine.calib_key(d) != engine.calib_key(a)


# ---------------------------------------------------------------- geometry: hand-computed fixture
def test_geometry_against_hand_computed_values():
    """Pins BOTH implementations (oracle/dsgn_ref.py and eval_driving_safety_b200/dsgn.py) to values computed here
    from first principles with Python floats -- not to each other: plane depths, plane shifts
    s_d = f_u * b / (z_d * 4), voxel centres and their projection P [x y z 1]^T normalised to
    [-1, 1] with align_corners=True against the 96 x 312 feature map and the 48 PSV plane centres (A4.2)."""
    from eval_driving_safety_b200 import dsgn, synthetic
    from oracle import dsgn_ref
    f, cu, cv, b = 721.5377, 609.5593, 172.854, 0.54
    fu, base, P, PR = synthetic.make_calib(1)
    assert abs(float(fu) - f) < 1e-12 and abs(float(base) - b) < 1e-9
    z_of = lambda d: 2.0 + (d + 0.5) * 0.8                 # PSV planes: 0.2 m * downsample 4
    for mod in (dsgn, dsgn_ref):
        cfg = mod.default_cfg()
        zs = mod.psv_depths(cfg)
        assert zs.shape == (48,) and abs(zs[0].item() - 2.4) < 1e-6 and abs(zs[47].item() - z_of(47)) < 1e-5
        sh = mod.plane_shifts(cfg, fu, base)
        assert sh.shape == (1, 48)
        for d in (0, 13, 47):
            assert abs(sh[0, d].item() - f * b / (z_of(d) * 4)) < 2e-5, (mod.__name__, d)
        assert abs(sh[0, 0].item() - 40.5865) < 1e-3       # 721.5377 * 0.54 / 9.6
        grid = mod.lifting_grid(cfg, P, (96, 312))
        assert grid.shape == (1, 192, 20, 304, 3)
        for (iz, iy, ix) in ((0, 0, 0), (100, 7, 152), (191, 19, 303), (37, 12, 5)):
            x, y, z = -30.4 + (ix + 0.5) * 0.2, -1.0 + (iy + 0.5) * 0.2, 2.0 + (iz + 0.5) * 0.2
            u, v = f * x / z + cu, f * y / z + cv
            want = (2 * (u / 4) / 311 - 1, 2 * (v / 4) / 95 - 1, 2 * (z - z_of(0)) / (z_of(47) - z_of(0)) - 1)
            got = grid[0, iz, iy, ix].tolist()
            for a_, w_ in zip(got, want):
                assert abs(a_ - w_) < 1e-4 * max(1.0, abs(w_)), (mod.__name__, (iz, iy, ix), got, want)
        # the voxel straight ahead on the optical axis projects to the principal point
        g0 = grid[0, 50, 9, 151:153].mean(0)
        z50 = 2.0 + 50.5 * 0.2
        assert abs(g0[0].item() - (2 * (cu / 4) / 311 - 1)) < 1e-5
        assert abs(g0[1].item() - (2 * ((f * 0.9 / z50 + cv) / 4) / 95 - 1)) < 1e-5


# ---------------------------------------------------------------- detections hand-off (8f-3)
def test_decode_detections_inverts_the_target_assignment_and_writes_kitti_files(tmp_path):
    """labels_from_box3d (fake car of attack/DSGN/patch_attack.py:342-354) -> ideal head outputs -> decode_detections ->
    kitti_io.write_detections (predict_and_save_pgd.py:249-283 format) -> read back: the car comes out where it was put."""
    from eval_driving_safety_b200 import attack, dsgn, kitti_io, synthetic
    cfg = dsgn.default_cfg()
    bbox, box3d = synthetic.make_targets(4, 2)
    attack.inject_fake_gt(bbox, box3d)
    lab = synthetic.labels_from_box3d(cfg, box3d)
    out = {"bbox_cls": lab["cls"] * 20 - 10 + lab["ctr"] * lab["cls"], "bbox_reg": lab["reg"]}    # most confident at the centre
    _, _, P, _ = synthetic.make_calib(1)
    dets = dsgn.decode_detections(cfg, out, P[0], score_thresh=0.5, topk=5)
    assert 1 <= len(dets) <= 5 and dets[0]["score"] > 0.99
    h, w, l, x, y, z, th = attack.FAKE_GT_BOX3D
    d = dets[0]
    assert max(abs(a - b) for a, b in zip(d["hwl"], (h, w, l))) < 1e-4
    assert max(abs(a - b) for a, b in zip(d["center3d"], (x, y, z))) < 1e-4
    assert abs(math.remainder(d["ry"] - th, 2 * math.pi)) < 1e-4
    # the projected 2-D box contains the projection of the centre and lies around the reference's fake 2-D box
    u = 721.5377 * x / z + 609.5593
    assert d["bbox"][0] < u < d["bbox"][2] and abs((d["bbox"][0] + d["bbox"][2]) / 2 - (569.33 + 613.91) / 2) < 8
    path = kitti_io.write_detections(str(tmp_path), 7, dets)
    back = kitti_io.read_detections(path)
    assert os.path.basename(path) == "000007.txt" and len(back) == len(dets) and back[0]["type"] == "Car"
    assert abs(back[0]["location"][2] - z) < 1e-4 and abs(back[0]["location"][1] - (y + h / 2)) < 1e-4
    # nothing above the threshold -> empty file, as the reference writes for an image without detections
    none = dsgn.decode_detections(cfg, {"bbox_cls": torch.full_like(lab["cls"], -10.0), "bbox_reg": lab["reg"]}, P[0])
    assert none == [] and kitti_io.read_detections(kitti_io.write_detections(str(tmp_path), 8, none)) == []


# ---------------------------------------------------------------- GnLink: the hand-off of backward sums is self-validating
def _gnlink_graph(second_consumer, hold):
    """y = norm(a); consumers of y: a 'conv' whose backward writes the gradient g2 (and offers sums computed from it
    through the link) and, optionally, a plain second consumer.  Returns what the norm's backward saw."""
    from eval_driving_safety_b200.ops import GnLink
    seen = {}
    link = GnLink() if hold else None

    class Conv(torch.autograd.Function):
        @staticmethod
        def forward(ctx, y):
            return y * 2.0

        @staticmethod
        def backward(ctx, g):
            g2 = g * 2.0
            seen["offered"] = (g2.data_ptr(), g2._version)
            seen["offered_values"] = g2.clone()
            if link is not None:
                link.offer(torch.ones(1), g2)
            return g2

    class Norm(torch.autograd.Function):
        @staticmethod
        def forward(ctx, a):
            return a + 1.0

        @staticmethod
        def backward(ctx, gy):
            seen["arrived"] = (gy.data_ptr(), gy._version)
            seen["arrived_values"] = gy.clone()
            seen["taken"] = link.take(gy) if link is not None else None
            return gy

    a = torch.randn(64, 64, requires_grad=True)
    y = Norm.apply(a)
    w = torch.randn(64, 64)
    if second_consumer == "before":
        other = (y * w).sum()
        loss = Conv.apply(y).sum() + other
    elif second_consumer == "after":
        loss = Conv.apply(y).sum() + (y * w).sum()
    else:
        loss = Conv.apply(y).sum()
    torch.autograd.grad(loss, a)
    return seen


def test_gnlink_hands_the_sums_over_only_for_the_untouched_gradient():
    """Single consumer: the norm's backward receives the very tensor the conv wrote -> the sums are taken.
    Second consumer without the fork: the engine adds the two gradients; because the link holds the conv's tensor the
    sum always arrives as ANOTHER tensor (never an in-place update of the offered one), and the sums are dropped."""
    s = _gnlink_graph(None, hold=True)
    assert s["arrived"] == s["offered"] and s["taken"] is not None
    for order in ("before", "after"):
        s = _gnlink_graph(order, hold=True)
        assert not torch.equal(s["arrived_values"], s["offered_values"])          # the accumulated gradient ...
        assert s["arrived"][0] != s["offered"][0] and s["taken"] is None          # ... is not the offered tensor


def test_autograd_accumulates_in_place_into_an_unreferenced_gradient():
    """The hazard the link guards against (documented behaviour of the engine's input buffer): without an extra
    reference, a second consumer's gradient can be added IN PLACE into the conv's gradient tensor -- same address,
    other values.  If this ever stops being true the guard is merely redundant, so only the safe direction is asserted:
    whenever the addresses coincide, the values must have changed (i.e. address equality alone proves nothing)."""
    hit = False
    for order in ("before", "after"):
        s = _gnlink_graph(order, hold=False)
        if s["arrived"][0] == s["offered"][0]:
            hit = True
            assert not torch.equal(s["arrived_values"], s["offered_values"])
    if not hit:
        pytest.skip("this torch build accumulated out of place in both orders")


# ---------------------------------------------------------------- pure-torch glue that also runs on the CPU
def test_spp_upsample_as_two_gemms_equals_bilinear_interpolate_on_cpu():
    """dsgn.upsample_bilinear_matmul (the SPP branches' upsampling as two GEMMs on the channels-last memory) against
    F.interpolate(bilinear, align_corners=False), forward and backward, incl. the 1x4 map of the largest pool."""
    import torch.nn.functional as F
    from eval_driving_safety_b200 import dsgn
    g = torch.Generator().manual_seed(13)
    for (h, w, size) in ((1, 4, (24, 78)), (3, 9, (24, 78)), (6, 19, (24, 78)), (2, 3, (7, 11))):
        x = torch.randn(2, 8, h, w, generator=g).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        ref = F.interpolate(x, size, mode="bilinear", align_corners=False)
        out = dsgn.upsample_bilinear_matmul(x, size)
        assert out.shape == ref.shape and out.permute(0, 2, 3, 1).is_contiguous()
        assert (out - ref).abs().max().item() < 1e-5
        gy = torch.randn(ref.shape, generator=g)
        (a,) = torch.autograd.grad(out, x, gy)
        (b,) = torch.autograd.grad(ref, x, gy)
        assert (a - b).abs().max().item() < 1e-4 * b.abs().max().item()


def test_split_batch_backward_is_one_concatenation():
    from eval_driving_safety_b200 import ops
    g = torch.Generator().manual_seed(14)
    x = torch.randn(3, 8, 5, 6, generator=g).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wa, wb = torch.randn(1, 8, 5, 6, generator=g), torch.randn(2, 8, 5, 6, generator=g)
    a, b = ops.split_batch(x, 1)
    assert torch.equal(a, x[:1]) and torch.equal(b, x[1:])
    (got,) = torch.autograd.grad((a * wa).sum() + (b * wb).sum(), x)
    (ref,) = torch.autograd.grad((x[:1] * wa).sum() + (x[1:] * wb).sum(), x)
    assert torch.equal(got, ref)
    (only_b,) = torch.autograd.grad((ops.split_batch(x, 1)[1] * wb).sum(), x)
    assert torch.equal(only_b[1:], wb) and only_b[:1].abs().max().item() == 0
