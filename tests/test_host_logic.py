"""CPU: host-side logic -- weight packing for the gather modes, the oracle's own
consistency (gradcheck, PSMNet integer-shift form), sharding and the gloo gather."""
import os
import subprocess
import sys

import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT
from helpers import gather_conv_ref, max_err
from oracle import dsgn_ref as R
from eval_driving_safety_b200 import dsgn as P, ops, parallel, synthetic


@pytest.mark.parametrize("stride", [1, 2])
def test_conv_packing_matches_torch_conv3d_and_its_dgrad(stride):
    torch.manual_seed(0)
    x = torch.randn(1, 8, 4, 6, 8, dtype=torch.float64, requires_grad=True)
    w = torch.randn(12, 8, 3, 3, 3, dtype=torch.float64)
    y = F.conv3d(x, w, None, stride, 1)
    assert max_err(gather_conv_ref(x.detach(), ops._packed(w, "conv_fwd"), stride, 0), y) < 1e-10
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, gy)
    if stride == 1:
        got = gather_conv_ref(gy, ops._packed(w, "conv_dgrad_s1"), 1, 0)
    else:
        got = gather_conv_ref(gy, ops._packed(w, "conv_dgrad_s2"), 2, 1)
    assert max_err(got, gx) < 1e-10


def test_deconv_packing_matches_torch_conv_transpose3d_and_its_dgrad():
    torch.manual_seed(1)
    x = torch.randn(1, 8, 3, 4, 5, dtype=torch.float64, requires_grad=True)
    wt = torch.randn(8, 12, 3, 3, 3, dtype=torch.float64)
    y = F.conv_transpose3d(x, wt, None, 2, 1, output_padding=1)
    assert max_err(gather_conv_ref(x.detach(), ops._packed(wt, "deconv_fwd"), 2, 1), y) < 1e-10
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, gy)
    assert max_err(gather_conv_ref(gy, ops._packed(wt, "deconv_dgrad"), 2, 0), gx) < 1e-10


def test_oracle_cost_volume_integer_shift_equals_psmnet_slices():
    torch.manual_seed(2)
    l, r = torch.randn(2, 4, 5, 16), torch.randn(2, 4, 5, 16)
    shifts = torch.tensor([[0., 1., 3., 7.], [2., 2., 5., 15.]])
    cost = R.build_cost_volume(l, r, shifts)
    for n in range(2):
        for d in range(4):
            s = int(shifts[n, d])
            exp = torch.zeros(8, 5, 16)
            exp[:4, :, s:] = l[n, :, :, s:]
            exp[4:, :, s:] = r[n, :, :, :16 - s]
            assert torch.equal(cost[n, :, d], exp)


def test_oracle_cost_volume_gradcheck_fp64():
    torch.manual_seed(3)
    l = torch.randn(1, 2, 3, 8, dtype=torch.float64, requires_grad=True)
    r = torch.randn(1, 2, 3, 8, dtype=torch.float64, requires_grad=True)
    shifts = torch.tensor([[0.0, 1.25, 2.5, 6.75]], dtype=torch.float64)
    assert torch.autograd.gradcheck(lambda a, b: R.build_cost_volume(a, b, shifts), (l, r))


def test_oracle_and_product_models_share_parameters_and_geometry():
    cfg_r, cfg_p = R.tiny_cfg(), P.tiny_cfg()
    mr, mp = R.build_model(cfg_r), P.StereoNet(cfg_p)
    sr, sp = mr.state_dict(), mp.state_dict()
    assert list(sr.keys()) == list(sp.keys())
    assert all(sr[k].shape == sp[k].shape for k in sr)
    mp.load_state_dict({"module." + k: v for k, v in sr.items()})       # DataParallel prefix accepted
    fu, b, Pm, PR = synthetic.make_calib(1, scale=32 / 384, cu=32, cv=16)
    assert torch.equal(R.plane_shifts(cfg_r, fu, b), P.plane_shifts(cfg_p, fu, b))
    assert torch.equal(R.lifting_grid(cfg_r, Pm, (8, 16)), P.lifting_grid(cfg_p, Pm, (8, 16)))
    full = R.default_cfg()
    assert R.psv_depths(full).numel() == 48 and R.voxel_grid(full).shape == (192, 20, 304, 3)


def test_synthetic_pairs_are_seeded_and_kitti_shaped():
    a, b = synthetic.make_pair(3), synthetic.make_pair(3)
    assert torch.equal(a["imgL"], b["imgL"]) and a["imgL"].shape == (1, 3, 384, 1248)
    assert not torch.equal(a["imgL"], synthetic.make_pair(4)["imgL"])
    frac_invalid = (a["disp_L"] == 0).float().mean().item()
    assert 0.25 < frac_invalid < 0.35
    fu, base, Pm, PR = synthetic.make_calib(2)
    assert Pm.dtype == torch.float64 and abs(base[0].item() - 0.54) < 1e-9


def test_shard_pairs_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in parallel.shard_pairs(64, r, world))
        assert seen == list(range(64))
    assert parallel.shard_pairs(5, 1, 2) == [1, 3]


def test_gloo_world2_gather_and_patch_allreduce(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(
        "import sys, torch\n"
        "sys.path.insert(0, %r)\n"
        "from eval_driving_safety_b200 import parallel\n"
        "r, w = parallel.init('gloo')\n"
        "rows = [torch.tensor([float(i)] + [i * 10.0 + k for k in range(7)]) for i in parallel.shard_pairs(5)]\n"
        "allr = parallel.gather_stats(rows, 5)\n"
        "assert allr.shape == (5, 8) and allr[:, 0].tolist() == [0., 1., 2., 3., 4.], allr\n"
        "assert allr[3, 1].item() == 30.0\n"
        "d = parallel.allreduce_patch_delta(torch.full((3, 5, 5), float(r + 1)))\n"
        "assert torch.all(d == 3.0)\n"
        "parallel.barrier(); print('ok', r)\n" % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2
