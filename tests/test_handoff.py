"""Hand-off formats to the reference's unchanged evaluation stage, pinned against
goldens produced by executing the reference's own lines (tests/golden/make_golden.py: tensor2im =
attack/DSGN/pgd_attack.py:153-178, detection line = attack/DSGN/predict_and_save_pgd.py:273-283)."""
import json
import os

import numpy as np
import pytest
import torch

from eval_driving_safety_b200 import kitti_io

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(HERE, "golden", "handoff.json")) as f:
        return json.load(f)


def test_tensor2im_matches_reference_bytes(gold):
    img = torch.tensor(gold["tensor2im"]["img"], dtype=torch.float32)
    want = np.array(gold["tensor2im"]["out"], dtype=np.uint8)
    got = kitti_io.tensor2im(img)
    assert got.dtype == np.uint8 and got.shape == want.shape == (5, 7, 3)
    assert np.array_equal(got, want)
    assert np.array_equal(kitti_io.tensor2im(img.numpy()), want)         # array input
    # truncation, not rounding: 0.999 * 255 = 254.745 -> 254
    assert got[0, 1, 1] == 254


def test_runner_uses_the_same_conversion(gold):
    from eval_driving_safety_b200 import runner
    img = torch.tensor(gold["tensor2im"]["img"], dtype=torch.float32)
    assert np.array_equal(runner.tensor2im(img), kitti_io.tensor2im(img))


def test_detection_lines_match_reference_text(gold):
    for c in gold["detections"]:
        line = kitti_io.format_detection(c["cls"], c["bbox"], c["hwl"], c["center3d"], c["ry"], c["score"])
        assert line == c["line"], (line, c["line"])
        assert len(line.split()) == 16


def test_write_and_read_detections(tmp_path, gold):
    dets = [dict(cls=c["cls"], bbox=c["bbox"], hwl=c["hwl"], center3d=c["center3d"], ry=c["ry"], score=c["score"])
            for c in gold["detections"]]
    path = kitti_io.write_detections(str(tmp_path / "kitti_output"), 42, dets)
    assert os.path.basename(path) == "000042.txt"                          # '{:06d}.txt' (:250)
    with open(path) as f:
        assert f.read() == "".join(c["line"] for c in gold["detections"])
    back = kitti_io.read_detections(path)
    assert [b["type"] for b in back] == ["Pedestrian", "Car", "Cyclist", "Car", "Pedestrian", "Car", "Cyclist", "Car"]
    for b, c in zip(back, gold["detections"]):
        assert b["truncated"] == -1 and b["occluded"] == -1
        assert abs(b["location"][1] - (c["center3d"][1] + c["hwl"][0] / 2)) < 1e-5   # bottom-centre convention
        assert abs(b["score"] - c["score"]) < 1e-7
    # a detection without a 3-D box: the reference's zero defaults
    p2 = kitti_io.write_detections(str(tmp_path / "kitti_output"), 7, [dict(cls=2, bbox=[1, 2, 3, 4], score=0.5)])
    assert open(p2).read() == "Car -1 -1 0.0000 1.0000 2.0000 3.0000 4.0000 0.000000 0.000000 0.000000 0.000000 0.000000 0.000000 0.000000 0.50000000\n"
    assert kitti_io.write_detections(str(tmp_path / "kitti_output"), 8, []) and open(
        os.path.join(str(tmp_path / "kitti_output"), "000008.txt")).read() == ""


def test_iteration_paths_layout():
    l, r = kitti_io.iteration_paths("/x", 3, 17)
    assert l == "/x/dsgn_pgd_iters_3/image_2/000017.png" and r == "/x/dsgn_pgd_iters_3/image_3/000017.png"
