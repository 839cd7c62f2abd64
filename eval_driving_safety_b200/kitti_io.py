"""Hand-off formats from the attack stage to the reference's UNCHANGED evaluation stage
(SURVEY 8f "next" row 3).  Host-side only: no kernels involved.

* per-iteration adversarial images: ``<dir>/dsgn_pgd_iters_{k}/image_{2,3}/%06d.png`` written with the
  reference's tensor2im convention (denormalise, *255, TRUNCATING uint8 cast) and its (0, 0, w, h)
  crop -- attack/DSGN/pgd_attack.py:157-193, 357-374;
* detections: one ``%06d.txt`` per image in the 16-column KITTI object format the reference's
  predict scripts write -- attack/DSGN/predict_and_save_pgd.py:249-283.

Both are pinned against goldens produced by executing the reference's own lines
(tests/golden/make_golden.py -> tests/golden/handoff.json).
"""
import os

import numpy as np

IMAGENET_MEAN = (0.485, 0.456, 0.406)     # attack/DSGN/pgd_attack.py:153
IMAGENET_STD = (0.229, 0.224, 0.225)      # :154

CLASS_NAMES = {1: "Pedestrian", 2: "Car"}  # anything else: 'Cyclist' (predict_and_save_pgd.py:273)


def tensor2im(img_norm):
    """[3,H,W] normalised image (tensor or array) -> HxWx3 uint8, pgd_attack.py:157-178: per channel
    x*std+mean, *255, transpose, ``astype(uint8)`` (truncation toward zero, no rounding, no clamp)."""
    a = np.array(img_norm.detach().cpu().float().numpy() if hasattr(img_norm, "detach") else img_norm,
                 dtype=np.float32, copy=True)
    if a.shape[0] == 1:                    # grayscale to RGB (:170-171)
        a = np.tile(a, (3, 1, 1))
    for i in range(3):
        a[i] = a[i] * IMAGENET_STD[i] + IMAGENET_MEAN[i]
    a = a * 255
    return np.transpose(a, (1, 2, 0)).astype(np.uint8)


def save_image(img_norm, path, w, h):
    """pgd_attack.py:181-193."""
    from PIL import Image
    Image.fromarray(tensor2im(img_norm)).crop((0, 0, w, h)).save(path)


def iteration_paths(save_dir, k, index):
    """(left, right) file names of iteration k, pgd_attack.py:357-374."""
    base = os.path.join(save_dir, "dsgn_pgd_iters_%d" % k)
    return (os.path.join(base, "image_2", "%06d.png" % index), os.path.join(base, "image_3", "%06d.png" % index))


def format_detection(cls, bbox, hwl, center3d, ry, score):
    """One line of the reference's detection file (predict_and_save_pgd.py:273-283):
    type, truncated = -1, occluded = -1, alpha = -atan2(x, z) + ry, 2-D box (4 x %.4f), h w l,
    x, y + h/2 (KITTI's bottom-centre convention), z, ry (%.6f), score (%.8f).
    ``center3d`` is the box CENTRE in camera coordinates, as the reference computes it from the
    8 corners (:259-261)."""
    h, w, l = (float(v) for v in hwl)
    # the reference holds the centre as float32 tensors and mixes them with Python floats: the two
    # derived columns are float32 results (torch / NumPy scalar promotion), reproduced explicitly
    x, y, z = (np.float32(v) for v in center3d)
    ry = float(ry)
    name = CLASS_NAMES.get(int(cls), "Cyclist")
    alpha = np.float32(np.float32(-np.arctan2(x, z)) + np.float32(ry))
    ybottom = np.float32(y + np.float32(h / 2.))
    b = [float(np.float32(v)) for v in bbox]
    return ('{} -1 -1 {:.4f} {:.4f} {:.4f} {:.4f} {:.4f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.8f}\n'
            .format(name, float(alpha), b[0], b[1], b[2], b[3], h, w, l, float(x), float(ybottom), float(z), ry,
                    float(np.float32(score))))


def write_detections(output_path, image_index, detections):
    """``detections``: iterable of dicts with keys cls, bbox, hwl, center3d, ry, score (a detection
    without a 3-D box uses hwl = center3d = (0,0,0), ry = 0 like :267-270).  Returns the file name
    ``<output_path>/%06d.txt`` (:250)."""
    os.makedirs(output_path, exist_ok=True)
    path = os.path.join(output_path, "{:06d}.txt".format(int(image_index)))
    with open(path, "w") as f:
        for d in detections:
            f.write(format_detection(d["cls"], d["bbox"], d.get("hwl", (0., 0., 0.)), d.get("center3d", (0., 0., 0.)),
                                     d.get("ry", 0.), d["score"]))
    return path


def read_detections(path):
    """Parse a detection file back (tests; KITTI column order)."""
    out = []
    with open(path) as f:
        for line in f:
            t = line.split()
            if len(t) != 16:
                raise ValueError("expected 16 columns, got %d: %r" % (len(t), line))
            v = [float(s) for s in t[1:]]
            out.append(dict(type=t[0], truncated=v[0], occluded=v[1], alpha=v[2], bbox=v[3:7], hwl=v[7:10],
                            location=v[10:13], ry=v[13], score=v[14]))
    return out
