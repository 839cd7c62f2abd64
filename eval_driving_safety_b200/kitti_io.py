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


class AsyncImageWriter:
    """Per-iteration image dump WITHOUT stalling the attack loop.  The reference copies both
    images to the host and PNG-encodes them inside every iteration (attack/DSGN/pgd_attack.py:357-374: a forced
    device sync + ~30 ms of encoding per image).  Here ``submit`` only enqueues a device-to-pinned-host copy on a
    side stream (ordered after the producing work by an event) and returns; worker threads wait for that copy,
    convert with the reference's ``tensor2im`` (truncating uint8 cast, (0,0,w,h) crop) and write the PNG.  The
    bytes on disk are identical to ``save_image``.  ``close()`` drains the queue."""

    def __init__(self, workers=4, max_pending=64):
        import queue
        import threading
        self._q = queue.Queue(maxsize=max_pending)
        self._threads = [threading.Thread(target=self._run, daemon=True) for _ in range(max(1, workers))]
        self._err = None
        self._stream = None
        for t in self._threads:
            t.start()

    def submit(self, img_norm, path, w, h):
        """``img_norm`` [3,H,W] (or [1,3,H,W]) normalised image, CUDA or CPU tensor; snapshotted at call time."""
        import torch
        if self._err is not None:
            raise self._err
        t = img_norm.detach()
        if t.dim() == 4:
            t = t[0]
        if t.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=t.device)
            host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            snap = t.clone()                       # device-side snapshot: the loop updates the image in place next
            self._stream.wait_stream(torch.cuda.current_stream(t.device))
            with torch.cuda.stream(self._stream):
                host.copy_(snap, non_blocking=True)
                done = torch.cuda.Event()
                done.record(self._stream)
            snap.record_stream(self._stream)
        else:
            host, done = t.clone(), None
        self._q.put((host, done, path, w, h))

    def _run(self):
        while True:
            item = self._q.get()
            try:
                if item is None:
                    return
                host, done, path, w, h = item
                if done is not None:
                    done.synchronize()
                os.makedirs(os.path.dirname(path), exist_ok=True)
                save_image(host, path, w, h)
            except Exception as e:          # surfaced by the next submit() / close()
                self._err = e
            finally:
                self._q.task_done()

    def close(self):
        self._q.join()
        for _ in self._threads:
            self._q.put(None)
        for t in self._threads:
            t.join()
        if self._err is not None:
            raise self._err

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


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
