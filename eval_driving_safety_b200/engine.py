"""CUDA-graph engine for the attack iteration.

One PGD iteration of one stereo pair (forward, loss, backward to the pixels, fused pixel
update) has static shapes, ~1,400 kernel launches and is host-launch-bound when issued from
Python (measured: 55 ms eager vs the ~40 ms of kernel time).  The iteration is therefore captured
ONCE into a CUDA graph over static buffers and replayed for every pair and every iteration
(reference loop: attack/DSGN/pgd_attack.py:300-354).  All libb2attack kernels launch on the
current torch stream and never synchronise, so they capture like any torch op; the TMA
descriptors baked into the conv launches stay valid because graph-pool addresses are fixed.
"""
import collections

import torch

from . import attack, dsgn, ops

MAX_CALIB_GRAPHS = 3      # captured graphs kept alive per engine (each owns ~10 GB of activations per lane)


def calib_key(calib):
    """Hashable identity of a calibration tuple (fu, baseline, Proj, Proj_R): the byte image of its values."""
    return tuple(torch.as_tensor(c).detach().cpu().double().contiguous().numpy().tobytes() for c in calib)


class PgdIterationGraph:
    """step(xL, xR, cleanL, cleanR, disp) runs ONE PGD iteration in place on xL/xR ([1,3,H,W]
    normalised images; clean* are the denormalised clean copies) and returns the loss (device
    scalar, valid until the next step).

    The captured graph bakes in everything that depends only on the calibration (plane shifts, lifting
    grid, CSR gather plans).  Real KITTI frames carry per-frame P2/P3: pass ``calib=`` to ``step`` /
    ``step_multi`` and a frame whose calibration differs from the captured one is routed to a graph
    captured for ITS calibration (an LRU of ``MAX_CALIB_GRAPHS`` graphs) -- never silently replayed with
    the first frame's geometry."""

    def __init__(self, model, cfg, labels, calib, alpha, eps, example, norm='linf', use_graph=True, warmup=2,
                 lanes=1):
        """``lanes`` > 1 captures that many independent pair-iterations on parallel streams inside
        ONE graph: the launch/latency-bound 2-D sections of one pair overlap the tensor-core
        sections of the other (``step_multi``)."""
        self.model, self.cfg, self.labels, self.calib = model, cfg, labels, calib
        self.calib_key = calib_key(calib)
        self._siblings = collections.OrderedDict()        # calib_key -> engine captured for that calibration
        self.alpha, self.eps, self.norm = alpha, eps, norm
        self.use_graph = use_graph
        self.lanes = lanes if use_graph else 1
        self.launches_per_step = None
        xL, xR, cL, cR, disp = example
        if not use_graph:
            return
        self.sl = [[t.clone() for t in (xL, xR, cL, cR, disp)] for _ in range(self.lanes)]
        self.s = self.sl[0]
        cur = torch.cuda.current_stream()
        self.streams = [torch.cuda.Stream() for _ in range(self.lanes)]
        for st, bufs in zip(self.streams, self.sl):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                for _ in range(warmup):                  # cuDNN autotune, plan/pack caches, workspaces
                    self._iteration(*bufs)
            cur.wait_stream(st)
        torch.cuda.synchronize()
        for bufs in self.sl:
            for dst, src in zip(bufs, (xL, xR, cL, cR, disp)):
                dst.copy_(src)
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.LAUNCH_COUNT
        with torch.cuda.graph(self.graph):
            if self.lanes == 1:
                self.losses = [self._iteration(*self.s)]
            else:
                cap = torch.cuda.current_stream()
                self.losses, grads = [], []
                joint = self.norm == 'linf' and 2 * self.lanes <= 4
                for st, bufs in zip(self.streams, self.sl):
                    st.wait_stream(cap)
                    with torch.cuda.stream(st):
                        if joint:
                            loss, gL, gR = self._grads(bufs[0], bufs[1], bufs[4])
                            grads.append((gL, gR))
                            self.losses.append(loss)
                        else:
                            self.losses.append(self._iteration(*bufs))
                for st in self.streams:
                    cap.wait_stream(st)
                if joint:
                    # ONE pixel-update launch for all lanes (L and R image of every pair): 92 MB per launch
                    # instead of 46 MB -- the kernel is launch-latency bound at one pair
                    self._update_joint(self.sl, grads)
        self.loss = self.losses[0]
        self.launches_per_step = (ops.LAUNCH_COUNT - n0) // self.lanes
        for bufs in self.sl:
            for dst, src in zip(bufs, (xL, xR, cL, cR, disp)):
                dst.copy_(src)

    def _grads(self, xL, xR, disp):
        a, b = xL.detach().requires_grad_(True), xR.detach().requires_grad_(True)
        out = self.model(a, b, self.calib[0], self.calib[1], self.calib[2], calibs_Proj_R=self.calib[3])
        loss = dsgn.attack_loss(self.cfg, out, disp, self.labels)
        gL, gR = torch.autograd.grad(loss, [a, b])
        return loss.detach(), gL.contiguous(), gR.contiguous()

    def _update_joint(self, bufs_list, grads):
        xs, gs, cs = [], [], []
        for bufs, (gL, gR) in zip(bufs_list, grads):
            xs += [bufs[0], bufs[1]]; gs += [gL, gR]; cs += [bufs[2], bufs[3]]
        attack._pgd_update_sets(xs, gs, cs, xs, self.alpha, self.eps, attack.IMAGENET_MEAN, attack.IMAGENET_STD, 0.0, 1.0)

    def iterate_eager(self, pairs):
        """Eager (no graph) iteration of up to 2 pairs with the launch shapes of the captured graph -- one joint
        pixel update for all of them.  Used by bench.py's live kernel timing."""
        assert self.norm == 'linf' and len(pairs) <= 2
        losses, grads = [], []
        for p in pairs:
            loss, gL, gR = self._grads(p[0], p[1], p[4])
            losses.append(loss); grads.append((gL, gR))
        self._update_joint(pairs, grads)
        return losses

    def _iteration(self, xL, xR, cL, cR, disp):
        loss, gL, gR = self._grads(xL, xR, disp)
        if self.norm == 'linf':
            attack.pgd_step_pair(xL, gL, cL, xR, gR, cR, self.alpha, self.eps, inplace=True)
        else:
            attack.pgd_step(xL, gL, cL, self.alpha, self.eps, norm=self.norm, out=xL)
            attack.pgd_step(xR, gR, cR, self.alpha, self.eps, norm=self.norm, out=xR)
        return loss

    def _for_calib(self, calib, example):
        """The engine that serves ``calib``: this one, or a sibling captured for that calibration."""
        if calib is None:
            return self
        key = calib_key(calib)
        if key == self.calib_key:
            return self
        eng = self._siblings.get(key)
        if eng is None:
            while len(self._siblings) >= MAX_CALIB_GRAPHS - 1:
                self._siblings.popitem(last=False)        # least recently used graph and its memory pool
            eng = PgdIterationGraph(self.model, self.cfg, self.labels, calib, self.alpha, self.eps,
                                    tuple(t.clone() for t in example), norm=self.norm, use_graph=self.use_graph,
                                    lanes=self.lanes)
            self._siblings[key] = eng
        else:
            self._siblings.move_to_end(key)
        return eng

    def step_multi(self, pairs, calib=None):
        """``pairs`` = list of ``lanes`` tuples (xL, xR, cL, cR, disp): one iteration of each, concurrently.
        ``calib``: the calibration shared by these pairs (None = the captured one)."""
        eng = self._for_calib(calib, pairs[0])
        if eng is not self:
            return eng.step_multi(pairs)
        assert self.use_graph and len(pairs) == self.lanes
        for bufs, pr in zip(self.sl, pairs):
            for dst, src in zip(bufs, pr):
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        for bufs, pr in zip(self.sl, pairs):
            pr[0].copy_(bufs[0], non_blocking=True)
            pr[1].copy_(bufs[1], non_blocking=True)
        return self.losses

    def step(self, xL, xR, cL, cR, disp, calib=None):
        eng = self._for_calib(calib, (xL, xR, cL, cR, disp))
        if eng is not self:
            return eng.step(xL, xR, cL, cR, disp)
        if not self.use_graph:
            n0 = ops.LAUNCH_COUNT
            loss = self._iteration(xL, xR, cL, cR, disp)
            self.launches_per_step = ops.LAUNCH_COUNT - n0
            return loss
        if self.lanes > 1:
            # a lone pair (odd remainder): replaying the multi-lane graph would recompute idle lanes,
            # so a single-lane graph is captured on first use
            if getattr(self, "_single", None) is None:
                self._single = PgdIterationGraph(self.model, self.cfg, self.labels, self.calib, self.alpha,
                                                 self.eps, (xL, xR, cL, cR, disp), norm=self.norm, lanes=1)
            return self._single.step(xL, xR, cL, cR, disp)
        for dst, src in zip(self.s, (xL, xR, cL, cR, disp)):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        xL.copy_(self.s[0], non_blocking=True)
        xR.copy_(self.s[1], non_blocking=True)
        return self.loss


class PatchIterationGraph:
    """One iteration of the universal-patch attack (attack/DSGN/patch_attack.py:367-430) as a replayable CUDA
    graph over static buffers: blend the patch into both images -> forward -> loss -> backward to the pixels ->
    crop the two gradients at the patch boxes, clipped descent step (-> all_reduce(sum) of the 71 KB step over
    the ranks) -> apply it to the patch.

    The patch positions differ from image to image; they live in a device int32 buffer the patch kernels read
    at run time (``b2_patch_apply_dev`` / ``b2_patch_update_dev``), so ONE captured graph serves every image.
    ``load()`` brings the next image in, ``iterate()`` runs one iteration on it (the reference runs ``iters``
    = 2 per image, re-blending the updated patch into the already patched image)."""

    def __init__(self, model, cfg, labels, calib, patch, radius, example, alpha=1e3, eps=8.0 / 255, lo=None, hi=None,
                 use_graph=True, allreduce=None, warmup=2):
        """``patch`` [1,3,dim,dim] is updated in place.  ``example`` = (imgL, imgR, disp).  ``allreduce``:
        callable(delta) -> delta summed over the ranks (None: single rank, fused one-launch update)."""
        self.model, self.cfg, self.labels, self.calib = model, cfg, labels, calib
        self.patch, self.radius, self.alpha, self.eps, self.lo, self.hi = patch, int(radius), alpha, eps, lo, hi
        self.allreduce, self.use_graph = allreduce, use_graph
        dev = patch.device
        self.imgL, self.imgR, self.disp = (t.clone() for t in example)
        self.centres = torch.zeros(4, dtype=torch.int32, device=dev)
        self.delta = torch.zeros_like(patch)
        self.loss = None
        self.launches_per_step = None
        if not use_graph:
            return
        h, w = self.imgL.shape[-2:]
        self._set_centres([h // 2, w // 2], [h // 2, w // 2])
        keep = patch.clone()
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            for _ in range(warmup):                   # plan / pack caches, workspaces, NCCL communicator
                self._iteration()
        torch.cuda.current_stream().wait_stream(st)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.LAUNCH_COUNT
        with torch.cuda.graph(self.graph):
            self.loss = self._iteration()
        self.launches_per_step = ops.LAUNCH_COUNT - n0
        patch.copy_(keep)                             # the warm-up iterations must not train the patch
        self.imgL.copy_(example[0]); self.imgR.copy_(example[1])

    def _set_centres(self, cl, cr):
        # 16 bytes from pageable memory: the copy is staged synchronously, so the host values cannot change under it
        self.centres.copy_(torch.tensor([int(cl[0]), int(cl[1]), int(cr[0]), int(cr[1])], dtype=torch.int32))

    def _iteration(self):
        attack.patch_apply(self.imgL, self.patch, self.centres[0:2], self.radius)
        attack.patch_apply(self.imgR, self.patch, self.centres[2:4], self.radius)
        a, b = self.imgL.detach().requires_grad_(True), self.imgR.detach().requires_grad_(True)
        out = self.model(a, b, self.calib[0], self.calib[1], self.calib[2], calibs_Proj_R=self.calib[3])
        loss = dsgn.attack_loss(self.cfg, out, self.disp, self.labels)
        gL, gR = torch.autograd.grad(loss, [a, b])
        if self.allreduce is None:
            attack.patch_update(self.patch, gL.contiguous(), gR.contiguous(), self.centres, None, self.radius,
                                self.alpha, self.eps, self.lo, self.hi)
        else:
            attack.patch_update(self.patch, gL.contiguous(), gR.contiguous(), self.centres, None, self.radius,
                                self.alpha, self.eps, delta_out=self.delta)
            attack.patch_axpy(self.patch, self.allreduce(self.delta), self.lo, self.hi)
        return loss.detach()

    def load(self, imgL, imgR, disp, center_l, center_r):
        h, w = self.imgL.shape[-2:]
        for c in (center_l, center_r):
            if c[0] - self.radius < 0 or c[0] + self.radius >= h or c[1] - self.radius < 0 or c[1] + self.radius >= w:
                raise RuntimeError("patch box at %s leaves the %dx%d frame" % (c, h, w))
        self.imgL.copy_(imgL, non_blocking=True)
        self.imgR.copy_(imgR, non_blocking=True)
        self.disp.copy_(disp, non_blocking=True)
        self._set_centres(center_l, center_r)

    def iterate(self):
        """One iteration on the loaded image; returns the loss (device scalar, valid until the next call)."""
        if not self.use_graph:
            n0 = ops.LAUNCH_COUNT
            self.loss = self._iteration()
            self.launches_per_step = ops.LAUNCH_COUNT - n0
            return self.loss
        self.graph.replay()
        return self.loss
