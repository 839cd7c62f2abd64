"""CUDA-graph engine for the attack iteration.

One PGD iteration of one stereo pair (forward, loss, backward to the pixels, fused pixel
update) has static shapes, ~1,400 kernel launches and is host-launch-bound when issued from
Python (measured: 55 ms eager vs the ~40 ms of kernel time).  The iteration is therefore captured
ONCE into a CUDA graph over static buffers and replayed for every pair and every iteration
(reference loop: attack/DSGN/pgd_attack.py:300-354).  All libb2attack kernels launch on the
current torch stream and never synchronise, so they capture like any torch op; the TMA
descriptors baked into the conv launches stay valid because graph-pool addresses are fixed.
"""
import torch

from . import attack, dsgn, ops


class PgdIterationGraph:
    """step(xL, xR, cleanL, cleanR, disp) runs ONE PGD iteration in place on xL/xR ([1,3,H,W]
    normalised images; clean* are the denormalised clean copies) and returns the loss (device
    scalar, valid until the next step)."""

    def __init__(self, model, cfg, labels, calib, alpha, eps, example, norm='linf', use_graph=True, warmup=2,
                 lanes=1):
        """``lanes`` > 1 captures that many independent pair-iterations on parallel streams inside
        ONE graph: the launch/latency-bound 2-D sections of one pair overlap the tensor-core
        sections of the other (``step_multi``)."""
        self.model, self.cfg, self.labels, self.calib = model, cfg, labels, calib
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
                self.losses = []
                for st, bufs in zip(self.streams, self.sl):
                    st.wait_stream(cap)
                    with torch.cuda.stream(st):
                        self.losses.append(self._iteration(*bufs))
                for st in self.streams:
                    cap.wait_stream(st)
        self.loss = self.losses[0]
        self.launches_per_step = (ops.LAUNCH_COUNT - n0) // self.lanes
        for bufs in self.sl:
            for dst, src in zip(bufs, (xL, xR, cL, cR, disp)):
                dst.copy_(src)

    def _iteration(self, xL, xR, cL, cR, disp):
        a, b = xL.detach().requires_grad_(True), xR.detach().requires_grad_(True)
        out = self.model(a, b, self.calib[0], self.calib[1], self.calib[2], calibs_Proj_R=self.calib[3])
        loss = dsgn.attack_loss(self.cfg, out, disp, self.labels)
        gL, gR = torch.autograd.grad(loss, [a, b])
        if self.norm == 'linf':
            attack.pgd_step_pair(xL, gL.contiguous(), cL, xR, gR.contiguous(), cR, self.alpha, self.eps, inplace=True)
        else:
            attack.pgd_step(xL, gL.contiguous(), cL, self.alpha, self.eps, norm=self.norm, out=xL)
            attack.pgd_step(xR, gR.contiguous(), cR, self.alpha, self.eps, norm=self.norm, out=xR)
        return loss.detach()

    def step_multi(self, pairs):
        """``pairs`` = list of ``lanes`` tuples (xL, xR, cL, cR, disp): one iteration of each, concurrently."""
        assert self.use_graph and len(pairs) == self.lanes
        for bufs, pr in zip(self.sl, pairs):
            for dst, src in zip(bufs, pr):
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        for bufs, pr in zip(self.sl, pairs):
            pr[0].copy_(bufs[0], non_blocking=True)
            pr[1].copy_(bufs[1], non_blocking=True)
        return self.losses

    def step(self, xL, xR, cL, cR, disp):
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
