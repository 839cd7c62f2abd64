#!/usr/bin/env python
"""bench.py -- PGD attack iterations/s on KITTI-shaped stereo pairs (BASELINE.json metric).

Workload at N=1: BASELINE.json configs[1] -- 10-iteration L-inf PGD (eps 0.03, alpha eps/4) on a
batch of 8 synthetic 384x1248 stereo pairs, DSGN-shaped model with seeded random-init weights.
One "step" = one PGD iteration (forward + backward + fused pixel update) applied to every pair
of the rank's batch; value = pair-iterations / s over all ranks (weak scaling: 8 pairs per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

``--impl reference`` times the CPU oracle (oracle/, a restatement of the reference's loop: the
reference itself cannot run, its model lives in the un-vendored DSGN package) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# stdout must carry exactly ONE JSON line, but NCCL / torch print a version banner to the C-level
# stdout during init: park fd 1 on stderr for the whole run and emit the line on the saved fd.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

import torch

EPS, ALPHA = 0.03, 0.03 / 4
H, W = 384, 1248
PAIRS_PER_GPU = 8
METRIC, UNIT = "pgd_attack_pair_iterations_per_second", "pair-iterations/s"
WORKLOAD = ("BASELINE configs[1]: 10-iter L-inf PGD eps=0.03 alpha=eps/4, DSGN-shaped model, "
            "8 synthetic 384x1248 stereo pairs per GPU, random-init weights (seed 1)")


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm_gbs=float(p["hbm_gbs"]), bf16_tflops=float(p["bf16_tflops"]),
                    bf16_tflops_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    except Exception:
        return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def make_batch(cfg, pair_ids, pin=False):
    from eval_driving_safety_b200 import synthetic
    L, R, D = [], [], []
    for i in pair_ids:
        p = synthetic.make_pair(i, H, W)
        L.append(p["imgL"]); R.append(p["imgR"]); D.append(p["disp_L"])
    batch = {"imgL": torch.cat(L), "imgR": torch.cat(R), "disp": torch.cat(D)}
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    return batch


def run_b200(args):
    from eval_driving_safety_b200 import build as b2build
    if not os.path.exists(b2build.LIB) and int(os.environ.get("LOCAL_RANK", "0")) == 0:
        b2build.build()                  # normally the in-tree .so built by __graft_entry__.build() travels with the repo
    from eval_driving_safety_b200 import attack, dsgn, engine, ops, parallel, synthetic
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    rank, world = parallel.init(device=dev)
    torch.backends.cudnn.benchmark = True
    # stock 2-D convolutions of the extractor / BEV head (cuDNN): TF32 like our 3-D tensor-core convs
    # unless --backbone-fp32 (fp32 CUDA-core cuDNN kernels: tighter parity, slower)
    torch.backends.cudnn.allow_tf32 = not args.backbone_fp32
    torch.backends.cuda.matmul.allow_tf32 = not args.backbone_fp32
    cfg = dsgn.default_cfg()
    model = dsgn.build_model(cfg, seed=1, device=dev)
    calib = synthetic.make_calib(1)
    pair_ids = [rank * PAIRS_PER_GPU + j for j in range(PAIRS_PER_GPU)]
    host = make_batch(cfg, pair_ids, pin=True)
    labels = {k: v.to(dev) for k, v in synthetic.make_labels(cfg, 1, 7).items()}
    mean = torch.tensor(attack.IMAGENET_MEAN, device=dev).view(1, 3, 1, 1)
    std = torch.tensor(attack.IMAGENET_STD, device=dev).view(1, 3, 1, 1)

    # ---- resident-input arm ("value") ------------------------------------------------------
    xL, xR, disp = host["imgL"].to(dev), host["imgR"].to(dev), host["disp"].to(dev)
    cleanL, cleanR = xL * std + mean, xR * std + mean
    example = (xL[0:1], xR[0:1], cleanL[0:1], cleanR[0:1], disp[0:1])
    # the pair-iteration is captured once into a CUDA graph and replayed per pair (engine.py)
    lanes = 1 if args.eager else args.lanes
    eng = engine.PgdIterationGraph(model, cfg, labels, calib, ALPHA, EPS, example, use_graph=not args.eager,
                                   lanes=lanes)

    def iteration(xL, xR, cleanL, cleanR, disp, eng=eng):
        """one PGD iteration of every pair of the batch, in place on xL/xR; returns summed loss"""
        total = torch.zeros((), device=dev)
        n = xL.shape[0]
        sl = lambda j: (xL[j:j + 1], xR[j:j + 1], cleanL[j:j + 1], cleanR[j:j + 1], disp[j:j + 1])
        j = 0
        while eng.lanes > 1 and j + eng.lanes <= n:
            for l in eng.step_multi([sl(j + k) for k in range(eng.lanes)]):
                total += l
            j += eng.lanes
        for j in range(j, n):
            total += eng.step(*sl(j))
        return total
    for _ in range(args.warmup):
        iteration(xL, xR, cleanL, cleanR, disp)
    torch.cuda.synchronize()
    parallel.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        iteration(xL, xR, cleanL, cleanR, disp)
    e1.record()
    torch.cuda.synchronize()
    parallel.barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launches_per_step * PAIRS_PER_GPU * args.steps * world
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end arm ("e2e"): host buffers, H2D + D2H inside every timed step ------------
    hL, hR = host["imgL"].clone().pin_memory(), host["imgR"].clone().pin_memory()
    h_clean_L = (host["imgL"] * std.cpu() + mean.cpu()).pin_memory()
    h_clean_R = (host["imgR"] * std.cpu() + mean.cpu()).pin_memory()
    h_loss = torch.zeros((), pin_memory=True)

    # Every step moves ALL of its inputs host -> device and its results device -> host; the copies of pair group g+1 ride
    # on a copy stream while group g computes (the user-facing loop would do the same: nothing is kept resident across steps)
    copy_stream = torch.cuda.Stream()
    dL, dR, dcL, dcR, dd = (torch.empty_like(t, device=dev) for t in (hL, hR, h_clean_L, h_clean_R, host["disp"]))
    group = eng.lanes if (eng.lanes > 1 and not args.eager) else 1

    def e2e_step():
        main = torch.cuda.current_stream()
        copy_stream.wait_stream(main)
        n = hL.shape[0]
        ready = []
        with torch.cuda.stream(copy_stream):
            for j in range(0, n, group):
                sl = slice(j, j + group)
                for dst, src in ((dL, hL), (dR, hR), (dcL, h_clean_L), (dcR, h_clean_R), (dd, host["disp"])):
                    dst[sl].copy_(src[sl], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                ready.append(ev)
        total = torch.zeros((), device=dev)
        done = []
        for gi, j in enumerate(range(0, n, group)):
            sl = slice(j, j + group)
            main.wait_event(ready[gi])
            total += iteration(dL[sl], dR[sl], dcL[sl], dcR[sl], dd[sl])
            ev = torch.cuda.Event()
            ev.record(main)
            done.append((sl, ev))
        with torch.cuda.stream(copy_stream):
            for sl, ev in done:
                copy_stream.wait_event(ev)
                hL[sl].copy_(dL[sl], non_blocking=True)
                hR[sl].copy_(dR[sl], non_blocking=True)
        h_loss.copy_(total, non_blocking=True)
        main.synchronize()
        copy_stream.synchronize()

    e2e_warm = min(args.warmup, 3)
    for _ in range(e2e_warm):
        e2e_step()
    parallel.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    h2d = sum(t.numel() * 4 for t in (hL, hR, h_clean_L, h_clean_R, host["disp"]))
    d2h = (hL.numel() + hR.numel()) * 4 + 4

    # ---- live kernel timing: CUDA events cannot bracket kernels inside a replayed graph, so the
    # same pair-iteration is run eagerly right after the timed region with an event pair around
    # every libb2attack launch (on the launching stream) ---------------------------------------
    prof_pairs = 2
    eager = engine.PgdIterationGraph(model, cfg, labels, calib, ALPHA, EPS, example, use_graph=False)
    pp = [(xL[j:j + 1], xR[j:j + 1], cleanL[j:j + 1], cleanR[j:j + 1], disp[j:j + 1]) for j in range(prof_pairs)]
    eager.iterate_eager(pp)                                                       # warm the eager path
    with ops.profile() as prof:
        eager.iterate_eager(pp)            # same launch shapes as the graph: per-pair kernels + ONE joint pixel update
    kern = prof.summary()

    # ---- per-pair statistics: the only collective of the path -------------------
    rows = [parallel.pair_stats(pair_ids[j], [torch.zeros((), device=dev), torch.zeros((), device=dev)],
                                xL[j:j + 1] * std + mean, cleanL[j:j + 1]) for j in range(PAIRS_PER_GPU)]
    stats = parallel.gather_stats(rows, PAIRS_PER_GPU * world)

    # max over ranks, device-timed
    t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, e2e_ms = t.tolist()
    if rank != 0:
        return
    units = PAIRS_PER_GPU * world * args.steps
    peaks = measured_peaks()
    conv = kern.get("conv3d_tcgen05", dict(calls=0, ms=0.0, work=0, per_s=0.0))
    # TF32 roofline: MEASURED cuBLAS TF32 rate of this GPU class (tools/measure_tf32_peak.py -> profiles/r2_tf32_peak.json,
    # 8192^3, sustained = back to back for 4 s, like the driver's bf16 figure); if that file is absent, half the driver's
    # measured bf16 rate (kind::tf32 issues at half the bf16 rate -- nominal, cuBLAS reaches less: 605 vs 691 TFLOP/s)
    tf32_peak, tf32_src = peaks["bf16_tflops_sustained"] / 2.0, "%s bf16_tflops_sustained / 2" % peaks["source"]
    try:
        with open(os.path.join(ROOT, "profiles", "r2_tf32_peak.json")) as f:
            tf32_peak = float(json.load(f)["tf32_tflops_sustained"])
        tf32_src = "profiles/r2_tf32_peak.json tf32_tflops_sustained (cuBLAS TF32 8192^3 back to back, measured on this pool's B200)"
    except Exception:
        pass
    kernels = {}
    for name, k in sorted(kern.items()):
        if name.startswith("conv3d_tcgen05") or name.startswith("conv3d_simt") or name.startswith("conv2d_tcgen05"):
            kernels[name] = {"calls": k["calls"], "ms_total": round(k["ms"], 3), "tflops": round(k["per_s"] / 1e12, 2)}
        else:
            kernels[name] = {"calls": k["calls"], "ms_total": round(k["ms"], 3), "gbs": round(k["per_s"] / 1e9, 1),
                             "frac_hbm": round(k["per_s"] / 1e9 / peaks["hbm_gbs"], 3)}
    line = {
        "metric": METRIC, "value": units / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "tf32 (fp32 storage, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "pairs_per_gpu": PAIRS_PER_GPU, "image": [H, W], "psv": [64, 48, 96, 312],
                   "voxels": [96, 192, 20, 304], "parallelism": "dp%d (pairs sharded, no data-path collective)" % world,
                   "backbone_2d": {"b2": "own tcgen05 kernels, %s" % (
                       ("forward: 3xTF32 error-compensated split in-kernel (fp32-class); data gradient: %s"
                        % ("3xTF32" if ops.CONV2D_SPLIT_BWD else "plain TF32, like the 3-D convs")) if ops.CONV2D_SPLIT else "plain TF32"),
                                   "cudnn": "cuDNN %s (A/B mode)" % ("fp32" if args.backbone_fp32 else "TF32"),
                                   "cudnn_tf32x3": "cuDNN TF32 x3 stacked on the host (A/B mode)"}[dsgn.BACKBONE_IMPL],
                   "execution": "eager" if args.eager else
                   "CUDA graph of %d concurrent pair-iteration(s) on parallel streams, replayed" % lanes,
                   "parity_of_this_mode": "tests/test_gpu_fullsize.py on this exact engine: per-iteration gradient-sign agreement mean "
                                          "99.942 % / min 99.893 %, updated pixels identical to the CPU oracle mean 98.33 % / min 98.14 % "
                                          "(10 iterations x 2 full-size pairs); cost volume bit-exact",
                   "l2_policy": "per-iteration working set (~10 GB of activations per pair) is far larger than the 126 MB L2"},
        "e2e": {"value": units / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps, "warmup": e2e_warm},
        "gpu_launches": launches,
        "roofline": {"kernel": "conv3d_tcgen05_kernel (all 3x3x3 conv/deconv fwd+dgrad launches of the timed region)",
                     "bound": "tensor", "achieved": conv["per_s"] / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                     "frac": conv["per_s"] / 1e12 / tf32_peak if tf32_peak else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant shape (64->64 at
                     # 48x96x312, algorithmic 736 MB) from the committed ncu --set full capture; `achieved` averages
                     # all conv launches of the step, whose shapes differ
                     "traffic": 709.2e6,
                     "traffic_source": "profiles/r1_conv3d_final_ncu.txt (conv3d_s1n_tcgen05_kernel, 64->64 48x96x312: "
                                       "389.3 MB read + 319.8 MB written vs 736 MB algorithmic)",
                     "peak_source": tf32_src + "; kernel timed inside a long step",
                     "frac_of_half_bf16_sustained": conv["per_s"] / 1e12 / (peaks["bf16_tflops_sustained"] / 2.0),
                     "frac_of_bf16_peak": conv["per_s"] / 1e12 / peaks["bf16_tflops_sustained"],
                     "launches": conv["calls"], "avg_launch_ms": conv["ms"] / max(conv["calls"], 1),
                     "note": "conv flops only: about a third of the launches also add up GroupNorm statistics / backward "
                             "sums or a second gradient in their epilogue (work of other kernels, not counted here); "
                             "B2_FUSE_GN_STATS=0 B2_FUSE_GRAD_ADD=0 B2_FUSE_GN_BWD=0 gives the plain-conv figure (607-612)",
                     "share_of_step": (conv["ms"] / prof_pairs) / (ms / (args.steps * PAIRS_PER_GPU)) if ms else None,
                     "measured_in": "eager pass of %d pair-iterations right after the timed region, CUDA event pair "
                                    "around every launch on the launching stream (events cannot bracket kernels "
                                    "inside a replayed CUDA graph); a device-side spin ahead of each pair keeps the "
                                    "host's launch latency out of the bracket" % prof_pairs},
        "kernels": kernels,
        "clocks": clocks,
        "stats_rows_gathered": int(stats.shape[0]),
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_sample()
    emit(line)


# ---------------------------------------------------------------------------------------------
# BASELINE configs[3] / configs[4]: the other two attack loops, measured the same way (profiles/ only: the
# driver's BENCH / SCALE records are the default config)
# ---------------------------------------------------------------------------------------------
def _timed(fn, steps, warmup, dev, world):
    """(ms of the timed steps: CUDA events, max over ranks; this rank's clocks sampled during them)"""
    from eval_driving_safety_b200 import parallel
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    parallel.barrier()
    sampler = ClockSampler(dev.index)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    parallel.barrier()
    clocks = sampler.stop()
    own = e0.elapsed_time(e1)
    t = torch.tensor([own], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        lo = torch.tensor([own], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
        clocks["ms_fastest_rank"] = lo.item() / steps
    return t.item(), clocks


def run_patch(args):
    """configs[3]: universal adversarial patch (ratio 0.2 -> 77x77, alpha 1e3, eps 8/255, fake ground truth), every
    rank attacks its own image of the step, the clipped patch step is all-reduced over NCCL INSIDE the captured
    iteration.  step = one patch iteration per rank; a new image (H2D from pinned memory in the e2e arm) every 2."""
    import random
    from eval_driving_safety_b200 import attack, dsgn, engine, ops, parallel, synthetic
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    rank, world = parallel.init(device=dev)
    cfg = dsgn.default_cfg()
    model = dsgn.build_model(cfg, seed=1, device=dev)
    calib = synthetic.make_calib(1)
    bbox, box3d = synthetic.make_targets(6, 1)
    attack.inject_fake_gt(bbox, box3d)
    labels = synthetic.labels_from_box3d(cfg, box3d, device=dev)
    dim, radius = attack.patch_dim_radius(H, 0.2)
    patch = torch.zeros(1, 3, dim, dim, device=dev)
    imgs = [synthetic.make_pair(rank * 4 + j, H, W) for j in range(4)]
    host = [{k: v.pin_memory() for k, v in p.items()} for p in imgs]
    devp = [{k: v.to(dev) for k, v in p.items()} for p in imgs]
    rng = random.Random(1)                                                # seeded centres
    centres = [attack.generate_round_mask(radius, rng, H, W) for _ in range(64)]
    hook = parallel.allreduce_patch_delta if world > 1 else None
    eng = engine.PatchIterationGraph(model, cfg, labels, calib, patch, radius, (devp[0]["imgL"], devp[0]["imgR"], devp[0]["disp_L"]),
                                     alpha=1e3, eps=8 / 255, allreduce=hook)
    state = {"k": 0}

    def step(src):
        k = state["k"]
        if k % 2 == 0:                                                    # iters = 2 per image (patch_attack.py:53)
            p = src[(k // 2) % len(src)]
            cl, cr = centres[(k // 2) % len(centres)]
            eng.load(p["imgL"], p["imgR"], p["disp_L"], cl, cr)
        state["k"] = k + 1
        return eng.iterate()

    ms, clocks = _timed(lambda: step(devp), args.steps, args.warmup, dev, world)
    h_loss = torch.zeros((), pin_memory=True)

    def e2e():
        h_loss.copy_(step(host), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    parallel.barrier()
    for _ in range(2):
        e2e()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    e2e_ms = t.item()
    if rank != 0:
        return
    units = world * args.steps
    emit({"metric": "patch_attack_pair_iterations_per_second", "value": units / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
          "scaling": "weak", "vs_baseline": None, "dtype": "tf32 (fp32 storage, fp32 accumulate); 2-D convs 3xTF32", "data": "synthetic",
          "config": {"workload": "BASELINE configs[3]: DSGN universal adversarial patch (77x77 circular patch, alpha 1e3, eps 8/255, "
                                 "fake ground truth), one image per GPU per iteration, 2 iterations per image, clipped patch step "
                                 "all-reduced (NCCL, 71 KB) inside the captured iteration",
                     "execution": "CUDA graph of one patch iteration incl. the all_reduce; patch centres in device memory",
                     "parallelism": "dp%d, synchronous mini-batch of %d images per patch step" % (world, world)},
          "e2e": {"value": units / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                  "h2d_bytes_per_step": (2 * 3 * H * W + H * W) * 4 // 2, "d2h_bytes_per_step": 4},
          "gpu_launches": (eng.launches_per_step or 0) * units, "clocks": clocks, "patch_absmax": patch.abs().max().item()})


def run_srcnn(args):
    """configs[4]: Stereo R-CNN PGD (RoIAlign fwd/bwd path) on 600x1987 pairs in mean-subtracted 0-255 space,
    R = 256 RoIs per view over FPN levels 2-5, alpha 1.0, eps 0.03*255; step = one PGD iteration of one pair per rank."""
    from eval_driving_safety_b200 import attack, ops, parallel, stereo_rcnn as S
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    rank, world = parallel.init(device=dev)
    model = S.SyntheticStereoRCNN(width=256, seed=1).to(dev)
    il, ir = S.synthetic_pair(rank, 600, 1987)
    rl, rr = S.synthetic_rois(256, 600, 1987, seed=rank)
    tg = {k: v.to(dev) for k, v in S.synthetic_targets(256, seed=rank).items()}
    hl, hr = il.pin_memory(), ir.pin_memory()
    rl, rr = rl.to(dev), rr.to(dev)
    xl, xr, cl, cr = il.to(dev), ir.to(dev), il.to(dev), ir.to(dev)
    eps255 = 255 * 0.03
    lo = [0 - m for m in attack.STEREO_RCNN_MEANS]
    hi = [255 - m for m in attack.STEREO_RCNN_MEANS]

    def iteration():
        """one PGD iteration in place on xl / xr (attack/Stereo-RCNN/pgd_attack.py:151-217)"""
        a, b = xl.detach().requires_grad_(True), xr.detach().requires_grad_(True)
        loss = model(a, b, rl, rr, tg)
        gl, gr = torch.autograd.grad(loss, [a, b])
        attack.pgd_step(xl, gl.contiguous(), cl, 1.0, eps255, mean=None, std=None, lo=lo, hi=hi, out=xl)
        attack.pgd_step(xr, gr.contiguous(), cr, 1.0, eps255, mean=None, std=None, lo=lo, hi=hi, out=xr)
        return loss.detach()

    # the one-launch RoIAlign dispatch has no host synchronisation, so the whole iteration is graph-capturable
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            iteration()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    n0 = ops.LAUNCH_COUNT
    graph = torch.cuda.CUDAGraph()
    if args.eager:
        loss_buf = None
    else:
        with torch.cuda.graph(graph):
            loss_buf = iteration()
    launches = ops.LAUNCH_COUNT - n0 if not args.eager else 8
    xl.copy_(cl); xr.copy_(cr)

    def step():
        if args.eager:
            return iteration()
        graph.replay()
        return loss_buf
    ms, clocks = _timed(step, args.steps, args.warmup, dev, world)
    h_loss = torch.zeros((), pin_memory=True)

    def e2e():
        xl.copy_(hl, non_blocking=True); xr.copy_(hr, non_blocking=True)
        h_loss.copy_(step(), non_blocking=True)
        hl.copy_(xl, non_blocking=True); hr.copy_(xr, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for _ in range(2):
        e2e()
    parallel.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e()
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    e2e_ms = t.item()
    if rank != 0:
        return
    units = world * args.steps
    emit({"metric": "stereo_rcnn_pgd_pair_iterations_per_second", "value": units / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
          "scaling": "weak", "vs_baseline": None, "dtype": "f32 (stock FPN stand-in; own RoIAlign fwd / deterministic bwd, own update kernel)",
          "data": "synthetic",
          "config": {"workload": "BASELINE configs[4]: Stereo R-CNN PGD (RoIAlign fwd/bwd path), 600x1987 pairs, 256 RoIs per view, "
                                 "alpha 1.0, eps 0.03*255; Stereo-R-CNN-shaped stand-in network (the real detector is un-vendored)",
                     "parallelism": "dp%d (pairs sharded, no data-path collective)" % world,
                     "execution": "eager" if args.eager else "CUDA graph of one PGD iteration, replayed"},
          "e2e": {"value": units / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                  "h2d_bytes_per_step": 2 * 3 * 600 * 1987 * 4, "d2h_bytes_per_step": 2 * 3 * 600 * 1987 * 4 + 4},
          "gpu_launches": launches * units, "clocks": clocks})


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle (the reference's own model lives in the un-vendored DSGN package and cannot run;
# oracle/ restates the loop with stock torch ops) at FULL SIZE -- the same 384x1248 pair, voxel grid and
# network as the GPU arm.  A "step" of this arm = one PGD iteration of ONE full-size pair (the bounded
# sample of the 8-pair GPU step: pairs are independent, so pair-iterations/s needs no scaling factor).
# ---------------------------------------------------------------------------------------------
CPU_ARM_BUDGET_S = 900.0     # the timed loop stops early (and says so) rather than run for an hour on a slow host


def cpu_iteration(model, cfg, pair, calib, labels):
    from oracle import attack_ref, dsgn_ref
    xL, xR = pair["imgL"].clone().requires_grad_(True), pair["imgR"].clone().requires_grad_(True)
    out = model(xL, xR, *calib[:3], calibs_Proj_R=calib[3])
    loss = dsgn_ref.attack_loss(cfg, out, pair["disp_L"], labels)
    gL, gR = torch.autograd.grad(loss, [xL, xR])
    pair["imgL"] = attack_ref.pgd_step_linf(xL.detach(), gL, pair["cleanL"], ALPHA, EPS)
    pair["imgR"] = attack_ref.pgd_step_linf(xR.detach(), gR, pair["cleanR"], ALPHA, EPS)
    return loss.item()


def cpu_setup():
    from eval_driving_safety_b200 import synthetic
    from oracle import attack_ref, dsgn_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = dsgn_ref.default_cfg()
    model = dsgn_ref.build_model(cfg, seed=1)
    for p in model.parameters():
        p.requires_grad_(False)      # like our arm: no wgrad (the reference wastes it, pgd_attack.py:333)
    pair = synthetic.make_pair(0, H, W)
    pair["cleanL"], pair["cleanR"] = attack_ref.denormalize(pair["imgL"]), attack_ref.denormalize(pair["imgR"])
    calib = synthetic.make_calib(1)
    labels = dsgn_ref.make_labels(cfg, 1, 7)
    return model, cfg, pair, calib, labels, cores


CPU_SAMPLE = ("oracle (pure-PyTorch fp32 CPU restatement of the reference loop), full size: each step = 1 PGD iteration "
              "(forward + backward to the pixels + update) of one 384x1248 pair, same network / voxel grid / eps / alpha "
              "as the GPU arm; no crop, no scaling")


def cpu_baseline_sample():
    model, cfg, pair, calib, labels, cores = cpu_setup()
    t0 = time.perf_counter()
    cpu_iteration(model, cfg, pair, calib, labels)
    dt = time.perf_counter() - t0
    return {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": CPU_SAMPLE + "; 1 iteration, %.1f s (first call, includes thread-pool / allocator warm-up)" % dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model, cfg, pair, calib, labels, cores = cpu_setup()
    t_w = time.perf_counter()
    for _ in range(min(args.warmup, 1)):          # one full-size iteration warms the thread pool and the allocator
        cpu_iteration(model, cfg, pair, calib, labels)
    t_w = time.perf_counter() - t_w
    done, t0 = 0, time.perf_counter()
    while done < args.steps:
        cpu_iteration(model, cfg, pair, calib, labels)
        done += 1
        if time.perf_counter() - t0 + t_w > CPU_ARM_BUDGET_S:
            break
    dt = time.perf_counter() - t0
    value = done / dt
    sample = CPU_SAMPLE + "; %d timed iteration(s) after %d warm-up" % (done, min(args.warmup, 1))
    if done < args.steps:
        sample += " (stopped after %d of %d requested steps: %.0f s wall budget)" % (done, args.steps, CPU_ARM_BUDGET_S)
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "steps_requested": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / done * 1e3,
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_gpu": PAIRS_PER_GPU, "image": [H, W], "psv": [64, 48, 96, 312],
                   "voxels": [96, 192, 20, 304],
                   "cpu_arm": "CPU oracle (the reference's model is in the un-vendored DSGN package and cannot run); "
                              "one pair per step -- pairs are independent, the metric is pair-iterations/s"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="no CUDA graph (debug)")
    ap.add_argument("--backbone-fp32", action="store_true", help="cuDNN 2-D convs in fp32 instead of TF32")
    ap.add_argument("--lanes", type=int, default=2, help="pair-iterations captured side by side in one graph")
    ap.add_argument("--config", default="pgd", choices=["pgd", "patch", "srcnn"],
                    help="pgd = BASELINE configs[1]/[2] (the headline, default); patch = configs[3]; srcnn = configs[4]")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "patch":
        run_patch(args)
    elif args.config == "srcnn":
        run_srcnn(args)
    else:
        run_b200(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
