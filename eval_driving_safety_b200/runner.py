"""Runnable mirrors of the reference's attack scripts on synthetic KITTI-shaped data.

  python -m eval_driving_safety_b200.runner pgd   --iter 10 --alpha 0.0075 --eps 0.03 --pairs 8
  python -m eval_driving_safety_b200.runner patch --ratio 0.2 --epochs 1 --iter 2 --pairs 8
  python -m eval_driving_safety_b200.runner srcnn-patch --ratio 0.1 --epochs 1 --iter 2 --pairs 8
  torchrun --nproc-per-node 8 -m eval_driving_safety_b200.runner pgd --pairs 64 --iter 20

Flags follow attack/DSGN/pgd_attack.py:53-55 (--iter/--alpha/--eps) and
attack/DSGN/patch_attack.py:53-56 (--iter/--eps/--epochs/--ratio).  Pairs are sharded over the
ranks (pair i -> rank i mod world); per-pair statistics are all-gathered at the end;
the universal patch is kept identical on all ranks by all-reducing the clipped step.
``--save-dir`` writes the per-iteration images in the reference's hand-off layout
``<dir>/dsgn_pgd_iters_{k}/image_{2,3}/%06d.png`` (attack/DSGN/pgd_attack.py:357-374) with the
reference's tensor2im convention (denormalise, *255, truncating uint8 cast, :157-178).
"""
import argparse
import os
import random

import numpy as np
import torch

from . import attack, dsgn, engine, kitti_io, parallel, synthetic


tensor2im = kitti_io.tensor2im        # attack/DSGN/pgd_attack.py:157-178 (truncating uint8 cast)


def save_pair(save_dir, k, index, imgL, imgR, w, h, writer=None):
    """Per-iteration hand-off images, attack/DSGN/pgd_attack.py:357-374.  With a ``writer``
    (kitti_io.AsyncImageWriter) the copy to the host and the PNG encoding happen off the attack loop."""
    for path, img in zip(kitti_io.iteration_paths(save_dir, k, index), (imgL, imgR)):
        if writer is not None:
            writer.submit(img[0], path, w, h)
        else:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            kitti_io.save_image(img[0], path, w, h)


def _setup(args):
    rank, world = parallel.init()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = dsgn.tiny_cfg() if args.tiny else dsgn.default_cfg()
    h, w = (32, 64) if args.tiny else (384, 1248)
    model = dsgn.build_model(cfg, seed=args.seed, device=dev)
    calib = synthetic.make_calib(1, scale=h / 384, cu=w / 2, cv=h / 2) if args.tiny else synthetic.make_calib(1)
    labels = {k: v.to(dev) for k, v in synthetic.make_labels(cfg, 1, 7).items()}
    return rank, world, dev, cfg, (h, w), model, calib, labels


def _write_detections(args, model, cfg, calib, index, xL, xR, tag):
    """Detections of the (clean / attacked) pair in the reference's KITTI hand-off format
    (attack/DSGN/predict_and_save_pgd.py:249-283) under ``<detections>/<tag>/%06d.txt``; returns their number."""
    with torch.no_grad():
        out = model(xL, xR, calib[0], calib[1], calib[2], calibs_Proj_R=calib[3])
    dets = dsgn.decode_detections(cfg, out, calib[2][0], score_thresh=args.score_thresh)
    kitti_io.write_detections(os.path.join(args.detections, tag), index, dets)
    return len(dets)


def run_pgd(args):
    rank, world, dev, cfg, (h, w), model, calib, labels = _setup(args)
    mean = torch.tensor(attack.IMAGENET_MEAN, device=dev).view(1, 3, 1, 1)
    std = torch.tensor(attack.IMAGENET_STD, device=dev).view(1, 3, 1, 1)
    rows, eng = [], None
    writer = kitti_io.AsyncImageWriter() if (args.save_dir and not args.sync_save) else None
    for i in parallel.shard_pairs(args.pairs, rank, world):
        p = synthetic.make_pair(i, h, w, max_depth=cfg.max_depth)
        xL, xR, disp = p["imgL"].to(dev), p["imgR"].to(dev), p["disp_L"].to(dev)
        cL, cR = xL * std + mean, xR * std + mean
        if eng is None:
            eng = engine.PgdIterationGraph(model, cfg, labels, calib, args.alpha, args.eps, (xL, xR, cL, cR, disp),
                                           norm=args.norm, use_graph=not args.eager)
        if args.save_dir:
            save_pair(args.save_dir, 0, i, xL, xR, w, h, writer)
        n_clean = n_adv = float('nan')
        if args.detections:
            n_clean = _write_detections(args, model, cfg, calib, i, xL, xR, "clean")
        losses = []
        for k in range(args.iter):
            losses.append(eng.step(xL, xR, cL, cR, disp, calib=calib).clone())
            if args.save_dir:
                save_pair(args.save_dir, k + 1, i, xL, xR, w, h, writer)
        if args.detections:
            n_adv = _write_detections(args, model, cfg, calib, i, xL, xR, "adv")
        rows.append(parallel.pair_stats(i, losses, xL * std + mean, cL, n_clean, n_adv))
    if writer is not None:
        writer.close()
    stats = parallel.gather_stats(rows, args.pairs)
    if rank == 0 and getattr(args, "stats_out", None):
        torch.save(stats.cpu(), args.stats_out)
    if rank == 0:
        print("pair  loss_0  loss_K  linf  l2  frac_changed")
        for r in stats.tolist():
            print("%4d  %.4f  %.4f  %.4f  %.3f  %.3f" % (int(r[0]), r[1], r[2], r[3], r[4], r[5]))
    return stats


def run_patch(args):
    """attack/DSGN/patch_attack.py:278-443 with seeded centres: every image is given the fake ground truth
    (one car at 29 m, :336-354) and the patch DESCENDS the loss towards it; the patch resumes from
    ``<save-dir>/epoch0/patch.npy`` (:211-234); ranks process different images of the same step and all-reduce
    the clipped patch step (world = 1 == the reference's sequence)."""
    rank, world, dev, cfg, (h, w), model, calib, _ = _setup(args)
    if args.save_dir:
        if rank == 0:
            dim, radius, patch0 = attack.init_patch(args.ratio, args.save_dir, short_side=h)      # resume / create
        parallel.barrier()
        if rank != 0:
            dim, radius, patch0 = attack.init_patch(args.ratio, args.save_dir, short_side=h)
        patch = torch.from_numpy(patch0).to(dev).contiguous()
    else:
        dim, radius = attack.patch_dim_radius(h, args.ratio)
        patch = torch.zeros(1, 3, dim, dim, device=dev)                 # init_patch: zeros (:229)
    # the fake ground truth is the same for every image (:336-354): all real boxes zeroed, box 0 = the fake car
    bbox, box3d = synthetic.make_targets(6, args.seed)
    attack.inject_fake_gt(bbox, box3d)
    if args.tiny:                                                       # shrunk world grid: move the car inside it
        box3d[0, 3:6] = torch.tensor([0.3, 0.5, 5.1])
    labels = synthetic.labels_from_box3d(cfg, box3d, device=dev)
    rng = random.Random(args.seed)
    alpha, losses_all = 1e3, []                                          # :279
    steps = (args.pairs + world - 1) // world
    eng = None
    hook = parallel.allreduce_patch_delta if world > 1 else None
    for epoch in range(args.epochs):
        for s in range(steps):
            centres = [attack.generate_round_mask(radius, rng, h, w) for _ in range(world)]
            i = (s * world + rank) % args.pairs
            cl, cr = centres[rank]
            if cr[1] - radius < 0:                                       # tiny frames: keep the right box inside
                cr = [cr[0], radius]
            p = synthetic.make_pair(i, h, w, max_depth=cfg.max_depth)
            xL, xR, disp = p["imgL"].to(dev), p["imgR"].to(dev), p["disp_L"].to(dev)
            if eng is None:
                eng = engine.PatchIterationGraph(model, cfg, labels, calib, patch, radius, (xL, xR, disp), alpha=alpha,
                                                 eps=args.eps, use_graph=not args.eager, allreduce=hook)
            eng.load(xL, xR, disp, cl, cr)
            for _ in range(args.iter):
                loss = eng.iterate()
            losses_all.append(loss.clone())
    if args.save_dir and rank == 0:
        d = os.path.join(args.save_dir, "epoch%d" % args.epochs)
        os.makedirs(d, exist_ok=True)
        np.save(os.path.join(d, "patch.npy"), patch.cpu().numpy())       # :438-443
    if rank == 0:
        print("patch %dx%d  mean loss %.4f  |patch|max %.4f" % (dim, dim, torch.stack(losses_all).mean().item(),
                                                               patch.abs().max().item()))
    return patch


def run_srcnn(args):
    """BASELINE config 5: PGD in 0-255 space against the Stereo-R-CNN-shaped stand-in (RoIAlign fwd/bwd
    path), attack/Stereo-RCNN/pgd_attack.py:151-217; flags :42-48 (--iter, --alpha 1.0, --eps 0.3)."""
    import time
    from . import stereo_rcnn as S
    rank, world = parallel.init()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    h, w, nroi = (96, 320, 24) if args.tiny else (600, 1987, 256)
    model = S.SyntheticStereoRCNN(width=64 if args.tiny else 256, seed=args.seed).to(dev)
    eps255 = 255 * args.eps                                               # pgd_attack.py:57
    done, t0 = 0, None
    for i in parallel.shard_pairs(args.pairs, rank, world):
        il, ir = S.synthetic_pair(i, h, w)
        rl, rr = S.synthetic_rois(nroi, h, w, seed=i)
        tg = {k: v.to(dev) for k, v in S.synthetic_targets(nroi, seed=i).items()}
        if t0 is None:                                                    # first pair = warm-up
            S.pgd_attack(model, il.to(dev), ir.to(dev), rl.to(dev), rr.to(dev), tg, 1, args.alpha, eps255)
            torch.cuda.synchronize(); t0 = time.perf_counter()
        al, ar, losses = S.pgd_attack(model, il.to(dev), ir.to(dev), rl.to(dev), rr.to(dev), tg, args.iter,
                                      args.alpha, eps255)
        done += args.iter
    torch.cuda.synchronize()
    rate = torch.tensor([done / (time.perf_counter() - t0)], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(rate)
    if rank == 0:
        print("stereo-rcnn PGD: %.2f pair-iterations/s on %d GPU(s); last loss %.4f -> %.4f; |delta|max %.3f"
              % (rate.item(), world, losses[0].item(), losses[-1].item(), (al - il.to(dev)).abs().max().item()))
    return rate.item()


def run_srcnn_patch(args):
    """Universal patch against the Stereo-R-CNN-shaped stand-in, attack/Stereo-RCNN/patch_attack.py:100-295:
    patch_dim = int(600 * ratio) made odd (:58-65), seeded centres in the reference's ranges (:79-83), fake
    single-box ground truth = the patch square (:187-207), clipped descent + per-channel clamp (:268-281),
    resume from / save to ``<save-dir>/epoch{0,E}/patch.npy`` (:66-76, :288-295)."""
    from . import stereo_rcnn as S
    rank, world = parallel.init()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    h, w, nroi = (96, 320, 24) if args.tiny else (600, 1987, 256)
    model = S.SyntheticStereoRCNN(width=64 if args.tiny else 256, seed=args.seed).to(dev)
    if args.save_dir:
        if rank == 0:
            dim, radius, patch0 = attack.init_patch(args.ratio, args.save_dir, short_side=h, resize=False)
        parallel.barrier()
        if rank != 0:
            dim, radius, patch0 = attack.init_patch(args.ratio, args.save_dir, short_side=h, resize=False)
        patch = torch.from_numpy(patch0).to(dev).contiguous()
    else:
        dim, radius = attack.patch_dim_radius(h, args.ratio)
        patch = torch.zeros(1, 3, dim, dim, device=dev)
    rng = random.Random(args.seed)
    hook = parallel.allreduce_patch_delta if world > 1 else None
    steps = (args.pairs + world - 1) // world
    last = None
    for epoch in range(args.epochs):
        for s in range(steps):
            centres = [attack.generate_round_mask(radius, rng, h, w) for _ in range(world)]
            i = (s * world + rank) % args.pairs
            cl, cr = centres[rank]
            if cr[1] - radius < 0:
                cr = [cr[0], radius]
            il, ir = S.synthetic_pair(i, h, w)
            rl, rr = S.synthetic_rois(nroi, h, w, seed=i)
            tg = {k: v.to(dev) for k, v in S.synthetic_targets(nroi, seed=i).items()}
            last = S.patch_attack_image(model, il.to(dev), ir.to(dev), rl.to(dev), rr.to(dev), tg, patch, cl, cr, radius,
                                        iters=args.iter, alpha=1e3, eps=args.eps, delta_hook=hook)
    if args.save_dir and rank == 0:
        d = os.path.join(args.save_dir, "epoch%d" % args.epochs)
        os.makedirs(d, exist_ok=True)
        np.save(os.path.join(d, "patch.npy"), patch.cpu().numpy())
    if rank == 0:
        print("stereo-rcnn patch %dx%d  last losses %s  patch range [%.2f, %.2f]"
              % (dim, dim, [round(v, 4) for v in last.tolist()], patch.min().item(), patch.max().item()))
    return patch


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    a = sub.add_parser("pgd")
    a.add_argument("--iter", type=int, default=4)                        # pgd_attack.py:53
    a.add_argument("--alpha", type=float, default=1 / 255)               # :54
    a.add_argument("--eps", type=float, default=0.3)                     # :55
    a.add_argument("--norm", default="linf", choices=["linf", "l2"])
    a.add_argument("--stats-out", default=None, help="rank 0 saves the gathered per-pair statistics [pairs, fields] here")
    a.add_argument("--detections", default=None, help="write clean / attacked detections as KITTI txt files under this directory")
    a.add_argument("--score-thresh", type=float, default=0.5)
    b = sub.add_parser("patch")
    b.add_argument("--iter", type=int, default=2)                        # patch_attack.py:53
    b.add_argument("--eps", type=float, default=8 / 255)                 # :54
    b.add_argument("--epochs", type=int, default=80)                     # :55
    b.add_argument("--ratio", type=float, default=0.2)                   # :56
    c = sub.add_parser("srcnn")
    c.add_argument("--iter", type=int, default=10)                       # Stereo-RCNN/pgd_attack.py:42-48
    c.add_argument("--alpha", type=float, default=1.0)
    c.add_argument("--eps", type=float, default=0.03)
    d = sub.add_parser("srcnn-patch")
    d.add_argument("--iter", type=int, default=2)                        # Stereo-RCNN/patch_attack.py:43-46
    d.add_argument("--eps", type=float, default=0.1)
    d.add_argument("--epochs", type=int, default=40)
    d.add_argument("--ratio", type=float, default=0.1)
    for q in (a, b, c, d):
        q.add_argument("--pairs", type=int, default=8)
        q.add_argument("--seed", type=int, default=1)                    # :41
        q.add_argument("--save-dir", default=None)
        q.add_argument("--tiny", action="store_true", help="32x64 frames, shrunk volumes (tests)")
        q.add_argument("--eager", action="store_true")
        q.add_argument("--sync-save", action="store_true", help="write the PNGs inside the loop like the reference")
    args = ap.parse_args(argv)
    out = {"pgd": run_pgd, "patch": run_patch, "srcnn": run_srcnn, "srcnn-patch": run_srcnn_patch}[args.cmd](args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
    return out


if __name__ == "__main__":
    main()
