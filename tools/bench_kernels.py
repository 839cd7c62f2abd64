"""Per-kernel micro-benchmark at KITTI sizes: CUDA events on the launching stream, 3 warm-ups,
L2 flushed (256 MB write) between timed launches.  Writes JSON for profiles/."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from eval_driving_safety_b200 import attack, dsgn, ops, synthetic

dev = torch.device("cuda", 0)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
flush_buf = torch.empty(64 * 1024 * 1024, device=dev)
g = torch.Generator().manual_seed(0)
R = {}


def timeit(name, fn, work, unit, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush_buf.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    rate = work / (ms * 1e-3)
    if unit == "GB/s":
        R[name] = {"ms": round(ms, 4), "GB/s": round(rate / 1e9, 1), "frac_of_measured_hbm": round(rate / 1e9 / peaks["hbm_gbs"], 3), "algorithmic_MB": round(work / 1e6, 1)}
    else:
        R[name] = {"ms": round(ms, 4), "TFLOP/s": round(rate / 1e12, 1), "frac_of_measured_tf32": round(rate / 1e12 / (peaks["bf16_tflops"] / 2), 3), "GFLOP": round(work / 1e9, 1)}
    print(name, R[name], flush=True)


def rnd(*shape):
    return torch.randn(*shape, generator=g).to(dev)


def cl3(n, c, d, h, w):
    return rnd(n, d, h, w, c).permute(0, 4, 1, 2, 3)


# (4) pixel update: batch of 8 pairs, L+R in one launch
x, gr, cl = rnd(8, 3, 384, 1248), rnd(8, 3, 384, 1248), torch.rand(8, 3, 384, 1248, generator=g).to(dev)
x2, gr2, cl2 = x.clone(), gr.clone(), cl.clone()
timeit("pgd_update L+R x8 pairs", lambda: attack.pgd_step_pair(x, gr, cl, x2, gr2, cl2, 0.0075, 0.03, inplace=True), 16 * 2 * x.numel(), "GB/s")
x1, g1, c1 = x[:1].contiguous(), gr[:1].contiguous(), cl[:1].contiguous()
timeit("pgd_update L+R 1 pair (latency bound)", lambda: attack.pgd_step_pair(x1, g1, c1, x1.clone(), g1, c1, 0.0075, 0.03), 16 * 2 * x1.numel(), "GB/s")
del x, gr, cl, x2, gr2, cl2
# (1) cost volume
cfg = dsgn.default_cfg()
fu, b, P, PR = synthetic.make_calib(1)
shifts = dsgn.plane_shifts(cfg, fu, b).to(dev)
L, Rr = rnd(1, 32, 96, 312), rnd(1, 32, 96, 312)
timeit("cost_volume_fwd", lambda: ops.build_cost_volume(L, Rr, shifts), 4 * (2 * L.numel() + 64 * 48 * 96 * 312), "GB/s")
gc = cl3(1, 64, 48, 96, 312)
Lr, Rq = L.clone().requires_grad_(True), Rr.clone().requires_grad_(True)
cv = ops.build_cost_volume(Lr, Rq, shifts)
timeit("cost_volume_bwd", lambda: torch.autograd.grad(cv, [Lr, Rq], gc, retain_graph=True), 4 * (2 * L.numel() + gc.numel()), "GB/s")
del cv
# (2) lifting
psv = cl3(1, 64, 48, 96, 312).requires_grad_(True)
img = rnd(1, 96, 312, 32).permute(0, 3, 1, 2).requires_grad_(True)
grid3 = dsgn.lifting_grid(cfg, P, (96, 312)).to(dev).contiguous()
plan3 = ops.GridPlan(grid3, (48, 96, 312), True)
grid2 = grid3[..., :2].contiguous().view(1, 192 * 20, 304, 2)
plan2 = ops.GridPlan(grid2, (96, 312), True)
nv = 192 * 20 * 304
timeit("grid_sample3d_fwd", lambda: ops.grid_sample(psv, grid3, True, plan3), 4 * (psv.numel() + grid3.numel() + nv * 64), "GB/s")
out3 = ops.grid_sample(psv, grid3, True, plan3); go3 = torch.randn_like(out3)
timeit("grid_sample3d_bwd (CSR gather)", lambda: torch.autograd.grad(out3, psv, go3, retain_graph=True), 4 * (nv * 64 + psv.numel()) + 8 * plan3.nnz + 4 * plan3.ncell, "GB/s")
timeit("grid_sample2d_fwd", lambda: ops.grid_sample(img, grid2, True, plan2), 4 * (img.numel() + grid2.numel() + nv * 32), "GB/s")
out2 = ops.grid_sample(img, grid2, True, plan2); go2 = torch.randn_like(out2)
timeit("grid_sample2d_bwd (CSR gather)", lambda: torch.autograd.grad(out2, img, go2, retain_graph=True), 4 * (nv * 32 + img.numel()) + 8 * plan2.nnz + 4 * plan2.ncell, "GB/s")
del out3, go3, out2, go2
# GroupNorm
xg = cl3(1, 64, 48, 96, 312).requires_grad_(True)
gam, bet = torch.ones(64, device=dev), torch.zeros(64, device=dev)
timeit("groupnorm+relu fwd (3 launches)", lambda: ops.groupnorm_act(xg, gam, bet, 32, 1e-5, relu=True), 4 * xg.numel() * 3, "GB/s")
yg = ops.groupnorm_act(xg, gam, bet, 32, 1e-5, relu=True); gyg = torch.randn_like(yg)
timeit("groupnorm+relu bwd (3 launches)", lambda: torch.autograd.grad(yg, xg, gyg, retain_graph=True), 4 * xg.numel() * 5, "GB/s")
del yg, gyg
# depth head, bev pool, c1
c1t = rnd(1, 1, 48, 96, 312).requires_grad_(True)
timeit("depth_head_fwd", lambda: ops.depth_head(c1t, (192, 384, 1248), 2.0, 0.2), 4 * (c1t.numel() + 384 * 1248), "GB/s")
dh = ops.depth_head(c1t, (192, 384, 1248), 2.0, 0.2); gdh = torch.randn_like(dh)
timeit("depth_head_bwd", lambda: torch.autograd.grad(dh, c1t, gdh, retain_graph=True), 4 * (2 * c1t.numel() + 384 * 1248) + 2 * 4 * 48 * 384 * 1248, "GB/s")
vv = cl3(1, 64, 192, 20, 304).requires_grad_(True)
timeit("bev_pool_fwd", lambda: ops.bev_pool(vv, 4), 4 * (vv.numel() + vv.numel() // 4), "GB/s")
w1 = rnd(1, 64, 3, 3, 3)
timeit("conv3d_c1_fwd (64->1)", lambda: ops.conv3d_c1(xg, w1), 4 * (xg.numel() + 48 * 96 * 312), "GB/s")
# conv layers
for name, cin, cout, sp, stride, tr in [("conv 64->64 s1 48x96x312", 64, 64, (48, 96, 312), 1, False),
                                        ("conv 96->64 s1 192x20x304", 96, 64, (192, 20, 304), 1, False),
                                        ("conv 64->96 s1 (dgrad of 96->64)", 64, 96, (192, 20, 304), 1, False),
                                        ("conv 128->128 s1 24x48x156", 128, 128, (24, 48, 156), 1, False),
                                        ("conv 64->128 s2 48x96x312", 64, 128, (48, 96, 312), 2, False),
                                        ("deconv 128->64 s2 24x48x156", 128, 64, (24, 48, 156), 2, True)]:
    xc = cl3(1, cin, *sp)
    w = (rnd(cin, cout, 3, 3, 3) if tr else rnd(cout, cin, 3, 3, 3)) * 0.03
    vox = sp[0] * sp[1] * sp[2] if (tr or stride == 1) else (sp[0] // 2) * (sp[1] // 2) * (sp[2] // 2)
    timeit(name, lambda: ops.conv3d(xc, w, stride, tr, impl=0), 2 * cin * cout * 27 * vox, "TFLOP/s")
    del xc
# RoIAlign, config-5 shapes (FPN level 2 of a 600x1987 frame, 256 RoIs)
feat = rnd(1, 256, 150, 497).requires_grad_(True)
x1r, y1r = torch.rand(256, generator=g) * 1700, torch.rand(256, generator=g) * 500
rois = torch.stack([torch.zeros(256), x1r, y1r, x1r + torch.rand(256, generator=g) * 200 + 8, y1r + torch.rand(256, generator=g) * 90 + 8], 1).to(dev)
for pooled in (7, 14):
    timeit("roi_align_fwd P=%d R=256 C=256" % pooled, lambda: ops.roi_align(feat, rois, pooled, 0.25), 4 * (feat.numel() + 256 * 256 * pooled * pooled), "GB/s")
    ro = ops.roi_align(feat, rois, pooled, 0.25); gro = torch.randn_like(ro)
    timeit("roi_align_bwd P=%d (gather, deterministic)" % pooled, lambda: torch.autograd.grad(ro, feat, gro, retain_graph=True), 4 * (feat.numel() + ro.numel()), "GB/s", iters=5)
json.dump({"peaks": peaks, "kernels": R}, open(os.path.join(ROOT, "gpurun_out", "kernel_microbench.json"), "w"), indent=1)
