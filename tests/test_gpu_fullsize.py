"""GPU parity at the REAL size and on the TIMED path.

BASELINE config 1 (single-step FGSM, one 384x1248 pair) and config 2 (10-iteration L-inf PGD, eps 0.03,
alpha eps/4) on full-size synthetic pairs with the random-init DSGN-shaped model: the CPU oracle loop
(oracle/, stock torch fp32 ops) against exactly what bench.py times -- ``engine.PgdIterationGraph(lanes=2)``
replaying the CUDA graph of two concurrent pair-iterations, tcgen05 TF32 3-D convs (impl 0), own 3xTF32
2-D convs, every fused kernel.  north_star: "outputs must match ... within a stated tolerance on cost
volume, input gradient and final perturbation, with an identical perturbation sign pattern outside
near-zero gradients".

Stated tolerances (measured values are printed by the tests and recorded in DESIGN.md):
  * cost volume: bit-exact;
  * loss / depth / class maps: 5e-2 relative (measured 1e-5 .. 1e-3);
  * input gradient: 0.10 relative L2 (measured 5.6e-2: TF32 operands in the 3-D convs -- the fp32 SIMT
    verification mode gives 1.2e-2, i.e. most of what remains is summation order in an ill-conditioned
    random-init GroupNorm network);
  * sign pattern on the pixels with |g| > 1 % of max |g|: >= 99.9 % identical, every iteration;
  * one update step from the same iterate: >= 98 % of ALL pixels bit-identical to the oracle's;
  * final perturbation after 10 free-running iterations: stated in the test from the per-step figure.
"""
import os

import pytest
import torch

from helpers import rel_err
from oracle import attack_ref as A
from oracle import dsgn_ref as R

pytestmark = pytest.mark.gpu

K, EPS = 10, 0.03
ALPHA = EPS / 4
PAIRS = (0, 1)
MIN_AGREE, MIN_SAME = 0.999, 0.98


@pytest.fixture(scope="module")
def world(built_lib):
    """Both models with the same weights, and the ORACLE's K-iteration trajectory of every pair (iterates,
    gradients, losses): ~20 full-size CPU iterations, computed once for the whole module."""
    from eval_driving_safety_b200 import dsgn, ops, synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    ops.set_conv_impl(0)
    ops.set_conv2d_split(1)
    dsgn.set_backbone_impl("b2")
    cfg_r, cfg_p = R.default_cfg(), dsgn.default_cfg()
    ref = R.build_model(cfg_r, seed=1)
    for p in ref.parameters():
        p.requires_grad_(False)
    model = dsgn.StereoNet(cfg_p)
    model.load_state_dict(ref.state_dict())
    model = model.freeze().cuda()
    calib = synthetic.make_calib(1)
    labels = R.make_labels(cfg_r, 1, 7)
    traj = {}
    for i in PAIRS:
        pair = synthetic.make_pair(i)
        xL, xR = pair["imgL"].clone(), pair["imgR"].clone()
        cL, cR = A.denormalize(xL), A.denormalize(xR)
        steps = []
        for k in range(K):
            a, b = xL.clone().requires_grad_(True), xR.clone().requires_grad_(True)
            out = ref(a, b, *calib[:3], calibs_Proj_R=calib[3])
            loss = R.attack_loss(cfg_r, out, pair["disp_L"], labels)
            gL, gR = torch.autograd.grad(loss, [a, b])
            rec = dict(xL=xL, xR=xR, gL=gL, gR=gR, loss=loss.item())
            if k == 0:
                rec["out"] = {key: v.detach() for key, v in out.items()}
            steps.append(rec)
            xL = A.pgd_step_linf(xL, gL, cL, ALPHA, EPS)
            xR = A.pgd_step_linf(xR, gR, cR, ALPHA, EPS)
        traj[i] = dict(pair=pair, cL=cL, cR=cR, steps=steps, finalL=xL, finalR=xR)
    return dict(cfg_r=cfg_r, cfg_p=cfg_p, ref=ref, model=model, calib=calib, labels=labels,
                labels_gpu={k: v.cuda() for k, v in labels.items()}, traj=traj)


@pytest.fixture(scope="module")
def graph_engine(world):
    """The engine bench.py times: one CUDA graph holding two concurrent pair-iterations."""
    from eval_driving_safety_b200 import engine
    w = world
    t = w["traj"][PAIRS[0]]
    ex = (t["pair"]["imgL"].cuda(), t["pair"]["imgR"].cuda(), t["cL"].cuda(), t["cR"].cuda(), t["pair"]["disp_L"].cuda())
    return engine.PgdIterationGraph(w["model"], w["cfg_p"], w["labels_gpu"], w["calib"], ALPHA, EPS, ex, lanes=2)


def _gpu_grads(w, xL, xR, disp):
    from eval_driving_safety_b200 import dsgn
    a, b = xL.cuda().requires_grad_(True), xR.cuda().requires_grad_(True)
    out = w["model"](a, b, *w["calib"][:3], calibs_Proj_R=w["calib"][3])
    loss = dsgn.attack_loss(w["cfg_p"], out, disp.cuda(), w["labels_gpu"])
    gL, gR = torch.autograd.grad(loss, [a, b])
    return out, loss, gL, gR


def _sign_agree(g, g_ref):
    big = g_ref.abs() > 1e-2 * g_ref.abs().max()
    return (g.sign() == g_ref.sign())[big].float().mean().item(), big.float().mean().item()


def test_config1_fgsm_parity(world):
    """BASELINE config 1: one FGSM step (alpha = eps = 8/255) on pair 0."""
    from eval_driving_safety_b200 import attack
    w = world
    t = w["traj"][0]
    s0 = t["steps"][0]
    out_g, loss_g, gL_g, gR_g = _gpu_grads(w, s0["xL"], s0["xR"], t["pair"]["disp_L"])
    eps = alpha = 8 / 255
    advL_r = A.pgd_step_linf(s0["xL"], s0["gL"], t["cL"], alpha, eps)
    advL_g = attack.pgd_step(s0["xL"].cuda(), gL_g.contiguous(), t["cL"].cuda(), alpha, eps)
    e_depth = rel_err(out_g["depth_preds"].cpu(), s0["out"]["depth_preds"])
    e_cls = rel_err(out_g["bbox_cls"].cpu(), s0["out"]["bbox_cls"])
    e_reg = rel_err(out_g["bbox_reg"].cpu(), s0["out"]["bbox_reg"])
    e_loss = abs(loss_g.item() - s0["loss"]) / abs(s0["loss"])
    e_gL, e_gR = rel_err(gL_g.cpu(), s0["gL"]), rel_err(gR_g.cpu(), s0["gR"])
    agree, frac = _sign_agree(gL_g.cpu(), s0["gL"])
    same = ((advL_g.cpu() - advL_r).abs() < 1e-5).float().mean().item()
    print("\nCONFIG 1 (FGSM, full size): depth %.2e cls %.2e reg %.2e loss %.2e gradL %.2e gradR %.2e | sign(|g|>1%%max, %.0f%% of px) "
          "%.5f | FGSM pixels identical %.5f" % (e_depth, e_cls, e_reg, e_loss, e_gL, e_gR, 100 * frac, agree, same))
    assert e_depth < 5e-2 and e_cls < 5e-2 and e_reg < 5e-2 and e_loss < 5e-2
    assert e_gL < 0.10 and e_gR < 0.10
    assert agree >= MIN_AGREE and same >= MIN_SAME


def test_cost_volume_equality_at_full_size(world):
    """The oracle's features through both cost volumes: 368 MB, bit for bit."""
    from eval_driving_safety_b200 import dsgn, ops
    w = world
    t = w["traj"][0]
    with torch.no_grad():
        fL, _ = w["ref"].feature_extraction(t["pair"]["imgL"])
        fR, _ = w["ref"].feature_extraction(t["pair"]["imgR"])
        shifts = R.plane_shifts(w["cfg_r"], w["calib"][0], w["calib"][1])
        cost_r = R.build_cost_volume(fL, fR, shifts)
        cost_g = ops.build_cost_volume(fL.cuda(), fR.cuda(), dsgn.plane_shifts(w["cfg_p"], w["calib"][0], w["calib"][1]).cuda())
    assert cost_g.shape == cost_r.shape == (1, 64, 48, 96, 312)
    assert torch.equal(cost_g.cpu(), cost_r)


def test_config2_every_iteration_on_the_timed_path(world, graph_engine):
    """Teacher-forced: at every iteration k both pairs' ORACLE iterates go through the captured two-lane graph
    (forward + backward + fused update, as benchmarked); the updated images must equal the oracle's next
    iterates on >= 98 % of all pixels, and the gradient sign must agree on >= 99.9 % of the pixels with
    |g| > 1 % max."""
    w, eng = world, graph_engine
    min_agree, min_same = 1.0, 1.0
    for k in range(K):
        bufs = []
        for i in PAIRS:
            t = w["traj"][i]
            s = t["steps"][k]
            bufs.append([s["xL"].cuda(), s["xR"].cuda(), t["cL"].cuda(), t["cR"].cuda(), t["pair"]["disp_L"].cuda()])
        losses = [l.item() for l in eng.step_multi([tuple(b) for b in bufs])]
        torch.cuda.synchronize()
        for j, i in enumerate(PAIRS):
            t = w["traj"][i]
            nxtL = t["steps"][k + 1]["xL"] if k + 1 < K else t["finalL"]
            nxtR = t["steps"][k + 1]["xR"] if k + 1 < K else t["finalR"]
            sameL = ((A.denormalize(bufs[j][0].cpu()) - A.denormalize(nxtL)).abs() < 1e-6).float().mean().item()
            sameR = ((A.denormalize(bufs[j][1].cpu()) - A.denormalize(nxtR)).abs() < 1e-6).float().mean().item()
            _, _, gL, gR = _gpu_grads(w, t["steps"][k]["xL"], t["steps"][k]["xR"], t["pair"]["disp_L"])
            agL, _ = _sign_agree(gL.cpu(), t["steps"][k]["gL"])
            agR, _ = _sign_agree(gR.cpu(), t["steps"][k]["gR"])
            e_loss = abs(losses[j] - t["steps"][k]["loss"]) / abs(t["steps"][k]["loss"])
            print("iter %d pair %d: loss rel %.1e | sign agreement L %.5f R %.5f | updated pixels identical L %.5f R %.5f"
                  % (k, i, e_loss, agL, agR, sameL, sameR))
            assert e_loss < 5e-2
            min_agree, min_same = min(min_agree, agL, agR), min(min_same, sameL, sameR)
    print("CONFIG 2 teacher-forced over %d iterations x %d pairs: min_agree %.5f  min_same %.5f" % (K, len(PAIRS), min_agree, min_same))
    assert min_agree >= MIN_AGREE and min_same >= MIN_SAME


def test_config2_final_perturbation_free_running(world, graph_engine):
    """Free-running: K replays of the graph from the clean pair against the oracle's K-iteration loop.  A pixel
    whose update differs in ONE iteration ends 2*alpha = eps/2 away (unless later clamped back), so with ~1.5 %
    of the pixels per iteration differing (previous test) and feedback through the network the final
    perturbations agree exactly on most pixels and by well under eps/2 on average."""
    w, eng = world, graph_engine
    xs = []
    for i in PAIRS:
        t = w["traj"][i]
        xs.append([t["pair"]["imgL"].cuda(), t["pair"]["imgR"].cuda(), t["cL"].cuda(), t["cR"].cuda(), t["pair"]["disp_L"].cuda()])
    for k in range(K):
        eng.step_multi([tuple(b) for b in xs])
    torch.cuda.synchronize()
    for j, i in enumerate(PAIRS):
        t = w["traj"][i]
        d_g = A.denormalize(xs[j][0].cpu()) - t["cL"]
        d_r = A.denormalize(t["finalL"]) - t["cL"]
        assert d_g.abs().max() <= EPS + 1e-6 and (t["cL"] + d_g).min() >= -1e-6 and (t["cL"] + d_g).max() <= 1 + 1e-6
        same = ((d_g - d_r).abs() < 1e-6).float().mean().item()
        mean_dev = (d_g - d_r).abs().mean().item() / EPS
        sign_same = (d_g.sign() == d_r.sign()).float().mean().item()
        print("CONFIG 2 free-running pair %d: final perturbation identical on %.4f of the pixels, mean |delta - delta_ref| = %.4f eps, "
              "same direction on %.4f" % (i, same, mean_dev, sign_same))
        assert same >= 0.85 and mean_dev <= 0.08 and sign_same >= 0.90
