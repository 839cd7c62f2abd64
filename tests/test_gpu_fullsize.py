"""GPU parity at the REAL size: one FGSM iteration on a 384x1248 synthetic pair, B200 path (tcgen05
TF32 convs, every fused kernel, as benchmarked) against the CPU oracle with the same weights --
loss, input gradient, sign pattern outside near-zero gradients, final perturbation (north_star)."""
import os

import pytest
import torch

from helpers import rel_err
from oracle import attack_ref as A
from oracle import dsgn_ref as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cudnn_tf32,tol_grad,min_agree,min_same", [(False, 0.10, 0.995, 0.97), (True, 0.25, 0.98, 0.93)])
def test_fullsize_fgsm_parity(built_lib, cudnn_tf32, tol_grad, min_agree, min_same):
    """cudnn_tf32=False isolates OUR kernels (TF32 tcgen05 3-D convs + everything else of libb2attack; the
    stock 2-D convolutions in fp32): measured 5.6e-2 gradient error, 99.93 % sign agreement on the pixels
    with |g| > 1 % of max, 98.4 % of FGSM pixels identical.  cudnn_tf32=True is the benchmarked setting
    (cuDNN's 2-D convs in TF32 as well, PyTorch's default): measured 1.45e-1 / 99.06 % / 95.5 % -- the stock
    2-D TF32 convolutions contribute most of the gap.  (All-fp32 GPU vs CPU already differs by 8.7e-3 in
    the gradient: at random init this deep GroupNorm network is ill-conditioned.)"""
    from eval_driving_safety_b200 import attack, dsgn, ops, synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    torch.backends.cudnn.allow_tf32 = cudnn_tf32
    torch.backends.cuda.matmul.allow_tf32 = cudnn_tf32
    ops.set_conv_impl(0)
    cfg_r, cfg_p = R.default_cfg(), dsgn.default_cfg()
    ref = R.build_model(cfg_r, seed=1)
    for p in ref.parameters():
        p.requires_grad_(False)
    model = dsgn.StereoNet(cfg_p)
    model.load_state_dict(ref.state_dict())
    model = model.freeze().cuda()
    pair = synthetic.make_pair(0)
    calib = synthetic.make_calib(1)
    labels = R.make_labels(cfg_r, 1, 7)
    eps = alpha = 8 / 255
    # CPU oracle
    xL, xR = pair["imgL"].clone().requires_grad_(True), pair["imgR"].clone().requires_grad_(True)
    out_r = ref(xL, xR, *calib[:3], calibs_Proj_R=calib[3])
    loss_r = R.attack_loss(cfg_r, out_r, pair["disp_L"], labels)
    gL_r, gR_r = torch.autograd.grad(loss_r, [xL, xR])
    advL_r = A.pgd_step_linf(pair["imgL"], gL_r, A.denormalize(pair["imgL"]), alpha, eps)
    # B200 path
    xLc, xRc = pair["imgL"].cuda().requires_grad_(True), pair["imgR"].cuda().requires_grad_(True)
    out_g = model(xLc, xRc, *calib[:3], calibs_Proj_R=calib[3])
    loss_g = dsgn.attack_loss(cfg_p, out_g, pair["disp_L"].cuda(), {k: v.cuda() for k, v in labels.items()})
    gL_g, gR_g = torch.autograd.grad(loss_g, [xLc, xRc])
    clean = pair["imgL"].cuda() * torch.tensor(A.IMAGENET_STD).view(1, 3, 1, 1).cuda() + \
        torch.tensor(A.IMAGENET_MEAN).view(1, 3, 1, 1).cuda()
    advL_g = attack.pgd_step(pair["imgL"].cuda(), gL_g.contiguous(), clean, alpha, eps)
    # --- report + bounds (TF32 tensor-core convs; bounds as in test_gpu_e2e, stated there) ---
    e_depth = rel_err(out_g["depth_preds"].cpu(), out_r["depth_preds"])
    e_cls = rel_err(out_g["bbox_cls"].cpu(), out_r["bbox_cls"])
    e_loss = abs(loss_g.item() - loss_r.item()) / abs(loss_r.item())
    e_gL, e_gR = rel_err(gL_g.cpu(), gL_r), rel_err(gR_g.cpu(), gR_r)
    big = gL_r.abs() > 1e-2 * gL_r.abs().max()
    agree = (gL_g.cpu().sign() == gL_r.sign())[big].float().mean().item()
    agree_all = (gL_g.cpu().sign() == gL_r.sign()).float().mean().item()
    same_pix = ((advL_g.cpu() - advL_r).abs() < 1e-5).float().mean().item()
    print("\nFULL-SIZE PARITY (cudnn_tf32=%s)" % cudnn_tf32, " depth %.2e  cls %.2e  loss %.2e  gradL %.2e  gradR %.2e  sign(|g|>1%%max, %.1f%% of px) %.5f  "
          "sign(all px) %.5f  FGSM pixels identical %.5f" % (e_depth, e_cls, e_loss, e_gL, e_gR, 100 * big.float().mean().item(),
                                                          agree, agree_all, same_pix))
    assert e_depth < 5e-2 and e_cls < 5e-2 and e_loss < 5e-2
    assert e_gL < tol_grad and e_gR < tol_grad
    assert agree >= min_agree and same_pix >= min_same
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
