"""Bisect: which part of the extractor changes its input gradient when the 2-D GroupNorm backward sums are fused."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from eval_driving_safety_b200 import dsgn, ops, synthetic

H, W = 32, 64
cfg = dsgn.tiny_cfg()
model = dsgn.build_model(cfg, seed=1, device=torch.device("cuda", 0))
fe = model.feature_extraction
pair = synthetic.make_pair(0, H, W, max_depth=8.4)
ops.set_conv_impl(1)


def sub(stage, x):
    out = x
    for i in (0, 2, 4):
        out = dsgn.run_convbn_2d(fe.firstconv[i], out, relu=True)
        if stage == "first%d" % i:
            return [out]
    out = fe.layer1(out)
    if stage == "layer1":
        return [out]
    raw = fe.layer2(out)
    if stage == "layer2":
        return [raw]
    l3 = fe.layer3(raw)
    if stage == "layer3":
        return [l3]
    if stage == "layer3+raw":
        return [l3, raw]
    skip = fe.layer4(l3)
    if stage == "layer4":
        return [skip]
    return [raw, skip]


for stage in ("first0", "first2", "first4", "layer1", "layer2", "layer3", "layer3+raw", "layer4", "raw+skip"):
    res = []
    for fuse in (False, True):
        ops.FUSE_GN_BWD = fuse
        x = torch.cat([pair["imgL"], pair["imgR"]], 0).cuda().requires_grad_(True)
        outs = sub(stage, x)
        g = torch.Generator().manual_seed(3)
        loss = sum((o * torch.randn(o.shape, generator=g).cuda()).sum() for o in outs)
        res.append(torch.autograd.grad(loss, x)[0])
    print(stage, "%.2e" % float((res[1] - res[0]).abs().max() / res[0].abs().max()), flush=True)


def staged(x, grads):
    def keep(name):
        def h(g):
            grads[name] = g.detach().clone()
        return h
    out = x
    for i in (0, 2, 4):
        out = dsgn.run_convbn_2d(fe.firstconv[i], out, relu=True)
        out.register_hook(keep("first%d" % i))
    for bi, blk in enumerate(fe.layer1):
        out = blk(out); out.register_hook(keep("layer1.%d" % bi))
    raw = out
    for bi, blk in enumerate(fe.layer2):
        raw = blk(raw); raw.register_hook(keep("layer2.%d" % bi))
    l3 = raw
    for bi, blk in enumerate(fe.layer3):
        l3 = blk(l3); l3.register_hook(keep("layer3.%d" % bi))
    skip = l3
    for bi, blk in enumerate(fe.layer4):
        skip = blk(skip); skip.register_hook(keep("layer4.%d" % bi))
    return [raw, skip]


allg = []
for fuse in (False, True):
    ops.FUSE_GN_BWD = fuse
    x = torch.cat([pair["imgL"], pair["imgR"]], 0).cuda().requires_grad_(True)
    grads = {}
    outs = staged(x, grads)
    g = torch.Generator().manual_seed(3)
    loss = sum((o * torch.randn(o.shape, generator=g).cuda()).sum() for o in outs)
    grads["input"] = torch.autograd.grad(loss, x)[0]
    allg.append(grads)
for k in allg[0]:
    a, b = allg[0][k], allg[1][k]
    print("grad at", k, "%.2e" % float((a - b).abs().max() / a.abs().max()), flush=True)
