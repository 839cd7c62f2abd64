"""Test-only helpers shared by CPU and GPU tests."""
import torch
import torch.nn.functional as F


def gather_conv_ref(x, wp, stride, mode):
    """Literal evaluation of the gather definition in include/b2attack.h on logical NCDHW
    tensors (any device): CONV out[o] = sum_k in[o*s+k-1].wp[k]; DECONV out[o] = sum_k
    [(o+1-k) even] in[(o+1-k)/2].wp[k].  wp [27,Cout,Cin]."""
    n, cin, di, hi, wi = x.shape
    cout = wp.shape[1]
    if mode == 0:
        do, ho, wo = (di - 1) // stride + 1, (hi - 1) // stride + 1, (wi - 1) // stride + 1
    else:
        do, ho, wo = 2 * di, 2 * hi, 2 * wi
    out = torch.zeros(n, cout, do, ho, wo, dtype=x.dtype, device=x.device)
    xp = x.permute(0, 2, 3, 4, 1)                                   # [N,D,H,W,Cin]
    od, oh, ow = torch.arange(do), torch.arange(ho), torch.arange(wo)

    def coords(o, k, isz):
        if mode == 0:
            i = o * stride + k - 1
            return i.clamp(0, isz - 1), (i >= 0) & (i < isz)
        t = o + 1 - k
        ok = (t >= 0) & (t % 2 == 0) & (t // 2 < isz)
        return (t // 2).clamp(0, isz - 1), ok

    for kd in range(3):
        idd, okd = coords(od, kd, di)
        for kh in range(3):
            ihh, okh = coords(oh, kh, hi)
            for kw in range(3):
                iww, okw = coords(ow, kw, wi)
                g = xp[:, idd][:, :, ihh][:, :, :, iww]             # [N,Do,Ho,Wo,Cin]
                m = (okd.view(-1, 1, 1) & okh.view(1, -1, 1) & okw.view(1, 1, -1)).to(x.dtype).to(x.device)
                w = wp[(kd * 3 + kh) * 3 + kw]                      # [Cout,Cin]
                out += torch.einsum('ndhwc,oc->nodhw', g * m.view(1, do, ho, wo, 1), w)
    return out


def rel_err(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def max_err(a, b):
    return (a.double() - b.double()).abs().max().item()
