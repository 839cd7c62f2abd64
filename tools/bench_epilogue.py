"""Cost of the fused epilogue options on the two big kernels (CUDA events, L2 flushed between launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eval_driving_safety_b200 import ops
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def cl(shape):
    n, c, d, h, w = shape
    return torch.randn(n, d, h, w, c, generator=g).to(dev).permute(0, 4, 1, 2, 3)

def timeit(fn, n=8):
    ts = []
    for _ in range(n + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts[2:])[len(ts[2:]) // 2]

for name, cin, cout, stride, mode, sp in (("s1n 64->64 48x96x312", 64, 64, 1, 0, (48, 96, 312)),
                                          ("dc 128->64 24x48x156", 128, 64, 2, 1, (24, 48, 156))):
    x = cl((1, cin) + sp)
    wp = (torch.randn(27, cout, cin, generator=g) * 0.02).to(dev)
    osp = sp if mode == 0 else tuple(2 * s for s in sp)
    addend = cl((1, cout) + osp)
    link = ops.GnLink()
    link.x = cl((1, cout) + osp); link.groups = 32
    link.stats = torch.cat([torch.zeros(64), torch.rand(cout) + 0.5, torch.randn(cout)]).view(1, -1).to(dev)
    for label, kw in (("plain", {}), ("fwd stats", dict(stats=True)), ("addend", dict(addend=addend)),
                      ("bwd sums (no mask)", dict(bstat=(link, 0))), ("bwd sums + mask", dict(bstat=(link, 2))),
                      ("addend + bwd sums + mask", dict(addend=addend, bstat=(link, 2)))):
        kw = dict(kw)
        if "bstat" in kw:
            l, m = kw["bstat"]; l.mode = m; kw["bstat"] = l
        t = timeit(lambda: ops._conv_call(x, wp, stride, mode, 0, **kw))
        print("%-24s %-28s %.3f ms" % (name, label, t))
