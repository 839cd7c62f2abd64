"""The two stride-2 kernel classes at their KITTI shapes (graph replay, 10 launches): transposed 128 -> 64 at 24x48x156
and 128 -> 128 at 12x24x78 (class-stacked kernel; B2_CONV_DC_PAIR=0 selects the single-CTA variant), stride-2
64 -> 128 at 48x96x312 (generic kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eval_driving_safety_b200 import ops
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)


def cl(n, c, d, h, w):
    return torch.randn(n, d, h, w, c, generator=g).to(dev).permute(0, 4, 1, 2, 3)


def timeit(fn, iters=10):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(iters):
                fn()
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            gr.replay()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * iters)


from eval_driving_safety_b200 import _lib
for pair in (0, 1):
  _lib.set_flag("conv_dc_pair", pair)
  _lib.set_flag("conv_s2_pair", pair)
  print("conv_dc_pair = conv_s2_pair =", pair)
  for name, x, w, stride, mode in (
          ("deconv 128->64  @24x48x156", cl(1, 128, 24, 48, 156), (torch.randn(27, 64, 128, generator=g) * 0.02).to(dev), 2, 1),
          ("deconv 128->128 @12x24x78 ", cl(1, 128, 12, 24, 78), (torch.randn(27, 128, 128, generator=g) * 0.02).to(dev), 2, 1),
          ("deconv 128->64  @96x10x152 (3DGV)", cl(1, 128, 96, 10, 152), (torch.randn(27, 64, 128, generator=g) * 0.02).to(dev), 2, 1),
          ("conv s2 64->128 @48x96x312", cl(1, 64, 48, 96, 312), (torch.randn(27, 128, 64, generator=g) * 0.02).to(dev), 2, 0),
        ("conv s2 128->128 @24x48x156", cl(1, 128, 24, 48, 156), (torch.randn(27, 128, 128, generator=g) * 0.02).to(dev), 2, 0),
        ("conv s2 64->128 @192x20x304 (3DGV)", cl(1, 64, 192, 20, 304), (torch.randn(27, 128, 64, generator=g) * 0.02).to(dev), 2, 0)):
      n, cin, d, h, wd = x.shape
      cout = w.shape[1]
      vox = n * d * h * wd if mode == 1 else n * (d // 2) * (h // 2) * (wd // 2)
      fl = 2 * cin * cout * 27 * vox
      t = timeit(lambda: ops._conv_call(x, w, stride, mode, 0))
      print("  %s: %.4f ms  %.0f TFLOP/s" % (name, t, fl / t / 1e9), flush=True)
_lib.set_flag("conv_dc_pair", None); _lib.set_flag("conv_s2_pair", None)
