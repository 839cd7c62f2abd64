"""Does a tensor-bound conv (stream A) overlap with HBM-bound GroupNorm (stream B)?  Serial vs concurrent."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eval_driving_safety_b200 import ops
g = torch.Generator().manual_seed(0)
dev = torch.device("cuda", 0)
x1 = torch.randn(1, 48, 96, 312, 64, generator=g).to(dev).permute(0, 4, 1, 2, 3)
x2 = torch.randn(1, 48, 96, 312, 64, generator=g).to(dev).permute(0, 4, 1, 2, 3)
w = (torch.randn(64, 64, 3, 3, 3, generator=g) * 0.02).to(dev)
gamma, beta = torch.ones(64, device=dev), torch.zeros(64, device=dev)
PRIO = int(os.environ.get('CONV_PRIO', '0'))
sa, sb = torch.cuda.Stream(priority=PRIO), torch.cuda.Stream()
NC, NG = 4, 10

def conv_work():
    for _ in range(NC):
        ops.conv3d(x1, w, stride=1, transposed=False)

def gn_work():
    for _ in range(NG):
        ops.groupnorm_act(x2, gamma, beta, 32, 1e-5, relu=True, res=None)

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def both():
    cur = torch.cuda.current_stream()
    sa.wait_stream(cur); sb.wait_stream(cur)
    with torch.cuda.stream(sa): conv_work()
    with torch.cuda.stream(sb): gn_work()
    cur.wait_stream(sa); cur.wait_stream(sb)

with torch.no_grad():
    tc, tg, tb = timeit(conv_work), timeit(gn_work), timeit(both)
print(f"conv x{NC}: {tc:.3f} ms   gn x{NG}: {tg:.3f} ms   serial sum {tc+tg:.3f}   concurrent {tb:.3f} ms")
buf = torch.empty_like(x2)
def copy_work():
    for _ in range(2 * NG):
        buf.copy_(x2)
gn_work_saved = gn_work
gn_work = copy_work
def both2():
    cur = torch.cuda.current_stream()
    sa.wait_stream(cur); sb.wait_stream(cur)
    with torch.cuda.stream(sa): conv_work()
    with torch.cuda.stream(sb): copy_work()
    cur.wait_stream(sa); cur.wait_stream(sb)
with torch.no_grad():
    tg, tb = timeit(copy_work), timeit(both2)
print(f"conv x{NC}: {tc:.3f} ms   copy x{2*NG}: {tg:.3f} ms   serial sum {tc+tg:.3f}   concurrent {tb:.3f} ms")
# copy launched FIRST (stream order reversed)
def both3():
    cur = torch.cuda.current_stream()
    sa.wait_stream(cur); sb.wait_stream(cur)
    with torch.cuda.stream(sb): copy_work()
    with torch.cuda.stream(sa): conv_work()
    cur.wait_stream(sa); cur.wait_stream(sb)
with torch.no_grad():
    tb = timeit(both3)
print(f"copy first: concurrent {tb:.3f} ms")

# per-stream durations inside the concurrent run
def both_timed(work_b):
    cur = torch.cuda.current_stream()
    ea0, ea1, eb0, eb1 = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    sa.wait_stream(cur); sb.wait_stream(cur)
    with torch.cuda.stream(sa):
        ea0.record(); conv_work(); ea1.record()
    with torch.cuda.stream(sb):
        eb0.record(); work_b(); eb1.record()
    cur.wait_stream(sa); cur.wait_stream(sb)
    torch.cuda.synchronize()
    return ea0.elapsed_time(ea1), eb0.elapsed_time(eb1), ea0.elapsed_time(eb1)
with torch.no_grad():
    for name, wk in (("gn", gn_work_saved), ("copy", copy_work)):
        both_timed(wk)
        print(name, "conv stream %.3f ms, other stream %.3f ms, conv-start->other-end %.3f" % both_timed(wk))
