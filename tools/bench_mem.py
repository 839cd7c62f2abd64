import torch, json
dev = torch.device("cuda", 0)
flush_buf = torch.empty(64 * 1024 * 1024, device=dev)
a = torch.empty(92012544, device=dev); b = torch.empty_like(a)
def t(name, fn, nbytes, flush=True):
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        if flush: flush_buf.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[5]
    print("%-40s %.4f ms  %.0f GB/s" % (name, ms, nbytes / ms / 1e6))
t("fill_ 368MB (write only), flushed", lambda: a.fill_(2.0), a.numel() * 4)
t("fill_ 368MB (write only), no flush", lambda: a.fill_(2.0), a.numel() * 4, False)
t("copy_ 368MB->368MB, flushed", lambda: b.copy_(a), a.numel() * 8)
t("copy_ no flush", lambda: b.copy_(a), a.numel() * 8, False)
t("sum 368MB (read only), flushed", lambda: a.sum(), a.numel() * 4)
t("empty launch (x.add_ on 1 elem)", lambda: flush_buf[:1].add_(1), 4)
