"""Measured dense TF32 throughput of this GPU (cuBLAS, 8192^3, fp32 storage with allow_tf32) next to the bf16
figure the driver's MEASURED_PEAKS.json holds: the denominator of the conv kernels' roofline (VERDICT r1 item 14:
"the /2 is an assumption; measure it").  Writes profiles/r2_tf32_peak.json."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda", 0)
n = 8192


def rate(dtype, tf32, seconds=0.0):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device=dev, dtype=dtype)
    b = torch.randn(n, n, device=dev, dtype=dtype)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):                                  # burst: best single matmul
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = max(best, 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 0
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < 4.0:               # sustained: back to back for 4 s
        for _ in range(20):
            a @ b
        k += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return best / 1e12, 2 * n ** 3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12


out = {}
out["tf32_tflops"], out["tf32_tflops_sustained"] = rate(torch.float32, True)
out["bf16_tflops"], out["bf16_tflops_sustained"] = rate(torch.bfloat16, False)
out["fp32_simt_tflops"], _ = rate(torch.float32, False)
out["how"] = "torch.matmul 8192^3 (2*N^3): best of 10 (burst) and back to back for 4 s (sustained); tf32 = fp32 tensors with allow_tf32"
out["gpu"] = torch.cuda.get_device_name(0)
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_tf32_peak.json"), "w"), indent=1)
