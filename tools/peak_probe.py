"""Measures the matmul peaks the roofline fractions divide by, with the probe MEASURED_PEAKS.json documents
(torch.matmul 8192^3: best of 10 = burst, back to back for 4 s = sustained), for bf16 AND tf32 - the driver's file has
no tf32 figure.  Library GEMM (cuBLAS) used as a yardstick only; nothing on the product path calls it.
usage: python tools/peak_probe.py > gpurun_out/peaks_probe.json"""
import json
import time

import torch


def probe(dtype, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    n = 8192
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    flops = 2.0 * n ** 3
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    it = 0
    e0.record()
    while time.perf_counter() - t0 < 4.0:
        for _ in range(20):
            a @ b
        it += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return round(best, 1), round(flops * it / (e0.elapsed_time(e1) * 1e-3) / 1e12, 1)


if __name__ == "__main__":
    out = {"gpu": torch.cuda.get_device_name(0), "how": "torch.matmul 8192^3, best of 10 (burst) / back to back 4 s (sustained)"}
    out["bf16_tflops"], out["bf16_tflops_sustained"] = probe(torch.bfloat16, False)
    out["tf32_tflops"], out["tf32_tflops_sustained"] = probe(torch.float32, True)
    # HBM copy, as MEASURED_PEAKS: b.copy_(a) over 1 Gi bf16 elements, read + write bytes, best of 10
    a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda")
    b = torch.empty_like(a)
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2 * a.numel() * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    out["hbm_gbs"] = round(best, 1)
    print(json.dumps(out))
