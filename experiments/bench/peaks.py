"""Dense tensor-core peaks of this box measured the way MEASURED_PEAKS.json measured bf16 (torch.matmul 8192^3: best of
10 = burst, back to back for 4 s = sustained), for the dtypes the roofline of pass 1 is quoted against: tf32 (fp32 inputs,
allow_tf32), fp16, bf16.  Library GEMMs are used here ONLY as the yard-stick, never on the product path."""
import json
import sys
import time

import torch


def measure(dtype, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    n = 8192
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    c = torch.empty(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    t0 = time.time()
    iters = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(20):
            torch.matmul(a, b, out=c)
        iters += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / iters
    fl = 2.0 * n ** 3
    return {"burst_tflops": fl / best / 1e9, "sustained_tflops": fl / sus / 1e9}


if __name__ == "__main__":
    out = {"gpu": torch.cuda.get_device_name(0), "how": __doc__.strip()}
    out["tf32"] = measure(torch.float32, True)
    out["fp16"] = measure(torch.float16, False)
    out["bf16"] = measure(torch.bfloat16, False)
    print(json.dumps(out))
    json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/peaks.json", "w"), indent=1)
