"""Why the Khatri-Rao stage takes 8.5 ms inside the c3 step and 5.95 ms alone: time the producers (and a plain device copy
of the same size) directly behind a pass-1 launch, i.e. at the SM clock the power cap leaves, next to the same kernels
timed back to back on a cool chip.  NVML clock samples beside each measurement."""
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
import gppvae_b200  # noqa: E402
from gppvae_b200 import _lib, ops  # noqa: E402
from gppvae_b200.synth import make_tables  # noqa: E402

dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
p, q, L = 256, 16, 256
Q = p * q
gen = torch.Generator(device=dev).manual_seed(0)
P = -(-N // q)
x0, v0 = make_tables(P, p, q, "trained", gen, dev)
xn = ops.normalize_rows_fwd(x0)
wn = ops.normalize_rows_fwd(v0)
perm = torch.randperm(N, device=dev)
d, w = (perm // q).contiguous(), (perm % q).contiguous()
Z = torch.randn(N, L, device=dev, generator=gen)
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
V = torch.empty(N, Q, device=dev)
buf = torch.empty(lib.gpp_planes_bytes(N, Q), dtype=torch.uint8, device=dev)
ws = torch.empty(lib.gpp_split_workspace_bytes(N, Q), dtype=torch.uint8, device=dev)
src = torch.empty(N * Q * 4, dtype=torch.uint8, device=dev)      # 16.4 GB copy: 32.8 GB of traffic, like the producer
dst = torch.empty_like(src)

import pynvml  # noqa: E402
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False


def sampler():
    while not stop:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
        time.sleep(0.002)


def kr():
    lib.gpp_khatri_rao_fwd_planes(xn.data_ptr(), P, p, wn.data_ptr(), q, q, d.data_ptr(), w.data_ptr(), N, V.data_ptr(), Q,
                                  buf.data_ptr(), buf.numel(), ws.data_ptr(), ws.numel(), st)


def kr32():
    lib.gpp_khatri_rao_fwd(xn.data_ptr(), P, p, wn.data_ptr(), q, q, d.data_ptr(), w.data_ptr(), N, V.data_ptr(), Q, st)


def copy():
    dst.copy_(src)


kr()
pV = ops.Planes(buf, N, Q, True)
pZ = ops.split_planes(Z, L, N, L)


def pass1():
    return ops.gram_vtz_planes(pV, pZ, N, Q, L)


def measure(name, fn, nbytes, hot):
    ms = []
    for it in range(6):
        if hot:
            pass1()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if not hot:
            time.sleep(0.05)
        ms.append(e0.elapsed_time(e1))
    ms = sorted(ms[1:])
    m = ms[len(ms) // 2]
    print(f"{name:28s} {'behind pass 1' if hot else 'cool chip    '} {m:7.3f} ms  {nbytes / m / 1e9:5.2f} TB/s", flush=True)


t = threading.Thread(target=sampler, daemon=True)
t.start()
for hot in (False, True):
    measure("khatri_rao_fwd_planes", kr, N * Q * 8, hot)
    measure("khatri_rao_fwd (fp32 only)", kr32, N * Q * 4, hot)
    measure("device copy 16.4 GB", copy, N * Q * 8, hot)
stop = True
t.join()
sm = sorted(s[1] for s in samples)
print("SM clock MHz min / median / max:", sm[0], sm[len(sm) // 2], sm[-1], " mem clock:", sorted(set(s[2] for s in samples)),
      " power W max:", max(s[3] for s in samples))
