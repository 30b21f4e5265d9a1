"""Streaming-write paths on a cool chip and directly behind a pass-1 launch (power-capped clock): per-thread STG.128 / STG.64
against cp.async.bulk shared -> global (experiments/tc/exp5_store_paths.cu), and a device copy for scale."""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from gppvae_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
lib = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "..", "tc", "libexp5_store_paths.so"))
N, Q, L = 1_000_000, 4096, 256
BYTES = N * Q * 4
out = torch.empty(BYTES, dtype=torch.uint8, device=dev)
src = torch.empty(BYTES, dtype=torch.uint8, device=dev)
V = torch.empty(N, Q, device=dev).uniform_(-1, 1)
pV = ops.split_planes(V, Q, N, Q, colsq=True, unit_bound=True)
del V
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
vp = ctypes.c_void_p(out.data_ptr())


def burst():
    ops.gram_vtz_planes(pV, None, N, Q, 0)


def measure(name, fn, nbytes):
    res = []
    for hot in (False, True):
        ms = []
        for _ in range(5):
            if hot:
                burst()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
            if not hot:
                time.sleep(0.05)
        ms = sorted(ms[1:])
        res.append(ms[len(ms) // 2])
    print(f"{name:44s} cool {res[0]:7.3f} ms {nbytes / res[0] / 1e9:5.2f} TB/s | behind pass 1 {res[1]:7.3f} ms {nbytes / res[1] / 1e9:5.2f} TB/s",
          flush=True)


measure("device copy (read 16.4 GB + write 16.4 GB)", lambda: out.copy_(src), 2 * BYTES)
measure("memset 16.4 GB", lambda: out.zero_(), BYTES)
for ctas in (148 * 2, 148 * 4, 148 * 8):
    measure(f"STG.128, {ctas} CTAs", lambda: lib.exp5_stg128(vp, ctypes.c_longlong(BYTES), ctas, st), BYTES)
measure("STG.64, 1184 CTAs", lambda: lib.exp5_stg64(vp, ctypes.c_longlong(BYTES), 1184, st), BYTES)
for ctas, chunk, depth in ((148, 16384, 4), (296, 16384, 4), (592, 16384, 4), (296, 4096, 8), (296, 32768, 2), (592, 32768, 4)):
    measure(f"bulk store {chunk} B x depth {depth}, {ctas} CTAs",
            lambda: lib.exp5_bulk(vp, ctypes.c_longlong(BYTES), ctas, chunk, depth, st), BYTES)
