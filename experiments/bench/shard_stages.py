"""Per-kernel timing of one rank's share of c3 at 8 ranks (n = 125k rows, Q = 4096, L = 256) on ONE GPU: Gram tiles and
V^T Z as the sharded path launches them (two launches), at different split counts (GPP_TC_SPLIT_ROWS), each timed alone
with CUDA events -- back to back (hot) and with pauses (cool) -- to see what the 5.4 ms of `pass1` at 8 GPUs are made of."""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from gppvae_b200 import _lib, ops  # noqa: E402
from gppvae_b200.synth import make_tables  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
p, q, L = 256, 16, 256
Q = p * q
gen = torch.Generator(device=dev).manual_seed(0)
P = -(-1_000_000 // q)
x0, v0 = make_tables(P, p, q, "trained", gen, dev)
xn, wn = ops.normalize_rows_fwd(x0), ops.normalize_rows_fwd(v0)
perm = torch.randperm(1_000_000, device=dev)[:n]
d, w = (perm // q).contiguous(), (perm % q).contiguous()
Z = torch.randn(n, L, device=dev, generator=gen)
V = ops.khatri_rao_fwd(xn, wn, d, w)
pV = ops.planes_of(V, Q)
pZ = ops.split_planes(Z, L, n, L)


def timeit(name, fn, pause, flops=None):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        if pause:
            time.sleep(pause)
    ms.sort()
    m = ms[len(ms) // 2]
    extra = f"  {flops / m / 1e9:7.1f} TFLOP/s executed" if flops else ""
    print(f"{name:44s} {'cool' if pause else 'hot '} {m:7.3f} ms (min {ms[0]:.3f}){extra}", flush=True)


fl_g = 3.0 * n * Q * (Q + 256)          # lower tiles incl. the diagonal ones in full: 136 tiles of 256 x 256, 3 MMAs
fl_c = 3.0 * 2 * n * Q * L
for pause in (0.05, 0.0):
    timeit("gram tiles (planes)", lambda: ops.gram_vtz_planes(pV, None, n, Q, 0), pause, fl_g)
    timeit("V^T Z (planes)", lambda: ops.atb_planes(pV, pZ, n, Q, L), pause, fl_c)
    timeit("gram + V^T Z in one launch", lambda: ops.gram_vtz_planes(pV, pZ, n, Q, L), pause, fl_g + fl_c)
    timeit("split Z", lambda: ops.split_planes(Z, L, n, L), pause)
    timeit("khatri_rao_fwd + planes", lambda: ops.khatri_rao_fwd(xn, wn, d, w), pause)
