"""Timing of the producers of V and its operand planes at c3 shape (N x 4096): plain Khatri-Rao map, fused map + planes,
split of an existing fp32 matrix with / without the column sums of squares."""
import sys

import torch

sys.path.insert(0, ".")
import gppvae_b200  # noqa: E402
from gppvae_b200 import _lib, ops  # noqa: E402
from gppvae_b200.synth import make_tables  # noqa: E402

dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
p, q = 256, 16
Q = p * q
gen = torch.Generator(device=dev).manual_seed(0)
P = -(-N // q)
x0, v0 = make_tables(P, p, q, "trained", gen, dev)
xn = ops.normalize_rows_fwd(x0)
wn = ops.normalize_rows_fwd(v0)
perm = torch.randperm(N, device=dev)
d, w = (perm // q).contiguous(), (perm % q).contiguous()
lib = _lib.load()


def timeit(fn, name, nbytes):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name:34s} {ms:8.3f} ms   {nbytes / ms / 1e9:6.2f} TB/s")


V = torch.empty(N, Q, device=dev)
st = torch.cuda.current_stream().cuda_stream
timeit(lambda: lib.gpp_khatri_rao_fwd(xn.data_ptr(), P, p, wn.data_ptr(), q, q, d.data_ptr(), w.data_ptr(), N, V.data_ptr(), Q, st),
       "khatri_rao_fwd (fp32 only)", N * Q * 4)
buf = torch.empty(lib.gpp_planes_bytes(N, Q), dtype=torch.uint8, device=dev)
ws = torch.empty(lib.gpp_split_workspace_bytes(N, Q), dtype=torch.uint8, device=dev)
timeit(lambda: lib.gpp_khatri_rao_fwd_planes(xn.data_ptr(), P, p, wn.data_ptr(), q, q, d.data_ptr(), w.data_ptr(), N, V.data_ptr(), Q,
                                             buf.data_ptr(), buf.numel(), ws.data_ptr(), ws.numel(), st),
       "khatri_rao_fwd_planes (V + planes)", N * Q * 8)
timeit(lambda: lib.gpp_split_planes(V.data_ptr(), Q, N, Q, 1 | 2, buf.data_ptr(), buf.numel(), ws.data_ptr(), ws.numel(), st),
       "split_planes colsq, unit bound", N * Q * 8)
timeit(lambda: lib.gpp_split_planes(V.data_ptr(), Q, N, Q, 2, buf.data_ptr(), buf.numel(), ws.data_ptr(), ws.numel(), st),
       "split_planes no colsq, unit bound", N * Q * 8)
timeit(lambda: lib.gpp_split_planes(V.data_ptr(), Q, N, Q, 1, buf.data_ptr(), buf.numel(), ws.data_ptr(), ws.numel(), st),
       "split_planes colsq + absmax scan", N * Q * 12)
