"""Time and grade pass 1 for a given build variant: GPPVAE_LIB=path python variant_eval.py"""
import os, sys, shutil
sys.path.insert(0, ".")
lib = os.environ.get("GPPVAE_LIB")
if lib:
    import gppvae_b200._lib as L
    L.LIB_PATH = lib
import torch
from gppvae_b200 import ops
from gppvae_b200.synth import make_problem
dev = torch.device("cuda:0")
for (N, p, q, Lz) in [(100_000, 64, 16, 256), (500_000, 256, 16, 256)]:
    pr = make_problem(N, p, q, Lz, seed=0, device=dev)
    Q = p * q
    xn = ops.normalize_rows_fwd(pr.x0); wn = ops.normalize_rows_fwd(pr.v0)
    V = ops.khatri_rao_fwd(xn, wn, pr.d, pr.w)
    for _ in range(2):
        GC = ops.gram_vtz(V, Q, pr.Z, Lz, N, Q, Lz)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        GC = ops.gram_vtz(V, Q, pr.Z, Lz, N, Q, Lz)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    idx = torch.arange(0, Q, 37, device=dev)
    ref = V[:, idx].double().t() @ torch.cat([V.double()[:, :512], pr.Z.double()], 1)
    got = torch.cat([GC[idx][:, :512], GC[idx][:, Q:]], 1).double()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    bias = ((got - ref) / ref.abs().clamp_min(1e-3 * ref.abs().max())).mean().item()
    print(f"{os.path.basename(lib or 'default')}: N={N} Q={Q}: pass1 {ms:.3f} ms  max-rel err {err:.2e}  mean signed rel {bias:.2e}")
