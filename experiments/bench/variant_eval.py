"""Time and grade pass 1 / pass 2 for a given build variant: GPPVAE_LIB=path python variant_eval.py [big]"""
import os, sys
sys.path.insert(0, ".")
lib = os.environ.get("GPPVAE_LIB")
if lib:
    import gppvae_b200._lib as L
    L.LIB_PATH = lib
import torch
from gppvae_b200 import ops
from gppvae_b200.synth import make_problem
dev = torch.device("cuda:0")
tag = os.path.basename(lib or "default") + (" v1" if os.environ.get("GPP_TC_V1") == "1" else "")

def timeit(fn, reps=3):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

# accuracy on a same-sign block (exposes coherent bias), as tests/test_gpu_parity.py does
n, Q, Lz = 20000, 1024, 256
torch.manual_seed(n)
V = torch.randn(n, Q, device=dev) * torch.rand(1, Q, device=dev)
V[:, : Q // 2] = V[:, : Q // 2].abs()
X = torch.randn(n, Lz, device=dev)
ref = V.double().t() @ torch.cat([V.double(), X.double()], 1)
GC = ops.gram_vtz(V, Q, X, Lz, n, Q, Lz)
err = ((GC.double() - ref).abs().max() / ref.abs().max()).item()
pos = ref[: Q // 2, : Q // 2]
bias = ((GC[: Q // 2, : Q // 2].double() - pos) / pos).mean().item()
print(f"{tag}: positive-block test max-rel err {err:.2e}  mean signed rel (positive block) {bias:.2e}")
del V, X, ref, GC

sizes = [(100_000, 64, 16, 256)] + ([(500_000, 256, 16, 256)] if "big" in sys.argv else [])
for (N, p, q, Lz) in sizes:
    pr = make_problem(N, p, q, Lz, seed=0, device=dev)
    Q = p * q
    xn = ops.normalize_rows_fwd(pr.x0); wn = ops.normalize_rows_fwd(pr.v0)
    V = ops.khatri_rao_fwd(xn, wn, pr.d, pr.w)
    ms1 = timeit(lambda: ops.gram_vtz(V, Q, pr.Z, Lz, N, Q, Lz))
    GC = ops.gram_vtz(V, Q, pr.Z, Lz, N, Q, Lz)
    idx = torch.arange(0, Q, 37, device=dev)
    ref = V[:, idx].double().t() @ torch.cat([V.double()[:, :512], pr.Z.double()], 1)
    got = torch.cat([GC[idx][:, :512], GC[idx][:, Q:]], 1).double()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    W = torch.randn(Q, Lz, device=dev) / Q ** 0.5
    ms2 = timeit(lambda: ops.x_minus_am(pr.Z, Lz, V, Q, W, Lz, N, Q, Lz, 1.0))
    fl1 = N * (Q * (Q + 1) + 2 * Q * Lz) * 3 / 1e12
    fl2 = N * 2 * Q * Lz * 3 / 1e12
    print(f"{tag}: N={N} Q={Q}: pass1 {ms1:.3f} ms ({fl1 / ms1 * 1e3:.0f} TF/s executed)  err {err:.2e} | "
          f"rows GEMM {ms2:.3f} ms ({fl2 / ms2 * 1e3:.0f} TF/s executed)")
    del V, GC, pr
