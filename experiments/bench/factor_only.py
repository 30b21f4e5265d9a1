"""Run the Q-space factorisation a few times (for ncu --cache-control none timing of its kernels)."""
import os, sys, time
import torch
sys.path.insert(0, ".")
if os.environ.get("GPPVAE_LIB"):
    import gppvae_b200._lib as L
    L.LIB_PATH = os.environ["GPPVAE_LIB"]
from gppvae_b200 import ops
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
V = torch.randn(4 * Q, Q, device=dev) / (Q ** 0.5)
G = (V.t() @ V).contiguous()
C = torch.randn(Q, 256, device=dev)
vs = torch.tensor([0.5, 0.5], device=dev)
for _ in range(2):
    f = ops.factor(G, Q, Q, vs, False)
    ops.solve_w(f, C, 256, 256, 256, 100000)      # (warm-up of solve_w too: its first calls allocate)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    f = ops.factor(G, Q, Q, vs, False)
torch.cuda.synchronize()
t1 = time.perf_counter()
for _ in range(reps):
    W, sc = ops.solve_w(f, C, 256, 256, 256, 100000)
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"Q={Q}: factor {(t1 - t0) / reps * 1e3:.3f} ms, solve_w {(t2 - t1) / reps * 1e3:.3f} ms")
