"""Phase profile of chol_step_kernel (library built with -DGPP_CHOL_PROF)."""
import ctypes, os, sys
sys.path.insert(0, ".")
import gppvae_b200._lib as L
L.LIB_PATH = os.environ["GPPVAE_LIB"]
import numpy as np, torch
from gppvae_b200 import ops
Q = 4096
dev = torch.device("cuda:0")
V = torch.randn(4 * Q, Q, device=dev) / Q ** 0.5
G = (V.t() @ V).contiguous()
vs = torch.tensor([0.5, 0.5], device=dev)
for _ in range(2):
    f = ops.factor(G, Q, Q, vs, False)
torch.cuda.synchronize()
lib = ctypes.CDLL(L.LIB_PATH)
buf = np.zeros((64, 8), dtype=np.int64)
assert lib.gpp_debug_chol_prof(buf.ctypes.data_as(ctypes.c_void_p)) == 0
d = np.diff(buf[:, :6], axis=1)
print("phase cycles (load, narrow update, factor64, panel solve, store) for steps 1, 10, 30, 60:")
for j in (1, 10, 30, 60):
    print(j, d[j])
print("mean over steps 1..62:", d[1:63].mean(axis=0), "total", d[1:63].sum(axis=1).mean())
