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
gt = np.diff(buf[:, 6]) / 1e3
print("step-to-step wall time (us, globaltimer at the entry of a panel CTA):")
print(" ".join(f"{x:.1f}" for x in gt))
print("sum over the 63 intervals: %.1f us" % gt.sum())
bw = np.zeros((64, 8), dtype=np.int64)
if lib.gpp_debug_chol_prof_w(bw.ctypes.data_as(ctypes.c_void_p)) == 0:
    print("wide CTA 0 (us after the panel CTA's entry): entry, main loop done, partial stored + counted, all chunks arrived, slice reduced | next step's entry")
    for j in (2, 8, 14, 20, 32, 40, 48, 56, 61):
        t0 = buf[j, 6]
        print(j, " ".join(f"{(bw[j, k] - t0) / 1e3:6.1f}" for k in (0, 1, 2, 4, 5)), f"| {(buf[j + 1, 6] - t0) / 1e3:6.1f}",
              f"| main loop {bw[j, 7] - bw[j, 6]} cycles = {(bw[j, 7] - bw[j, 6]) / max(1, bw[j, 1] - bw[j, 0]):.2f} GHz")
print("mean over steps 1..62:", d[1:63].mean(axis=0), "total", d[1:63].sum(axis=1).mean())
