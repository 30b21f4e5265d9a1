"""Role profile of the tcgen05 pass-1 kernel (library built with -DGPP_TC_PROF): where each warp role waits.
GPPVAE_LIB=experiments/bench/variants/lib_w4g2s6prof.so python experiments/bench/prof_roles.py [N Q]"""
import ctypes, os, sys
sys.path.insert(0, ".")
import gppvae_b200._lib as L
L.LIB_PATH = os.environ["GPPVAE_LIB"]
import numpy as np, torch
from gppvae_b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda:0")
V = torch.randn(N, Q, device=dev); Z = torch.randn(N, 256, device=dev)
for _ in range(2):
    ops.gram_vtz(V, Q, Z, 256, N, Q, 256)
torch.cuda.synchronize()
lib = ctypes.CDLL(L.LIB_PATH)
buf = np.zeros((512, 16), dtype=np.uint64)
assert lib.gpp_debug_prof(buf.ctypes.data_as(ctypes.c_void_p)) == 0
names = ["producer (wait empty)", "mma (wait tempty | wait conv)", "A converter (wait full | wait lo_empty)", "drain (wait tfull)", "B converter (wait full | wait lo_empty)"]
for rank in (0, 1):
    rows = buf[rank::2][:74].astype(np.float64)
    for r, nm in enumerate(names):
        w0, w1, tot = rows[:, 3 * r], rows[:, 3 * r + 1], rows[:, 3 * r + 2]
        ok = tot > 0
        if not ok.any():
            continue
        print(f"rank {rank} {nm:34s}: total {tot[ok].mean() / 1e6:8.2f} Mclk  wait0 {100 * (w0[ok] / tot[ok]).mean():5.1f} %  "
              f"wait1 {100 * (w1[ok] / tot[ok]).mean():5.1f} %")
