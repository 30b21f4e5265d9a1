"""Pass 1 and pass 2 on operand planes, alone, at the c3 shape (or N Q-related args): the command the ncu captures of
the round wrap.  Usage: python planes_only.py [N] [reps]"""
import sys

import torch

sys.path.insert(0, ".")
import gppvae_b200  # noqa: E402
from gppvae_b200 import ops  # noqa: E402
from gppvae_b200.synth import make_problem  # noqa: E402

dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pr = make_problem(N, 256, 16, 256, seed=0, device=dev)
vm = gppvae_b200.Vmodel(pr.x0.shape[0], 16, 256, 16).to(dev)
with torch.no_grad():
    vm.x0.copy_(pr.x0); vm.v0.copy_(pr.v0)
    V = vm(pr.d, pr.w)
pV = ops.PLANES.get(V, 4096)
pX = ops.split_planes(pr.Z, 256, N, 256)
W = torch.randn(4096, 256, device=dev) / 64
scal = torch.zeros(8, device=dev, dtype=torch.float64); scal[1] = 0.5
torch.cuda.synchronize()
for i in range(reps):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    ops.gram_vtz_planes(pV, pX, N, 4096, 256)
    e1.record()
    ops.xb_nll_planes(pV, pr.Z, 256, W, N, 4096, 256, scal)
    e2.record()
    torch.cuda.synchronize()
    print(f"rep {i}: pass 1 {e0.elapsed_time(e1):.2f} ms, pass 2 {e1.elapsed_time(e2):.2f} ms")
