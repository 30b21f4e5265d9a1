"""Smoke-sized shapes through every hand-written tensor-core / Cholesky kernel, for compute-sanitizer
(racecheck / synccheck; one tool per run):

    compute-sanitizer --tool racecheck --kernel-name regex=pl_|tc_|chol_ python experiments/bench/sanitize_small.py

Shapes: N=1536 Q=128 L=64 (planes pass 1 / pass 2, fp32-entry pass 1 / rows kernel, Vb) and a Q=576 factorisation (nine
Cholesky panel steps with both roles) -- results are checked against float64 torch so that a run under the tool is also a
correctness run."""
import sys

import torch

sys.path.insert(0, ".")
import gppvae_b200  # noqa: E402
from gppvae_b200 import ops  # noqa: E402
from gppvae_b200.synth import make_problem  # noqa: E402

dev = torch.device("cuda:0")
pr = make_problem(1536, 16, 8, 64, kind="trained", lvs=(0.4, -0.6), seed=1, device=dev)
vm = gppvae_b200.Vmodel(pr.x0.shape[0], 8, 16, 8).to(dev)
gp = gppvae_b200.GP().to(dev)
with torch.no_grad():
    vm.x0.copy_(pr.x0); vm.v0.copy_(pr.v0); gp.lvs.copy_(pr.lvs)
    V = vm(pr.d, pr.w)
    Xb, Vbs, vbs, nll = gp.taylor_coeff(pr.Z, [V])                      # planes: pass 1, pass 2, Vb
    GC = ops.gram_vtz(V, 128, pr.Z, 64, 1536, 128, 64)                   # fp32 entry: tc_pass1_kernel
    W = torch.randn(128, 64, device=dev) * 0.1
    R = ops.x_minus_am(pr.Z, 64, V, 128, W, 64, 1536, 128, 64, 1.0)      # fp32 entry: tc_rows_kernel
torch.cuda.synchronize()
ref = V.double().t() @ torch.cat([V.double(), pr.Z.double()], 1)
print("pass 1 (fp32 entry) err", float((GC.double() - ref).abs().max() / ref.abs().max()))
print("rows (fp32 entry) err", float((R.double() - (pr.Z.double() - V.double() @ W.double())).abs().max()))
G = gp._cache.G[:, :128].double()
print("pass 1 (planes) err", float((G - ref[:, :128]).abs().max() / ref.abs().max()))

Q = 576
torch.manual_seed(0)
A = torch.randn(4 * Q, Q, device=dev) / Q ** 0.5
Gq = (A.t() @ A).contiguous()
vs = torch.tensor([0.5, 0.5], device=dev)
f = ops.factor(Gq, Q, Q, vs, True)
torch.cuda.synchronize()
B = torch.eye(Q, device=dev, dtype=torch.float64) + Gq.double()
print("Binv err", float((f.Binv.double() - torch.linalg.inv(B)).abs().max()))
print("done")
