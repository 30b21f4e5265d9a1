"""Grade and time the operand-plane kernels for a build variant: GPPVAE_LIB=path python planes_eval.py [N]
  * coherent bias of pass 1 on an all-positive block (mean signed relative error against float64);
  * relative NLL error of the full evaluation on small shapes (float64 oracle on the CPU);
  * pass 1 / pass 2 time at the c3 shape (N rows, Q=4096, L=256)."""
import os
import sys

sys.path.insert(0, ".")
lib = os.environ.get("GPPVAE_LIB")
if lib:
    import gppvae_b200._lib as L
    L.LIB_PATH = lib
import torch  # noqa: E402

import gppvae_b200  # noqa: E402
from gppvae_b200 import ops  # noqa: E402
from gppvae_b200.synth import make_problem  # noqa: E402
from oracle import gp_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
tag = os.path.basename(lib or "default")

n, Q, Lz = 20000, 1024, 256
torch.manual_seed(n)
V = torch.randn(n, Q, device=dev) * torch.rand(1, Q, device=dev)
V[:, : Q // 2] = V[:, : Q // 2].abs()
X = torch.randn(n, Lz, device=dev)
ref = V.double().t() @ torch.cat([V.double(), X.double()], 1)
GC = ops.gram_vtz_planes(ops.split_planes(V, Q, n, Q, colsq=True), ops.split_planes(X, Lz, n, Lz), n, Q, Lz)
err = ((GC.double() - ref).abs().max() / ref.abs().max()).item()
pos = ref[: Q // 2, : Q // 2]
rel = (GC[: Q // 2, : Q // 2].double() - pos) / pos
off = ~torch.eye(Q // 2, dtype=torch.bool, device=dev)
print(f"{tag}: positive block: max-rel err {err:.2e}  mean signed rel err off-diagonal {rel[off].mean().item():+.2e}")
del V, X, ref, GC

for (nn, p, q, Lz, lvs, kind) in [(1536, 16, 8, 64, (0.4, -0.6), "trained"), (4005, 64, 9, 256, (0.0, 0.0), "trained"),
                                  (4005, 64, 9, 256, (0.0, 0.0), "init"), (20000, 32, 16, 256, (2.0, -4.0), "trained"),
                                  (20000, 32, 16, 256, (0.0, 0.0), "init")]:
    pr = make_problem(nn, p, q, Lz, kind=kind, lvs=lvs, seed=1)
    V64 = O.feature_map(pr.x0.double(), pr.v0.double(), pr.d, pr.w)
    onll, oXb = O.nll_and_grad(pr.Z.double(), [V64], pr.lvs.double())
    vm = gppvae_b200.Vmodel(pr.x0.shape[0], q, p, q).to(dev)
    gp = gppvae_b200.GP().to(dev)
    with torch.no_grad():
        vm.x0.copy_(pr.x0.to(dev)); vm.v0.copy_(pr.v0.to(dev)); gp.lvs.copy_(pr.lvs.to(dev))
        Vd = vm(pr.d.to(dev), pr.w.to(dev))
        Xb, _, _, nll = gp.taylor_coeff(pr.Z.to(dev), [Vd], need_vb=False)
    e = (nll.double().sum().item() - onll.sum().item()) / abs(onll.sum().item())
    ex = ((Xb.cpu().double() - oXb).abs().max() / oXb.abs().max()).item()
    print(f"{tag}: n={nn} Q={p * q} L={Lz} {kind} lvs={lvs}: rel NLL err {e:+.2e}  Xb {ex:.2e}")

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
pr = make_problem(N, 256, 16, 256, seed=0, device=dev)
vm = gppvae_b200.Vmodel(pr.x0.shape[0], 16, 256, 16).to(dev)
with torch.no_grad():
    vm.x0.copy_(pr.x0); vm.v0.copy_(pr.v0)
    V = vm(pr.d, pr.w)
pV = ops.PLANES.get(V, 4096)
pX = ops.split_planes(pr.Z, 256, N, 256)


def timeit(fn, reps=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t1 = timeit(lambda: ops.gram_vtz_planes(pV, pX, N, 4096, 256))
W = torch.randn(4096, 256, device=dev) / 64
scal = torch.zeros(8, device=dev, dtype=torch.float64); scal[1] = 0.5
t2 = timeit(lambda: ops.xb_nll_planes(pV, pr.Z, 256, W, N, 4096, 256, scal))
fl1 = N * (4096 * 4097 + 2 * 4096 * 256.0)
print(f"{tag}: N={N} Q=4096: pass 1 {t1:.2f} ms ({fl1 / t1 / 1e9:.0f} algorithmic TFLOP/s), pass 2 {t2:.2f} ms")
