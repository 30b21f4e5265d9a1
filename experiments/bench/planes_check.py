"""Operand-plane kernels against float64 torch on the GPU (checker only): pass 1 (Gram + V^T Z, exact diagonal),
pass 2 (Xb, nll partials) for a few shapes including ragged ones.  Usage: python planes_check.py [N Q L]..."""
import sys
import time

import torch

sys.path.insert(0, ".")
import gppvae_b200  # noqa: E402
from gppvae_b200 import ops  # noqa: E402
from gppvae_b200._lib import NSCAL, S_VN  # noqa: E402
from gppvae_b200.synth import make_problem  # noqa: E402

dev = torch.device("cuda:0")


def check(N, p, q, L, kind="trained"):
    Q = p * q
    pr = make_problem(N, p, q, L, kind=kind, lvs=(0.0, 0.0), seed=0, device=dev)
    vm = gppvae_b200.Vmodel(pr.x0.shape[0], q, p, q).to(dev)
    with torch.no_grad():
        vm.x0.copy_(pr.x0); vm.v0.copy_(pr.v0)
        V = vm(pr.d, pr.w)
    assert ops.planes_supported(N, Q, L), "planes not supported?"
    Vm, ldv = ops.as_matrix(V, "V")
    Zm, ldz = ops.as_matrix(pr.Z, "Z")
    Lk = Zm.shape[1]
    pV = ops.PLANES.get(Vm, ldv)
    print(f"N={N} Q={Q} L={L} kind={kind}: planes from KR: {pV is not None}")
    pV2 = ops.split_planes(Vm, ldv, N, Vm.shape[1], colsq=True)
    pX = ops.split_planes(Zm, ldz, N, Lk)
    torch.cuda.synchronize()
    V64, Z64 = Vm.double(), Zm.double()
    G64 = V64.t() @ V64
    C64 = V64.t() @ Z64
    for name, pl in (("kr-planes", pV), ("split-planes", pV2)):
        if pl is None:
            continue
        GC = ops.gram_vtz_planes(pl, pX, N, Vm.shape[1], Lk)
        torch.cuda.synchronize()
        G, C = GC[:, :Vm.shape[1]].double(), GC[:, Vm.shape[1]:].double()
        eg = ((G - G64).abs().max() / G64.abs().max()).item()
        ec = ((C - C64).abs().max() / C64.abs().max()).item()
        ed = ((G.diagonal() - G64.diagonal()).abs() / G64.diagonal().abs().clamp_min(1e-30)).max().item()
        sym = (G - G.t()).abs().max().item()
        print(f"  {name}: G err {eg:.2e}  C err {ec:.2e}  diag rel err {ed:.2e}  asym {sym:.1e}")
    # pass 2
    W = (torch.randn(Vm.shape[1], Lk, device=dev) * 0.05)
    scal = torch.zeros(NSCAL, device=dev, dtype=torch.float64)
    scal[S_VN] = 0.5
    Xb, nll = ops.xb_nll_planes(pV2, Zm, ldz, W, N, Vm.shape[1], Lk, scal)
    torch.cuda.synchronize()
    Xb64 = (Z64 - V64 @ W.double()) / 0.5
    ex = ((Xb.double() - Xb64).abs().max() / Xb64.abs().max()).item()
    quad64 = 0.5 * (Z64 * Xb64).sum(1)
    en = ((nll.double().view(-1) - quad64).abs().max() / quad64.abs().max()).item()
    print(f"  pass 2: Xb err {ex:.2e}  nll(quad) err {en:.2e}  xb2 rel {abs(scal[6].item() - (Xb64 ** 2).sum().item()) / (Xb64 ** 2).sum().item():.2e}")
    # timing
    for _ in range(2):
        ops.gram_vtz_planes(pV2, pX, N, Vm.shape[1], Lk)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.gram_vtz_planes(pV2, pX, N, Vm.shape[1], Lk)
    e1.record(); torch.cuda.synchronize()
    t1 = e0.elapsed_time(e1) / 3
    e0.record()
    for _ in range(3):
        ops.xb_nll_planes(pV2, Zm, ldz, W, N, Vm.shape[1], Lk, scal)
    e1.record(); torch.cuda.synchronize()
    t2 = e0.elapsed_time(e1) / 3
    fl = N * Q * (Q + 1) + 2.0 * N * Q * L
    print(f"  pass 1 {t1:.3f} ms ({fl / t1 / 1e9:.0f} algorithmic TFLOP/s), pass 2 {t2:.3f} ms")


if __name__ == "__main__":
    shapes = [(4005, 64, 9, 256, "trained"), (20000, 32, 16, 256, "init"), (100000, 64, 16, 256, "trained"),
              (3001, 24, 7, 100, "trained")]
    if len(sys.argv) > 1:
        a = sys.argv[1:]
        shapes = [(int(a[0]), int(a[1]), int(a[2]), int(a[3]), a[4] if len(a) > 4 else "trained")]
    for s in shapes:
        t0 = time.time()
        check(*s)
        print(f"  ({time.time() - t0:.1f}s)")
