"""One GPPVAE training epoch at BASELINE.json configs[0] shape (N=4005 images 128x128x3, L=256, 9 views, p=64) on one
GPU: stock-torch conv VAE (reported, not optimised) around the B200 GP term; phase split by CUDA events.

    python experiments/bench/epoch_bench.py [N]
"""
import sys, json
sys.path.insert(0, ".")
import torch
import gppvae_b200
from gppvae_b200.vae import FaceVAE
from gppvae_b200.epoch import train_epoch, eval_step

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4005
dev = torch.device("cuda:0")
torch.manual_seed(0)
q, p, L, bs = 9, 64, 256, 64
P = -(-N // q)
perm = torch.randperm(N, device=dev)
D, W = (perm // q), (perm % q)
Y = torch.rand(N, 3, 128, 128, device=dev)
Nv = max(bs, N // 8)
Yv = torch.rand(Nv, 3, 128, 128, device=dev)
Dv, Wv = torch.randint(0, P, (Nv,), device=dev), torch.randint(0, q, (Nv,), device=dev)
vae = FaceVAE().to(dev)
vm = gppvae_b200.Vmodel(P, q, p, q).to(dev)
gp = gppvae_b200.GP().to(dev)
vae_opt = torch.optim.Adam(vae.parameters(), lr=2e-4)
gp_opt = torch.optim.Adam(list(vm.parameters()) + list(gp.parameters()), lr=1e-3)
res = {}
for lazy in (False, True):
    for it in range(3):
        prof = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rv = train_epoch(vae, vm, gp, Y, D, W, vae_opt, gp_opt, bs=bs, lazy=lazy, profile=prof)
        e1.record(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        Zm = torch.randn(N, L, device=dev)
    ev0.record()
    ev = eval_step(vae, vm, gp, Yv, Dv, Wv, Zm, D, W, bs=bs, lazy=lazy)
    ev1.record(); torch.cuda.synchronize()
    res["structured" if lazy else "dense"] = dict(epoch_ms=e0.elapsed_time(e1), phases_ms=prof, eval_step_ms=ev0.elapsed_time(ev1),
                                                  loss=rv["loss"], gp_nll=rv["gp_nll"])
print(json.dumps(dict(N=N, bs=bs, q=q, p=p, L=L, image="3x128x128", **res)))
