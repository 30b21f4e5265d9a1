"""BASELINE.json configs[3]: one full GPPVAE training epoch at N = 100k synthetic faces (3 x 128 x 128, L = 256, 9 views,
p = 64, bs = 64) on the GPUs of one box, rows (images) sharded over the ranks: stock-torch conv VAE (reported, not
optimised) around the B200 GP term, structured and dense route, phase split by CUDA events (max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        experiments/bench/epoch_dist.py [N] > gpurun_out/epoch_100k.json
"""
import json
import os
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gppvae_b200  # noqa: E402
from gppvae_b200.epoch import eval_step, make_vt, train_epoch  # noqa: E402
from gppvae_b200.vae import FaceVAE  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
group = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
q, p, L, bs = 9, 64, 256, 64
P = -(-N // q)
torch.manual_seed(0)                                  # same parameters and the same global permutation on every rank
perm = torch.randperm(N, device=dev)
per = -(-N // world)
rows = perm[rank * per: min(N, (rank + 1) * per)]
n = rows.numel()
D, W = (rows // q).contiguous(), (rows % q).contiguous()
vae = FaceVAE().to(dev)
vm = gppvae_b200.Vmodel(P, q, p, q).to(dev)
gp = gppvae_b200.GP().to(dev)
gen = torch.Generator(device=dev).manual_seed(1000 + rank)
Y = torch.rand(n, 3, 128, 128, device=dev, generator=gen)
Nv = max(bs, n // 8)
Yv = torch.rand(Nv, 3, 128, 128, device=dev, generator=gen)
Dv = torch.randint(0, P, (Nv,), device=dev, generator=gen)
Wv = torch.randint(0, q, (Nv,), device=dev, generator=gen)
vae_opt = torch.optim.Adam(vae.parameters(), lr=2e-4)
gp_opt = torch.optim.Adam(list(vm.parameters()) + list(gp.parameters()), lr=1e-3)


def max_over_ranks(d):
    keys = sorted(d)
    t = torch.tensor([d[k] for k in keys], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {k: float(v) for k, v in zip(keys, t)}


res = {}
for lazy in (True, False):
    for it in range(2):                               # first epoch warms cuDNN autotuning and the allocator
        prof = {}
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        Vt = make_vt(vm, D, W, lazy)
        with torch.no_grad():
            Zm = torch.randn(n, L, device=dev, generator=gen)
        ev = eval_step(vae, vm, gp, Yv, Dv, Wv, Zm, D, W, bs=bs, lazy=lazy, Vt=Vt)
        e1.record()
        rv = train_epoch(vae, vm, gp, Y, D, W, vae_opt, gp_opt, bs=bs, lazy=lazy, profile=prof, group=group, n_total=N, Vt=Vt)
        e2.record()
        torch.cuda.synchronize()
    t = max_over_ranks(dict(prof, eval_step=e0.elapsed_time(e1), train_epoch=e1.elapsed_time(e2), epoch=e0.elapsed_time(e2)))
    res["structured" if lazy else "dense"] = dict(ms=t, loss=rv["loss"], gp_nll=rv["gp_nll"], mse=rv["mse"],
                                                  cache_hits=gp.cache_hits)
    gp.invalidate_cache()
    torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps(dict(config="BASELINE.json configs[3]", N=N, n_gpus=world, rows_per_rank=per, bs=bs, q=q, p=p, Q=p * q, L=L,
                          image="3x128x128", images_per_s={k: N / (v["ms"]["epoch"] * 1e-3) for k, v in res.items()},
                          gp_term_share={k: v["ms"]["gp_term"] / v["ms"]["epoch"] for k, v in res.items()}, **res)))
if world > 1:
    dist.destroy_process_group()
