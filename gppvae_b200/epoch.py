"""One GPPVAE training epoch around the B200 GP term: the sequence of
/root/reference/pysrc/faceplace/train_gppvae.py:151-190, 204-220, 223-261, 264-311 on stock torch + gppvae_b200.

Same mathematics and the same one-optimizer-step-per-epoch schedule; what changes is the orchestration the SURVEY
flags as dominating the epoch once the GP term is fast:

  * `Eps` is drawn on the device (the reference draws it on the CPU and copies it, :157);
  * per-minibatch metrics stay on the device and are read once per epoch (the reference syncs three times per
    minibatch, :304-306);
  * `Vt` can be kept in factored form (`Vmodel.lazy`); build it ONCE per epoch with `make_vt` and pass it to
    both `eval_step` and `train_epoch` (`Vt=`): the factorisation the evaluation step builds (:235) is then the one
    `taylor_coeff` finds in the cache (:166 factors the same (Vt, vs)) -- two separately built objects never match;
  * rows (images) shard over ranks: every rank encodes / decodes its own images, the GP term all-reduces its small
    Q-space partials (`GP.shard_rows`), gradients are all-reduced once per epoch, before the single optimiser step.

`eps` and `batches` can be passed in to replay a fixed noise draw and minibatch order (parity tests).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import torch


def encode_all(vae, Y: torch.Tensor, bs: int, device) -> tuple:
    """Posterior means and scales of all images, in minibatches, without grad (train_gppvae.py:204-220)."""
    was_training = vae.training
    vae.eval()
    n = Y.shape[0]
    zm_all, zs_all = [], []
    with torch.no_grad():
        for a in range(0, n, bs):
            zm, zs = vae.encode(Y[a:a + bs].to(device, non_blocking=True))
            zm_all.append(zm)
            zs_all.append(zs)
    vae.train(was_training)
    return torch.cat(zm_all, 0), torch.cat(zs_all, 0)


def make_batches(n: int, bs: int, device, generator: Optional[torch.Generator] = None) -> List[torch.Tensor]:
    """A shuffled partition of range(n) into minibatches (DataLoader(shuffle=True), train_gppvae.py:119)."""
    perm = torch.randperm(n, device=device, generator=generator)
    return [perm[a:a + bs] for a in range(0, n, bs)]


def make_vt(vm, D: torch.Tensor, W: torch.Tensor, lazy: bool = True):
    """`Vt = vm(Dt, Wt).detach()` of train_gppvae.py:161, dense or in factored form; hand the SAME object to eval_step and
    train_epoch so that the second use finds the first one's factorisation."""
    with torch.no_grad():
        return vm.lazy(D, W) if lazy else vm(D, W).detach()


def _all_reduce_grads(params: Iterable[torch.nn.Parameter], group) -> None:
    import torch.distributed as dist
    for prm in params:
        if prm.grad is not None:
            dist.all_reduce(prm.grad, op=dist.ReduceOp.SUM, group=group)


def train_epoch(vae, vm, gp, Y: torch.Tensor, D: torch.Tensor, W: torch.Tensor, vae_optimizer, gp_optimizer, bs: int = 64,
                eps: Optional[torch.Tensor] = None, batches: Optional[Sequence[torch.Tensor]] = None, lazy: bool = True,
                n_total: Optional[int] = None, group=None, generator: Optional[torch.Generator] = None,
                step: bool = True, profile: Optional[dict] = None, Vt=None) -> Dict[str, float]:
    """Steps 1, 2, 4, 5 of the epoch (train_gppvae.py:153-186); step 3 is `eval_step` below.

    Y (n x C x H x W; device or pinned host memory), D / W (n,) int64 on the device are THIS rank's rows; `n_total`
    is the number of rows over all ranks (default: n) and `group` the process group when rows are sharded (the GP is
    told both here: `gp.shard_rows(group, n_total=n_total)`).  `Vt`: the epoch's `make_vt(...)` when `eval_step` already
    used it.  Returns the epoch's metrics as Python floats (one sync).
    `profile`, when a dict, receives the device time of the phases in ms (encode, gp_term, minibatches, update)."""
    device = D.device
    marks = []

    def mark(name):
        if profile is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    mark("start")
    n = Y.shape[0]
    n_total = n if n_total is None else n_total
    K = float(vae.K)
    if group is not None:
        gp.shard_rows(group, n_total=n_total)

    # 1. encode, 2. sample                                                          (:154-158)
    Zm, Zs = encode_all(vae, Y, bs, device)
    if eps is None:
        eps = torch.randn(Zs.shape, device=device, generator=generator)
    Z = Zm + eps * Zs
    mark("encode")

    # 4. Taylor coefficients of the GP term over all rows                           (:161, :166-167)
    if Vt is None:
        Vt = make_vt(vm, D, W, lazy)
    Zb, Vbs, vbs, gp_nll = gp.taylor_coeff(Z, [Vt])
    gp_nll_sum = gp_nll.sum()
    mark("gp_term")

    # 5. accumulate gradients over minibatches, one optimiser step                  (:264-311)
    vae_optimizer.zero_grad()
    gp_optimizer.zero_grad()
    vae.train(); gp.train(); vm.train()
    if batches is None:
        batches = make_batches(n, bs, device, generator)
    acc = torch.zeros(3, device=device, dtype=torch.float64)      # sums of mse, recon_term, pen_term
    for idx in batches:
        y = Y[idx.to(Y.device)].to(device, non_blocking=True)
        zm, zs = vae.encode(y)
        z = zm + zs * eps[idx]
        recon_term, mse = vae.nll(y, vae.decode(z))
        V_mb = vm(D[idx], W[idx])                                                   # with grad to x0, v0 (:292)
        gp_nll_fo = gp.taylor_expansion(z, [V_mb], Zb[idx], [Vbs[0][idx]], vbs) / K  # (:293)
        pen_term = -0.5 * zs.sum(1, keepdim=True) / K                               # (:296)
        (recon_term + gp_nll_fo + pen_term).sum().backward()
        acc += torch.stack([mse.detach().sum(), recon_term.detach().sum(), pen_term.detach().sum()]).double()
    mark("minibatches")
    if group is not None:
        import torch.distributed as dist
        _all_reduce_grads(list(vae.parameters()) + list(vm.parameters()) + list(gp.parameters()), group)
        red = torch.cat([acc, gp_nll_sum.double().reshape(1),
                         torch.tensor([float(len(batches))], device=device, dtype=torch.float64)])
        dist.all_reduce(red, group=group)
        acc, gp_nll_sum = red[:3], red[3]
        # gp.py:131-132 adds <vbs, vs> once per MINIBATCH, so the reference's lvs gradient is ceil(N / bs) times J^T vbs
        # (SURVEY section 9); the ranks together ran sum_r ceil(n_r / bs) minibatches, which differs whenever the shards
        # do not divide evenly -- bring the accumulated lvs gradient back to the reference's count
        if gp.lvs.grad is not None:
            gp.lvs.grad.mul_(float(-(-n_total // bs)) / red[4].to(gp.lvs.grad.dtype))
    if step:
        vae_optimizer.step()
        gp_optimizer.step()

    mark("update")
    out = torch.cat([acc / n_total, (gp_nll_sum.double() / n_total / K).reshape(1)]).tolist()    # the epoch's one sync
    if profile is not None:
        for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
            profile[name] = e0.elapsed_time(e1)
    rv = {"mse": out[0], "recon_term": out[1], "pen_term": out[2], "gp_nll": out[3]}
    rv["loss"] = rv["recon_term"] + rv["gp_nll"] + rv["pen_term"]                   # (:184-186)
    return rv


def eval_step(vae, vm, gp, Yv: torch.Tensor, Dv: torch.Tensor, Wv: torch.Tensor, Zm: torch.Tensor, D: torch.Tensor,
              W: torch.Tensor, bs: int = 64, lazy: bool = True, Vt=None) -> Dict[str, float]:
    """Out-of-sample prediction of the validation latents through the GP and the two reconstruction errors
    (train_gppvae.py:223-261): Zo = v0 Vv (Vt^T K^-1 Zm); mse_out decodes Zo, mse_val decodes the encoder's own code."""
    device = D.device
    with torch.no_grad():
        vs = gp.get_vs()
        if Vt is None:
            Vt = make_vt(vm, D, W, lazy)
        U, UBi, _ = gp.U_UBi_Shb([Vt], vs)
        Kiz = gp.solve(Zm, U, UBi, vs)
        VtKiz = Vt.t().mm(Kiz)
        gp._all_reduce(VtKiz)          # rows sharded over ranks: V^T (K^-1 Zm) sums over all rows
        Zo = vs[0] * vm(Dv, Wv).mm(VtKiz)
        was_training = vae.training
        vae.eval()
        acc = torch.zeros(2, device=device, dtype=torch.float64)
        for a in range(0, Yv.shape[0], bs):
            y = Yv[a:a + bs].to(device, non_blocking=True)
            yr = vae.decode(vae.encode(y)[0])
            yo = vae.decode(Zo[a:a + bs])
            acc += torch.stack([((y - yo) ** 2).mean(), ((y - yr) ** 2).mean()]).double() * y.shape[0]
        vae.train(was_training)
        out = (acc / Yv.shape[0]).tolist()
        return {"mse_out": out[0], "mse_val": out[1], "vars": vs.tolist()}
