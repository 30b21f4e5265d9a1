#!/usr/bin/env python
"""Golden vectors for one GPPVAE training epoch, produced by the UNMODIFIED reference classes.

    python tests/golden/make_golden_epoch.py          # in the build container (/root/reference mounted)

train_gppvae.py itself cannot be imported (module-level optparse / h5py / file side effects), so
this script re-drives its sequence -- encode_Y (:204-220), Eps / Z (:157-158), eval_step (:229-259),
taylor_coeff (:166-167), backprop_and_update (:264-311) and the two Adam steps -- on small synthetic tensors with
the reference's own `FaceVAE`, `Vmodel` and `GP` classes, imported as they are under the usual shim (stub `h5py`
and `pylab`, `.cuda()` -> identity), in float64 with float32-representable inputs.  The noise draw and the
minibatch order are stored so that the harness under test can replay them.  Nothing is copied from the reference:
the script only calls it.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/pysrc/faceplace"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    for name in ("h5py", "pylab"):
        sys.modules.setdefault(name, types.ModuleType(name))
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF)
    import gp as ref_gp
    import vae as ref_vae
    import vmod as ref_vmod
    return ref_gp, ref_vmod, ref_vae


def main():
    ref_gp, ref_vmod, ref_vae = load_reference()
    rng = np.random.RandomState(7)
    N, Nv, P, Q, p, L, bs = 40, 12, 10, 4, 6, 8, 16
    cfg = dict(img_size=32, nf=4, zdim=L, steps=3, colors=3, act="elu", vy=1e-3)
    perm = rng.permutation(N)
    D, W = (perm // Q).astype(np.int64), (perm % Q).astype(np.int64)
    Dv, Wv = rng.randint(0, P, Nv).astype(np.int64), rng.randint(0, Q, Nv).astype(np.int64)
    Y = rng.rand(N, 3, 32, 32).astype(np.float32)
    Yv = rng.rand(Nv, 3, 32, 32).astype(np.float32)
    Eps = rng.randn(N, L).astype(np.float32)
    order = rng.permutation(N)
    batches = [order[a:a + bs] for a in range(0, N, bs)]

    torch.manual_seed(3)
    torch.set_default_dtype(torch.float32)
    vae32 = ref_vae.FaceVAE(**cfg)
    vm32 = ref_vmod.Vmodel(P, Q, p, Q)
    with torch.no_grad():                      # move the tables away from the degenerate init so every gradient is exercised
        vm32.x0.add_(0.3 * torch.randn(P, p))
        vm32.v0.add_(0.3 * torch.randn(Q, Q))
    init = {f"vae.{k}": v.detach().numpy().copy() for k, v in vae32.state_dict().items()}
    init.update({"vm.x0": vm32.x0.detach().numpy().copy(), "vm.v0": vm32.v0.detach().numpy().copy(),
                 "gp.lvs": np.array([0.3, -0.2], np.float32)})

    torch.set_default_dtype(torch.float64)
    vae = ref_vae.FaceVAE(**cfg)
    vae.load_state_dict({k[4:]: torch.as_tensor(v, dtype=torch.float64) for k, v in init.items() if k.startswith("vae.")})
    vm = ref_vmod.Vmodel(P, Q, p, Q)
    gpm = ref_gp.GP(n_rand_effs=1)
    with torch.no_grad():
        vm.x0.copy_(torch.as_tensor(init["vm.x0"], dtype=torch.float64))
        vm.v0.copy_(torch.as_tensor(init["vm.v0"], dtype=torch.float64))
        gpm.lvs.copy_(torch.as_tensor(init["gp.lvs"], dtype=torch.float64))
    K = vae.K
    t64 = lambda a: torch.as_tensor(a, dtype=torch.float64)
    Yt, Yvt, Epst = t64(Y), t64(Yv), t64(Eps)
    Dt, Wt, Dvt, Wvt = (torch.as_tensor(a) for a in (D, W, Dv, Wv))
    vae_opt = torch.optim.Adam(vae.parameters(), lr=2e-4)
    gp_opt = torch.optim.Adam(list(vm.parameters()) + list(gpm.parameters()), lr=1e-3)
    out = {}

    # 1. encode_Y (:204-220), 2. sample Z (:157-158)
    vae.eval()
    with torch.no_grad():
        Zm, Zs = vae.encode(Yt)
    Z = Zm + Epst * Zs
    out["Zm"], out["Zs"] = Zm.numpy(), Zs.numpy()

    # 3. eval_step (:229-259)
    with torch.no_grad():
        Vt = vm(Dt, Wt).detach()
        Vv = vm(Dvt, Wvt).detach()
        vs = gpm.get_vs()
        U, UBi, _ = gpm.U_UBi_Shb([Vt], vs)
        Kiz = gpm.solve(Zm, U, UBi, vs)
        Zo = vs[0] * Vv.mm(Vt.transpose(0, 1).mm(Kiz))
        Zv = vae.encode(Yvt)[0]
        Yr, Yo = vae.decode(Zv), vae.decode(Zo)
        out["mse_out"] = float(((Yvt - Yo) ** 2).view(Nv, -1).mean(1).mean())
        out["mse_val"] = float(((Yvt - Yr) ** 2).view(Nv, -1).mean(1).mean())
        out["Zo"] = Zo.numpy()

    # 4. Taylor coefficients (:166-167)
    Zb, Vbs, vbs, gp_nll = gpm.taylor_coeff(Z, [Vt])
    out["gp_nll"] = float(gp_nll.mean()) / K

    # 5. backprop_and_update (:264-311) with the stored minibatch order
    vae_opt.zero_grad(); gp_opt.zero_grad()
    vae.train(); gpm.train(); vm.train()
    sums = np.zeros(3)
    for idx in batches:
        ix = torch.as_tensor(idx)
        y = Yt[ix]
        zm, zs = vae.encode(y)
        z = zm + zs * Epst[ix]
        yr = vae.decode(z)
        recon_term, mse = vae.nll(y, yr)
        gp_nll_fo = gpm.taylor_expansion(z, [vm(Dt[ix], Wt[ix])], Zb[ix], [Vbs[0][ix]], vbs) / K
        pen_term = -0.5 * zs.sum(1)[:, None] / K
        (recon_term + gp_nll_fo + pen_term).sum().backward()
        sums += np.array([float(mse.sum()), float(recon_term.sum()), float(pen_term.sum())]) / N
    out["mse"], out["recon_term"], out["pen_term"] = sums
    out["loss"] = out["recon_term"] + out["gp_nll"] + out["pen_term"]
    grads = {f"grad.vae.{k}": v.grad.numpy().copy() for k, v in vae.named_parameters() if v.grad is not None}
    grads.update({"grad.vm.x0": vm.x0.grad.numpy().copy(), "grad.vm.v0": vm.v0.grad.numpy().copy(),
                  "grad.gp.lvs": gpm.lvs.grad.numpy().copy()})
    vae_opt.step(); gp_opt.step()
    after = {"after.vm.x0": vm.x0.detach().numpy().copy(), "after.gp.lvs": gpm.lvs.detach().numpy().copy(),
             "after.vae.dense_zm.weight": vae.dense_zm.weight.detach().numpy().copy()}

    np.savez_compressed(os.path.join(HERE, "epoch", "epoch_small.npz"), Y=Y, Yv=Yv, D=D, W=W, Dv=Dv, Wv=Wv, Eps=Eps,
                        order=order, bs=np.int64(bs), cfg_img_size=np.int64(32), cfg_nf=np.int64(4), cfg_zdim=np.int64(L),
                        cfg_steps=np.int64(3), P=np.int64(P), Q=np.int64(Q), p=np.int64(p),
                        **{f"init.{k}": v for k, v in init.items()}, **{f"out.{k}": np.asarray(v) for k, v in out.items()},
                        **grads, **after)
    print("wrote epoch_small.npz:", {k: (v if np.isscalar(v) else np.asarray(v).shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
