#!/usr/bin/env python
"""Generate the golden vectors in this directory from the UNMODIFIED reference.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

It imports ``gp.py`` / ``vmod.py`` from /root/reference/pysrc/faceplace exactly
as they are, under the three-line shim SURVEY.md section 8(c) describes (stub
``h5py``; ``.cuda()`` -> identity because this container has no GPU), drives
the reference classes in float32 and again in float64, and stores inputs and
outputs as ``<case>.npz``.  The reference cannot travel to the GPU box, these
small fixtures can.  Nothing here is copied from the reference: the script only
*calls* it.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/pysrc/faceplace"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF)
    import gp as ref_gp      # noqa: E402
    import vmod as ref_vmod  # noqa: E402
    return ref_gp, ref_vmod


def np32(t):
    return t.detach().cpu().numpy()


def run_case(ref_gp, ref_vmod, name, *, x0=None, v0=None, d=None, w=None, Vdirect=None, Z, lvs, mb):
    """Drive the reference on one input set in fp32 and fp64; return dict of arrays."""
    out = {}
    lvs = np.asarray(lvs, np.float32)   # every input is float32-representable in both runs
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        torch.set_default_dtype(dt)
        gpm = ref_gp.GP(n_rand_effs=1)
        gpm.lvs.data[:] = torch.as_tensor(lvs, dtype=dt)
        if Vdirect is None:
            vm = ref_vmod.Vmodel(x0.shape[0], v0.shape[0], x0.shape[1], v0.shape[1])
            vm.x0.data[:] = torch.as_tensor(x0, dtype=dt)
            vm.v0.data[:] = torch.as_tensor(v0, dtype=dt)
            dd, ww = torch.as_tensor(d), torch.as_tensor(w)
            out[f"{tag}_xn"] = np32(vm.x())
            out[f"{tag}_wn"] = np32(vm.v())
            V = vm(dd, ww).detach()
            # gradient of a fixed linear functional of V wrt the raw tables (vmod.py:28-35 backward)
            probe = torch.as_tensor(np.cos(np.arange(V.numel(), dtype=np.float64)).reshape(V.shape), dtype=dt)
            (vm(dd, ww) * probe).sum().backward()
            out[f"{tag}_gx0"] = np32(vm.x0.grad)
            out[f"{tag}_gv0"] = np32(vm.v0.grad)
        else:
            V = torch.as_tensor(Vdirect, dtype=dt)
        X = torch.as_tensor(Z, dtype=dt)
        out[f"{tag}_V"] = np32(V)
        vs = gpm.get_vs()
        out[f"{tag}_vs"] = np32(vs)
        U, UBi, Shb = gpm.U_UBi_Shb([V], vs)
        out[f"{tag}_Shb"] = np32(Shb)
        out[f"{tag}_U"] = np32(U)
        out[f"{tag}_UBi"] = np32(UBi)
        out[f"{tag}_KiX"] = np32(gpm.solve(X, U, UBi, vs))
        Xb, Vbs, vbs, nll = gpm.taylor_coeff(X, [V])
        out[f"{tag}_Xb"], out[f"{tag}_Vb"] = np32(Xb), np32(Vbs[0])
        out[f"{tag}_vbs"], out[f"{tag}_nll"] = np32(vbs), np32(nll)
        out[f"{tag}_nll_attached"] = np32(gpm.nll(X, [V]))
        out[f"{tag}_nll_ineff"] = np32(gpm.nll_ineff(X, [V]))
        # Taylor surrogate on a minibatch (train_gppvae.py:279-293) and its gradients
        idx = torch.as_tensor(mb)
        xm = X[idx].clone().requires_grad_(True)
        vm_ = V[idx].clone().requires_grad_(True)
        gpm.lvs.grad = None
        te = gpm.taylor_expansion(xm, [vm_], Xb[idx], [Vbs[0][idx]], vbs)
        te.sum().backward()
        out[f"{tag}_te"] = np32(te)
        out[f"{tag}_te_gX"], out[f"{tag}_te_gV"] = np32(xm.grad), np32(vm_.grad)
        out[f"{tag}_te_glvs"] = np32(gpm.lvs.grad)
        # exact gradients of sum(nll) by autograd through svd/inverse (gp.py:205-214)
        xf = X.clone().requires_grad_(True)
        vf = V.clone().requires_grad_(True)
        gpm.lvs.grad = None
        gpm.nll(xf, [vf]).sum().backward()
        out[f"{tag}_nll_gX"], out[f"{tag}_nll_gV"] = np32(xf.grad), np32(vf.grad)
        out[f"{tag}_nll_glvs"] = np32(gpm.lvs.grad)
    torch.set_default_dtype(torch.float32)
    if Vdirect is None:
        out.update(x0=np.asarray(x0, np.float32), v0=np.asarray(v0, np.float32),
                   d=np.asarray(d, np.int64), w=np.asarray(w, np.int64))
    else:
        out.update(Vdirect=np.asarray(Vdirect, np.float32))
    out.update(Z=np.asarray(Z, np.float32), lvs=np.asarray(lvs, np.float32), mb=np.asarray(mb, np.int64))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, {len(out)} arrays")


def faceplace_like(rng, N, p, q, L, kind):
    """Synthetic inputs of SURVEY.md section 8(d), float32-representable."""
    P = -(-N // q)
    perm = rng.permutation(N)
    d = (np.arange(N) // q)[perm]
    w = (np.arange(N) % q)[perm]
    if kind == "init":       # vmod.py:37-40
        x0 = np.concatenate([np.ones((P, 1)), 1e-3 * rng.standard_normal((P, p - 1))], 1)
        v0 = np.eye(q) + 1e-3 * rng.standard_normal((q, q))
    else:                    # "trained-like"
        x0 = rng.standard_normal((P, p))
        v0 = np.eye(q) + 0.5 * rng.standard_normal((q, q))
    x0, v0 = x0.astype(np.float32), v0.astype(np.float32)
    xn = x0 / np.sqrt((x0 * x0).sum(1, keepdims=True))
    wn = v0 / np.sqrt((v0 * v0).sum(1, keepdims=True))
    V = (xn[d][:, :, None] * wn[w][:, None, :]).reshape(N, -1)
    Z = (0.5 * rng.standard_normal((N, L)) + V @ rng.standard_normal((p * q, L))).astype(np.float32)
    return x0, v0, d, w, Z


def main():
    ref_gp, ref_vmod = load_reference()
    rng = np.random.default_rng(20261018)

    # toy Vmodel of vmod.py:45-60 (P=Q=4, p=q=2, every object in two views)
    x0 = rng.standard_normal((4, 2)).astype(np.float32)
    v0 = (np.eye(4, 2) + 0.3 * rng.standard_normal((4, 2))).astype(np.float32)
    d = np.kron(np.arange(4), np.ones(2)).astype(np.int64)
    w = np.kron(np.ones(2), np.arange(4)).astype(np.int64)
    Z = rng.standard_normal((8, 4)).astype(np.float32)
    run_case(ref_gp, ref_vmod, "toy_vmod", x0=x0, v0=v0, d=d, w=w, Z=Z, lvs=[0.0, 0.0], mb=[1, 5, 6])

    x0, v0, d, w, Z = faceplace_like(rng, 90, 8, 9, 16, "init")
    run_case(ref_gp, ref_vmod, "faceplace_init", x0=x0, v0=v0, d=d, w=w, Z=Z, lvs=[0.0, 0.0],
             mb=rng.permutation(90)[:16])

    x0, v0, d, w, Z = faceplace_like(rng, 120, 6, 4, 12, "trained")
    run_case(ref_gp, ref_vmod, "faceplace_trained", x0=x0, v0=v0, d=d, w=w, Z=Z, lvs=[2.0, -4.0],
             mb=rng.permutation(120)[:16])

    # general V, in the manner of gp.py:138-157 (standardised binary design, graded signal)
    N, S, L = 160, 40, 24
    G = 1.0 * (rng.random((N, S)) < 0.2)
    G -= G.mean(0)
    G /= G.std(0) * np.sqrt(S)
    Zg = G @ rng.standard_normal((S, L))
    Zn = rng.standard_normal((N, L))
    vg = np.linspace(0.8, 0, L)
    Zg *= np.sqrt(vg / Zg.var(0))
    Zn *= np.sqrt((1 - vg) / Zn.var(0))
    run_case(ref_gp, ref_vmod, "genetics", Vdirect=G.astype(np.float32), Z=(Zg + Zn).astype(np.float32),
             lvs=[0.3, -0.2], mb=rng.permutation(N)[:32])


if __name__ == "__main__":
    main()
