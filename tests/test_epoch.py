"""the conv VAE restatement (stock torch) and the epoch harness of gppvae_b200/epoch.py against
golden vectors produced by the UNMODIFIED reference classes driven through train_gppvae.py's sequence
(tests/golden/make_golden_epoch.py; float64 reference, float32-representable inputs)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, rel_err


def _golden():
    with np.load(os.path.join(GOLDEN_DIR, "epoch", "epoch_small.npz")) as z:
        return {k: z[k] for k in z.files}


def _vae(g, dtype, device="cpu"):
    from gppvae_b200.vae import FaceVAE
    vae = FaceVAE(img_size=int(g["cfg_img_size"]), nf=int(g["cfg_nf"]), zdim=int(g["cfg_zdim"]), steps=int(g["cfg_steps"]))
    sd = {k[len("init.vae."):]: torch.as_tensor(v) for k, v in g.items() if k.startswith("init.vae.")}
    assert set(sd) == set(vae.state_dict()), "state_dict keys differ from the reference's FaceVAE"
    vae.load_state_dict(sd)
    return vae.to(device=device, dtype=dtype)


def test_facevae_matches_reference_cpu():
    """encode() of the restated FaceVAE with the reference's weights reproduces the reference's Zm, Zs (float64)."""
    g = _golden()
    vae = _vae(g, torch.float64).eval()
    with torch.no_grad():
        zm, zs = vae.encode(torch.as_tensor(g["Y"], dtype=torch.float64))
        elbo, mse, nll, kld = vae(torch.as_tensor(g["Y"], dtype=torch.float64), torch.as_tensor(g["Eps"], dtype=torch.float64))
    assert rel_err(zm, g["out.Zm"]) < 1e-12 and rel_err(zs, g["out.Zs"]) < 1e-12
    assert elbo.shape == (40, 1) and torch.isfinite(elbo).all() and torch.allclose(elbo, nll + kld)


@pytest.mark.gpu
@pytest.mark.parametrize("lazy", [True, False])
def test_epoch_against_reference_sequence(lazy):
    """eval_step + train_epoch on cuda:0 (fp32 kernels) against the reference sequence in float64: metrics, the
    accumulated gradients of every parameter before the optimiser step, and parameters after it."""
    import gppvae_b200
    from gppvae_b200.epoch import eval_step, train_epoch
    dev = torch.device("cuda:0")
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        g = _golden()
        vae = _vae(g, torch.float32, dev)
        vm = gppvae_b200.Vmodel(int(g["P"]), int(g["Q"]), int(g["p"]), int(g["Q"])).to(dev)
        gp = gppvae_b200.GP().to(dev)
        with torch.no_grad():
            vm.x0.copy_(torch.as_tensor(g["init.vm.x0"])); vm.v0.copy_(torch.as_tensor(g["init.vm.v0"]))
            gp.lvs.copy_(torch.as_tensor(g["init.gp.lvs"]))
        Y, Yv, Eps = (torch.as_tensor(g[k]).to(dev) for k in ("Y", "Yv", "Eps"))
        D, W, Dv, Wv = (torch.as_tensor(g[k]).to(dev) for k in ("D", "W", "Dv", "Wv"))
        bs = int(g["bs"])
        order = torch.as_tensor(g["order"]).to(dev)
        batches = [order[a:a + bs] for a in range(0, order.numel(), bs)]
        vae_opt = torch.optim.Adam(vae.parameters(), lr=2e-4)
        gp_opt = torch.optim.Adam(list(vm.parameters()) + list(gp.parameters()), lr=1e-3)

        from gppvae_b200.epoch import encode_all
        Zm, Zs = encode_all(vae, Y, bs, dev)
        assert rel_err(Zm.cpu(), g["out.Zm"]) < 1e-4 and rel_err(Zs.cpu(), g["out.Zs"]) < 1e-4
        ev = eval_step(vae, vm, gp, Yv, Dv, Wv, Zm, D, W, bs=bs, lazy=lazy)
        assert abs(ev["mse_out"] - float(g["out.mse_out"])) < 1e-4 * float(g["out.mse_out"])
        assert abs(ev["mse_val"] - float(g["out.mse_val"])) < 1e-4 * float(g["out.mse_val"])

        rv = train_epoch(vae, vm, gp, Y, D, W, vae_opt, gp_opt, bs=bs, eps=Eps, batches=batches, lazy=lazy, step=False)
        for key in ("mse", "recon_term", "pen_term", "gp_nll", "loss"):
            assert abs(rv[key] - float(g["out." + key])) < 2e-4 * abs(float(g["out." + key])), key
        worst = 0.0
        for name, prm in list(vae.named_parameters()):
            if prm.grad is not None:
                worst = max(worst, rel_err(prm.grad.cpu(), g["grad.vae." + name]))
        e_x0, e_v0 = rel_err(vm.x0.grad.cpu(), g["grad.vm.x0"]), rel_err(vm.v0.grad.cpu(), g["grad.vm.v0"])
        e_lvs = rel_err(gp.lvs.grad.cpu(), g["grad.gp.lvs"])
        print(f"[epoch lazy={lazy}] grads vs reference (fp64): vae {worst:.2e}  x0 {e_x0:.2e}  v0 {e_v0:.2e}  lvs {e_lvs:.2e}")
        assert worst < 2e-3 and e_x0 < 2e-3 and e_v0 < 2e-3 and e_lvs < 2e-3
        vae_opt.step(); gp_opt.step()
        # Adam's first step moves every weight by lr * sign(grad): compare the displacement, not the sign-fragile value
        assert rel_err((vm.x0.detach().cpu() - torch.as_tensor(g["init.vm.x0"])),
                       g["after.vm.x0"] - g["init.vm.x0"].astype(np.float64)) < 5e-2
        assert rel_err(gp.lvs.detach().cpu(), g["after.gp.lvs"]) < 1e-4
    finally:
        torch.backends.cudnn.allow_tf32 = tf32


# ---- the same epoch on CPU through the test-only stand-in engine (tests/fake_engine.py): covers the HOST logic of
# epoch.py -- sequencing, minibatch indexing, the lazy route, and (world_size 2, gloo) the row sharding of the images
# with its single all-reduce of the gradients -- against the same golden vectors of the unmodified reference.
def _cpu_epoch(g, rows, lazy, group=None, n_total=None):
    import gppvae_b200
    from gppvae_b200.epoch import train_epoch
    vae = _vae(g, torch.float32)
    vm = gppvae_b200.Vmodel(int(g["P"]), int(g["Q"]), int(g["p"]), int(g["Q"]))
    gp = gppvae_b200.GP()
    if group is not None:
        gp.shard_rows(group)
    with torch.no_grad():
        vm.x0.copy_(torch.as_tensor(g["init.vm.x0"])); vm.v0.copy_(torch.as_tensor(g["init.vm.v0"]))
        gp.lvs.copy_(torch.as_tensor(g["init.gp.lvs"]))
    Y, Eps, D, W = (torch.as_tensor(g[k])[rows] for k in ("Y", "Eps", "D", "W"))
    bs = int(g["bs"])
    # the reference's minibatch order, restricted to this rank's rows and renumbered locally
    order = torch.as_tensor(g["order"])
    local = {int(r): i for i, r in enumerate(rows.tolist())}
    mine = torch.tensor([local[int(r)] for r in order.tolist() if int(r) in local])
    batches = [mine[a:a + bs] for a in range(0, mine.numel(), bs)]
    vae_opt = torch.optim.Adam(vae.parameters(), lr=2e-4)
    gp_opt = torch.optim.Adam(list(vm.parameters()) + list(gp.parameters()), lr=1e-3)
    rv = train_epoch(vae, vm, gp, Y, D, W, vae_opt, gp_opt, bs=bs, eps=Eps, batches=batches, lazy=lazy, step=False,
                     group=group, n_total=n_total)
    grads = {"vae." + n: p.grad.clone() for n, p in vae.named_parameters() if p.grad is not None}
    grads.update({"vm.x0": vm.x0.grad.clone(), "vm.v0": vm.v0.grad.clone(), "gp.lvs": gp.lvs.grad.clone()})
    return rv, grads


def _check_epoch(g, rv, grads):
    for key in ("mse", "recon_term", "pen_term", "gp_nll", "loss"):
        assert abs(rv[key] - float(g["out." + key])) < 2e-4 * abs(float(g["out." + key])), key
    for name, gr in grads.items():
        assert rel_err(gr, g["grad." + name]) < 2e-3, name


@pytest.mark.parametrize("lazy", [True, False])
def test_epoch_host_logic_cpu(monkeypatch, lazy):
    import fake_engine
    fake_engine.install(monkeypatch)
    g = _golden()
    rv, grads = _cpu_epoch(g, torch.arange(g["Y"].shape[0]), lazy)
    _check_epoch(g, rv, grads)


def _epoch_shard_worker(rank, world, port, out):
    import torch.distributed as dist
    from _pytest.monkeypatch import MonkeyPatch
    import fake_engine
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mpch = MonkeyPatch()
    fake_engine.install(mpch)
    try:
        torch.set_num_threads(2)
        g = _golden()
        n = g["Y"].shape[0]
        cut = int(os.environ.get("GPP_TEST_CUT", (n * 3) // 5))    # default: deliberately unequal shards
        rows = torch.arange(0, cut) if rank == 0 else torch.arange(cut, n)
        rv, grads = _cpu_epoch(g, rows, True, group=dist.group.WORLD, n_total=n)
        torch.save(dict(rv=rv, grads=grads), os.path.join(out, f"r{rank}.pt"))
    finally:
        mpch.undo()
        dist.destroy_process_group()


def test_epoch_row_sharding_world2_gloo(tmp_path):
    """Two ranks hold 24 and 16 of the 40 images: every rank encodes / decodes its own images, the GP term all-reduces
    its Q-space partials, gradients are all-reduced once -- metrics and gradients equal the unsharded reference's on
    both ranks."""
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_epoch_shard_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g = _golden()
    r0 = torch.load(os.path.join(tmp_path, "r0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "r1.pt"))
    for r in (r0, r1):
        _check_epoch(g, r["rv"], r["grads"])
    for name in r0["grads"]:
        assert torch.equal(r0["grads"][name], r1["grads"][name]), name


def test_epoch_row_sharding_even_split_lvs_gradient(tmp_path, monkeypatch):
    """20 + 20 of the 40 images with 16-image minibatches: the ranks run 2 + 2 = 4 minibatches where the reference runs
    ceil(40 / 16) = 3, and gp.py:131-132 adds <vbs, vs> once per minibatch -- the accumulated lvs gradient has to be
    brought back to the reference's count (it would be 4/3 of it otherwise)."""
    import torch.multiprocessing as mp
    g = _golden()
    n, bs = g["Y"].shape[0], int(g["bs"])
    assert 2 * -(-(n // 2) // bs) != -(-n // bs)            # the split really changes the minibatch count
    monkeypatch.setenv("GPP_TEST_CUT", str(n // 2))
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_epoch_shard_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "r0.pt"))
    # minibatch composition differs from the reference's, so only what does not depend on it is compared: the lvs
    # gradient (J^T vbs times the minibatch count) and the epoch metrics
    assert rel_err(r0["grads"]["gp.lvs"], g["grad.gp.lvs"]) < 2e-3
    for key in ("gp_nll", "pen_term"):
        assert abs(r0["rv"][key] - float(g["out." + key])) < 2e-4 * abs(float(g["out." + key])), key
