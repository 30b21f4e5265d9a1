"""Host-side logic of the drop-in classes on CPU (fake engine from the oracle's Q-space model):
module surface, padding plumbing, factorisation cache, and the world_size-2 gloo row-sharding path."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fake_engine
from conftest import load_golden, rel_err
from oracle import gp_oracle as O


def test_module_surface_matches_reference():
    """Names, ctor signatures, parameters and state_dict keys of gp.py:11-22 / vmod.py:15-40."""
    import gppvae_b200
    gp = gppvae_b200.GP(n_rand_effs=1, vsum2one=True)
    assert list(gp.state_dict()) == ["lvs"] and gp.lvs.shape == (2,) and torch.all(gp.lvs == 0)
    for name in ("U_UBi_Shb", "solve", "get_vs", "taylor_coeff", "nll", "nll_ineff", "taylor_expansion"):
        assert callable(getattr(gp, name))
    assert torch.allclose(gp.get_vs(), torch.tensor([0.5, 0.5]))
    torch.manual_seed(0)
    vm = gppvae_b200.Vmodel(10, 9, 6, 9)
    assert list(vm.state_dict()) == ["x0", "v0"] and vm.x0.shape == (10, 6) and vm.v0.shape == (9, 9)
    assert torch.all(vm.x0[:, 0] == 1) and vm.x0[:, 1:].abs().max() < 1e-2          # vmod.py:38-39
    assert (vm.v0 - torch.eye(9)).abs().max() < 1e-2                                 # vmod.py:40
    with pytest.raises(NotImplementedError):
        gppvae_b200.GP(vsum2one=False)
    with pytest.raises(ValueError, match="CUDA"):
        vm(torch.zeros(3, dtype=torch.long), torch.zeros(3, dtype=torch.long))       # no CPU path, loudly


def test_product_does_not_import_oracle():
    import gppvae_b200
    root = os.path.dirname(gppvae_b200.__file__)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


@pytest.mark.parametrize("case", ["faceplace_init", "faceplace_trained", "genetics", "toy_vmod"])
def test_gp_host_logic_against_golden(monkeypatch, case):
    """GP.taylor_coeff / solve / nll plumbing (padding, scalar block, cache) with the fake engine."""
    import gppvae_b200
    fake_engine.install(monkeypatch)
    g = load_golden(case)
    V = torch.as_tensor(g["f64_V"], dtype=torch.float32)
    Z = torch.as_tensor(g["Z"])
    gp = gppvae_b200.GP()
    with torch.no_grad():
        gp.lvs.copy_(torch.as_tensor(g["lvs"]))
    Xb, Vbs, vbs, nll = gp.taylor_coeff(Z, [V])
    assert rel_err(nll, g["f64_nll"]) < 1e-5 and rel_err(Xb, g["f64_Xb"]) < 1e-5
    assert rel_err(Vbs[0], g["f64_Vb"]) < 1e-4 and rel_err(vbs, g["f64_vbs"]) < 1e-5
    vs = gp.get_vs()
    U, UBi, _ = gp.U_UBi_Shb([V], vs)
    assert gp.cache_hits == 1                               # same (V, lvs): factorisation reused
    assert rel_err(gp.solve(Z, U, UBi, vs), g["f64_KiX"]) < 1e-5
    with torch.no_grad():
        gp.lvs.add_(0.1)                                    # in-place update -> version bump -> cache miss
    gp.taylor_coeff(Z, [V], need_vb=False)
    assert gp.cache_hits == 1


def test_odd_widths_are_padded_and_unpadded(monkeypatch):
    import gppvae_b200
    fake_engine.install(monkeypatch)
    torch.manual_seed(1)
    V, Z = torch.randn(50, 7), torch.randn(50, 5)
    lvs = torch.tensor([0.2, -0.3])
    gp = gppvae_b200.GP()
    with torch.no_grad():
        gp.lvs.copy_(lvs)
    Xb, Vbs, vbs, nll = gp.taylor_coeff(Z, [V])
    oXb, oVbs, ovbs, onll = O.taylor_coeff(Z.double(), [V.double()], lvs.double())
    assert Xb.shape == (50, 5) and Vbs[0].shape == (50, 7)
    assert rel_err(Xb, oXb) < 1e-5 and rel_err(nll, onll) < 1e-5
    assert rel_err(Vbs[0], oVbs[0]) < 1e-4 and rel_err(vbs, ovbs) < 1e-5


def _shard_worker(rank, world, port, case, out):
    import gppvae_b200
    from _pytest.monkeypatch import MonkeyPatch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mpch = MonkeyPatch()
    fake_engine.install(mpch)
    try:
        g = load_golden(case)
        V = torch.as_tensor(g["f64_V"], dtype=torch.float32)
        Z = torch.as_tensor(g["Z"])
        n = V.shape[0]
        cut = (n * 3) // 5                                   # deliberately unequal shards
        rows = slice(0, cut) if rank == 0 else slice(cut, n)
        gp = gppvae_b200.GP().shard_rows()
        with torch.no_grad():
            gp.lvs.copy_(torch.as_tensor(g["lvs"]))
        Xb, Vbs, vbs, nll = gp.taylor_coeff(Z[rows].contiguous(), [V[rows].contiguous()])
        vs = gp.get_vs()
        U, UBi, _ = gp.U_UBi_Shb([V[rows].contiguous()], vs)
        KiX = gp.solve(Z[rows].contiguous(), U, UBi, vs)
        torch.save(dict(Xb=Xb, Vb=Vbs[0], vbs=vbs, nll=nll, KiX=KiX, hits=gp.cache_hits,
                        ntot=gp._n_total(Xb.shape[0], "cpu")), os.path.join(out, f"r{rank}.pt"))
    finally:
        mpch.undo()
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["faceplace_trained", "genetics"])
def test_row_sharding_world2_gloo(tmp_path, case):
    """two ranks each hold a row shard; GC and sum Xb^2 are all-reduced; results equal the
    unsharded reference (fp64 golden) row for row, and both ranks agree on vbs."""
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_shard_worker, args=(2, port, case, str(tmp_path)), nprocs=2, join=True)
    g = load_golden(case)
    r0 = torch.load(os.path.join(tmp_path, "r0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "r1.pt"))
    n = g["Z"].shape[0]
    assert r0["ntot"] == r1["ntot"] == n
    for key, ref in (("Xb", "f64_Xb"), ("nll", "f64_nll"), ("Vb", "f64_Vb"), ("KiX", "f64_KiX")):
        both = torch.cat([r0[key], r1[key]], 0)
        assert both.shape == g[ref].shape
        assert rel_err(both, g[ref]) < 1e-4, key
    assert torch.equal(r0["vbs"], r1["vbs"])
    assert rel_err(r0["vbs"], g["f64_vbs"]) < 1e-5
    assert r0["hits"] == r1["hits"] == 1


def _golden_kr(g):
    from gppvae_b200.vmod import KhatriRao
    return KhatriRao(torch.as_tensor(g["f64_xn"], dtype=torch.float32), torch.as_tensor(g["f64_wn"], dtype=torch.float32),
                     torch.as_tensor(g["d"]).long(), torch.as_tensor(g["w"]).long())


@pytest.mark.parametrize("case", ["faceplace_init", "faceplace_trained"])
def test_structured_route_host_logic_against_golden(monkeypatch, case):
    """The structured route (vmod.KhatriRao through GP.taylor_coeff / U_UBi_Shb / solve, LazyVb, V^T X) with the fake
    engine, against the outputs of the unmodified reference in float64: the slot index, the assembly identities of
    csrc/structured.cu and the plumbing around them."""
    import gppvae_b200
    fake_engine.install(monkeypatch)
    g = load_golden(case)
    kr = _golden_kr(g)
    Z = torch.as_tensor(g["Z"])
    order, slot_start = kr.index()
    key = kr.d * kr.nviews + kr.w
    assert torch.equal(key[order], torch.sort(key, stable=True)[0]) and slot_start[-1] == kr.n
    assert rel_err(kr.dense(), g["f64_V"]) < 1e-6
    gp = gppvae_b200.GP()
    with torch.no_grad():
        gp.lvs.copy_(torch.as_tensor(g["lvs"]))
    Xb, Vbs, vbs, nll = gp.taylor_coeff(Z, [kr])
    assert rel_err(nll, g["f64_nll"]) < 1e-5 and rel_err(Xb, g["f64_Xb"]) < 1e-5 and rel_err(vbs, g["f64_vbs"]) < 1e-5
    mb = torch.as_tensor(g["mb"]).long()
    assert rel_err(Vbs[0][mb], g["f64_Vb"][mb.numpy()]) < 1e-4            # the gather of train_gppvae.py:283
    assert rel_err(Vbs[0].dense(), g["f64_Vb"]) < 1e-4
    vs = gp.get_vs()
    U, UBi, _ = gp.U_UBi_Shb([kr], vs)
    assert gp.cache_hits == 1
    KiX = gp.solve(Z, U, UBi, vs)
    assert rel_err(KiX, g["f64_KiX"]) < 1e-5
    assert rel_err(kr.t().mm(KiX), torch.as_tensor(g["f64_V"]).t() @ KiX.double()) < 1e-5   # train_gppvae.py:237


def _kr_shard_worker(rank, world, port, case, out):
    import gppvae_b200
    from gppvae_b200.vmod import KhatriRao
    from _pytest.monkeypatch import MonkeyPatch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mpch = MonkeyPatch()
    fake_engine.install(mpch)
    try:
        g = load_golden(case)
        n = g["Z"].shape[0]
        cut = (n * 2) // 5
        rows = slice(0, cut) if rank == 0 else slice(cut, n)
        kr = KhatriRao(torch.as_tensor(g["f64_xn"], dtype=torch.float32), torch.as_tensor(g["f64_wn"], dtype=torch.float32),
                       torch.as_tensor(g["d"]).long()[rows], torch.as_tensor(g["w"]).long()[rows])
        gp = gppvae_b200.GP().shard_rows()
        with torch.no_grad():
            gp.lvs.copy_(torch.as_tensor(g["lvs"]))
        Xb, _, vbs, nll = gp.taylor_coeff(torch.as_tensor(g["Z"])[rows].contiguous(), [kr], need_vb=False)
        torch.save(dict(Xb=Xb, vbs=vbs, nll=nll), os.path.join(out, f"k{rank}.pt"))
    finally:
        mpch.undo()
        dist.destroy_process_group()


def test_structured_row_sharding_world2_gloo(tmp_path):
    """Row shards of the structured route: the ranks all-reduce ST (the slot GEMM) instead of GC."""
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_kr_shard_worker, args=(2, port, "faceplace_trained", str(tmp_path)), nprocs=2, join=True)
    g = load_golden("faceplace_trained")
    r0, r1 = torch.load(os.path.join(tmp_path, "k0.pt")), torch.load(os.path.join(tmp_path, "k1.pt"))
    assert rel_err(torch.cat([r0["Xb"], r1["Xb"]], 0), g["f64_Xb"]) < 1e-4
    assert rel_err(torch.cat([r0["nll"], r1["nll"]], 0), g["f64_nll"]) < 1e-4
    assert torch.equal(r0["vbs"], r1["vbs"]) and rel_err(r0["vbs"], g["f64_vbs"]) < 1e-5


def test_synth_generator_is_seeded_and_shardable():
    from gppvae_b200.synth import make_problem
    a = make_problem(200, 4, 5, 8, seed=3)
    b = make_problem(200, 4, 5, 8, seed=3)
    assert torch.equal(a.Z, b.Z) and torch.equal(a.d, b.d)
    s = make_problem(200, 4, 5, 8, seed=3, row_offset=120, n_rows=80)
    assert torch.equal(s.d, a.d[120:]) and torch.equal(s.w, a.w[120:]) and torch.equal(s.x0, a.x0)
    assert torch.equal(s.Z, a.Z[120:])          # a shard regenerates exactly its rows of the unsharded problem
    big = make_problem(70000, 2, 2, 4, seed=1)
    part = make_problem(70000, 2, 2, 4, seed=1, row_offset=65000, n_rows=5000)   # straddles a noise chunk boundary
    assert torch.equal(part.Z, big.Z[65000:])
    assert int(a.d.max()) < a.x0.shape[0] and int(a.w.max()) < 5
    # every (object, view) pair appears at most once: "every object seen in every view" under a permutation
    assert len({(int(x), int(y)) for x, y in zip(a.d, a.w)}) == 200
