"""GPU parity of the structured (never-materialised) Khatri-Rao route (csrc/structured.cu, vmod.KhatriRao;
SURVEY.md 8(f) row 4) against the dense CUDA route, the float64 oracle and the golden-pinned semantics of
train_gppvae.py:161-166, 235-237, 283."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

NLL_TOL = 1e-5      # relative error of sum(nll)            (BASELINE.json north_star)
GRAD_TOL = 1e-4     # max-relative error of dNLL/dZ          (BASELINE.json north_star)


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _models(pr, p, q, dev):
    import gppvae_b200
    vm = gppvae_b200.Vmodel(pr.x0.shape[0], pr.v0.shape[0], p, q).to(dev)
    gp = gppvae_b200.GP().to(dev)
    with torch.no_grad():
        vm.x0.copy_(pr.x0.to(dev)); vm.v0.copy_(pr.v0.to(dev)); gp.lvs.copy_(pr.lvs.to(dev))
    return vm, gp


@pytest.mark.parametrize("kind,lvs", [("trained", (0.0, 0.0)), ("init", (0.0, 0.0)), ("trained", (2.0, -4.0))])
def test_structured_c1_against_oracle_and_dense(dev, kind, lvs):
    """c1 shape (N=4005, p=64, q=9, L=256): structured taylor_coeff vs the fp64 oracle and vs the dense CUDA route."""
    from gppvae_b200.synth import make_problem
    from oracle import gp_oracle as O
    pr = make_problem(4005, 64, 9, 256, kind=kind, lvs=lvs, seed=11)
    V64 = O.feature_map(pr.x0.double(), pr.v0.double(), pr.d, pr.w)
    o64 = O.taylor_coeff(pr.Z.double(), [V64], pr.lvs.double())
    vm, gp = _models(pr, 64, 9, dev)
    d, w, Z = pr.d.to(dev), pr.w.to(dev), pr.Z.to(dev)
    kr = vm.lazy(d, w)
    assert tuple(kr.shape) == (4005, 576)
    Xb, Vbs, vbs, nll = gp.taylor_coeff(Z, [kr])
    e_nll = abs(nll.double().sum().item() - o64[3].sum().item()) / abs(o64[3].sum().item())
    idx = torch.arange(0, 4005, 31, device=dev)
    e_vb = rel_err(Vbs[0][idx].cpu(), o64[1][0][idx.cpu()])
    print(f"[structured c1 {kind} lvs={lvs}] vs fp64: NLL {e_nll:.2e}  Xb {rel_err(Xb.cpu(), o64[0]):.2e}  "
          f"vbs {rel_err(vbs.cpu(), o64[2]):.2e}  Vb rows {e_vb:.2e}")
    assert e_nll < NLL_TOL
    assert rel_err(Xb.cpu(), o64[0]) < GRAD_TOL
    assert rel_err(vbs.cpu(), o64[2]) < 1e-4
    assert e_vb < 2e-3
    gp2 = _models(pr, 64, 9, dev)[1]
    with torch.no_grad():
        Vd = vm(d, w)
    Xb_d, _, vbs_d, nll_d = gp2.taylor_coeff(Z, [Vd], need_vb=False)
    assert rel_err(Xb.cpu(), Xb_d.cpu()) < GRAD_TOL
    assert abs(nll.double().sum().item() - nll_d.double().sum().item()) / abs(nll_d.double().sum().item()) < NLL_TOL


@pytest.mark.parametrize("n,P,nv,p,q,L", [(700, 37, 5, 8, 4, 12), (3000, 50, 7, 16, 7, 260), (513, 600, 3, 33, 4, 8),
                                          (6000, 640, 6, 128, 4, 72), (5000, 1500, 3, 130, 3, 64)])
def test_structured_repeated_and_empty_slots(dev, n, P, nv, p, q, L):
    """Random (object, view) indices: slots that hold several rows and slots that hold none, widths that need padding
    (p, L not multiples of 4), more views than view features -- against the fp64 oracle.  The last two shapes (P >= 512,
    p >= 128) take the operand-planes form of the P-long products (slot sums written as planes, scale from the slot-count
    bound)."""
    import gppvae_b200
    from oracle import gp_oracle as O
    g = torch.Generator().manual_seed(n)
    x0 = torch.randn(P, p, generator=g); v0 = torch.eye(nv, q) + 0.5 * torch.randn(nv, q, generator=g)
    d = torch.randint(0, P, (n,), generator=g); w = torch.randint(0, nv, (n,), generator=g)
    lvs = torch.tensor([0.3, -0.5])
    V64 = O.feature_map(x0.double(), v0.double(), d, w)
    Z = (0.5 * torch.randn(n, L, generator=g).double() + V64 @ torch.randn(p * q, L, generator=g).double()).float()
    o64 = O.taylor_coeff(Z.double(), [V64], lvs.double())
    vm = gppvae_b200.Vmodel(P, nv, p, q).to(dev)
    gp = gppvae_b200.GP().to(dev)
    with torch.no_grad():
        vm.x0.copy_(x0.to(dev)); vm.v0.copy_(v0.to(dev)); gp.lvs.copy_(lvs.to(dev))
    kr = vm.lazy(d.to(dev), w.to(dev))
    assert rel_err(kr.dense().cpu(), V64) < 1e-6
    Xb, Vbs, vbs, nll = gp.taylor_coeff(Z.to(dev), [kr])
    assert Xb.shape == (n, L) and nll.shape == (n, 1)
    assert abs(nll.double().sum().item() - o64[3].sum().item()) / abs(o64[3].sum().item()) < NLL_TOL
    assert rel_err(Xb.cpu(), o64[0]) < GRAD_TOL
    assert rel_err(vbs.cpu(), o64[2]) < 1e-4
    rows = torch.arange(0, n, 7, device=dev)
    assert rel_err(Vbs[0][rows].cpu(), o64[1][0][rows.cpu()]) < 2e-3
    assert rel_err(Vbs[0].dense().cpu(), o64[1][0]) < 2e-3


def test_structured_eval_step_calls(dev):
    """train_gppvae.py:235-237 with V in factored form: U_UBi_Shb -> solve (second use of the cached factorisation) and
    Vt.t().mm(Kiz), against the dense route."""
    from gppvae_b200.synth import make_problem
    pr = make_problem(2500, 32, 8, 64, kind="trained", lvs=(0.5, -1.0), seed=4)
    vm, gp = _models(pr, 32, 8, dev)
    d, w, Z = pr.d.to(dev), pr.w.to(dev), pr.Z.to(dev)
    kr = vm.lazy(d, w)
    with torch.no_grad():
        Vd = vm(d, w)
        vs = gp.get_vs()
        U, UBi, _ = gp.U_UBi_Shb([kr], vs)
        Kiz = gp.solve(Z, U, UBi, vs)
        hits = gp.cache_hits
        Xb, _, _, _ = gp.taylor_coeff(Z, [kr], need_vb=False)
        assert gp.cache_hits == hits + 1                   # :166 reuses the factorisation of :235
        gpd = _models(pr, 32, 8, dev)[1]
        Ud, UBid, _ = gpd.U_UBi_Shb([Vd], vs)
        Kiz_d = gpd.solve(Z, Ud, UBid, vs)
    assert rel_err(Kiz.cpu(), Kiz_d.cpu()) < GRAD_TOL
    assert rel_err(Xb.cpu(), Kiz_d.cpu()) < GRAD_TOL
    ref = Vd.double().t() @ Kiz.double()
    assert rel_err(kr.t().mm(Kiz).cpu(), ref.cpu()) < 1e-5
    assert rel_err(U.dense().cpu(), Ud.dense().cpu()) < 1e-6


def test_structured_nll_under_autograd(dev):
    """gp.nll(Z, [vm.lazy(d, w)]).sum().backward(): the factored V is a detached snapshot (no gradient flows to it); the
    gradients to X and lvs must equal the dense route's (this used to raise in save_for_backward)."""
    from gppvae_b200.synth import make_problem
    pr = make_problem(2500, 32, 8, 64, kind="trained", lvs=(0.5, -1.0), seed=6)
    vm, gp = _models(pr, 32, 8, dev)
    d, w, Z = pr.d.to(dev), pr.w.to(dev), pr.Z.to(dev)
    x1 = Z.clone().requires_grad_(True)
    gp.lvs.grad = None
    out = gp.nll(x1, [vm.lazy(d, w)])
    out.sum().backward()
    g_lvs = gp.lvs.grad.clone()
    with torch.no_grad():
        Vd = vm(d, w)
    x2 = Z.clone().requires_grad_(True)
    gp.lvs.grad = None
    gp.nll(x2, [Vd]).sum().backward()
    assert rel_err(x1.grad.cpu(), x2.grad.cpu()) < GRAD_TOL
    assert rel_err(g_lvs.cpu(), gp.lvs.grad.cpu()) < 1e-3
    with torch.no_grad():                       # no gradient requested anywhere: the forward alone
        assert rel_err(gp.nll(Z, [vm.lazy(d, w)]).cpu(), out.detach().cpu()) < 1e-5


def test_structured_bad_index_gives_nan_row(dev):
    import gppvae_b200
    vm = gppvae_b200.Vmodel(10, 4, 8, 4).to(dev)
    gp = gppvae_b200.GP().to(dev)
    d = torch.tensor([0, 3, 99, 5] * 8, device=dev); w = torch.tensor([0, 1, 2, 3] * 8, device=dev)
    Z = torch.randn(32, 8, device=dev)
    Xb, _, _, nll = gp.taylor_coeff(Z, [vm.lazy(d, w)], need_vb=False)
    bad = torch.arange(2, 32, 4, device=dev)
    good = torch.tensor([i for i in range(32) if i % 4 != 2], device=dev)
    assert torch.isnan(Xb[bad]).all() and torch.isfinite(Xb[good]).all()


def test_structured_full_size_c2(dev):
    """c2 at full size (N=100k, Q=1024, L=256): structured vs dense CUDA route on the same inputs."""
    from gppvae_b200.synth import CONFIGS, make_problem
    cfg = CONFIGS["c2"]
    pr = make_problem(cfg["N"], cfg["p"], cfg["q"], cfg["L"], kind="trained", lvs=(0.0, 0.0), seed=5, device=dev)
    vm, gp = _models(pr, cfg["p"], cfg["q"], dev)
    with torch.no_grad():
        Xb_s, _, vbs_s, nll_s = gp.taylor_coeff(pr.Z, [vm.lazy(pr.d, pr.w)], need_vb=False)
        gpd = _models(pr, cfg["p"], cfg["q"], dev)[1]
        Xb_d, _, vbs_d, nll_d = gpd.taylor_coeff(pr.Z, [vm(pr.d, pr.w)], need_vb=False)
    e_nll = abs(nll_s.double().sum().item() - nll_d.double().sum().item()) / abs(nll_d.double().sum().item())
    print(f"[structured c2] vs dense route: NLL {e_nll:.2e}  Xb {rel_err(Xb_s.cpu(), Xb_d.cpu()):.2e}  "
          f"vbs {rel_err(vbs_s.cpu(), vbs_d.cpu()):.2e}")
    assert e_nll < NLL_TOL
    assert rel_err(Xb_s.cpu(), Xb_d.cpu()) < GRAD_TOL
    assert rel_err(vbs_s.cpu(), vbs_d.cpu()) < 1e-4
