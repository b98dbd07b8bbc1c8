"""CPU: pins oracle/lie.py with the reference's own lietorch test properties
(devo/lietorch/run_tests.py:16-226, fp64; Sim3 tolerance relaxed as there, :263-266)."""
import os

import pytest
import torch

from oracle import lie as olie
from lie_harness import make_group, numeric_jacobian, analytic_jacobian

GIDS = [1, 2, 3, 4]
dt = torch.float64


@pytest.mark.parametrize("gid", GIDS)
def test_exp_log(gid):
    G = make_group(olie, gid)
    torch.manual_seed(gid)
    a = 0.2 * torch.randn(500, G.manifold_dim, dtype=dt)
    assert torch.allclose(a, G.exp(a).log(), atol=1e-8)


@pytest.mark.parametrize("gid", GIDS)
def test_inv(gid):
    G = make_group(olie, gid)
    torch.manual_seed(10 + gid)
    X = G.exp(0.1 * torch.randn(200, G.manifold_dim, dtype=dt))
    a = (X * X.inv()).log()
    assert torch.allclose(a, torch.zeros_like(a), atol=1e-8)


@pytest.mark.parametrize("gid", GIDS)
def test_adj_commutes(gid):
    """X * Exp(a) == Exp(Adj_X a) * X"""
    G = make_group(olie, gid)
    torch.manual_seed(20 + gid)
    X = G.exp(torch.randn(100, G.manifold_dim, dtype=dt))
    a = torch.randn(100, G.manifold_dim, dtype=dt)
    c = ((X * G.exp(a)) * (G.exp(X.adj(a)) * X).inv()).log()
    assert torch.allclose(c, torch.zeros_like(c), atol=1e-8)


@pytest.mark.parametrize("gid", GIDS)
def test_act_matches_matrix(gid):
    G = make_group(olie, gid)
    torch.manual_seed(30 + gid)
    X = G.exp(torch.randn(50, G.manifold_dim, dtype=dt))
    p = torch.randn(50, 3, dtype=dt)
    ph = torch.cat([p, torch.ones(50, 1, dtype=dt)], -1)
    p2 = torch.matmul(X.matrix(), ph[..., None])[..., 0]
    assert torch.allclose(X.act(p), p2[:, :3], atol=1e-8)
    assert torch.allclose(X.act(ph), p2, atol=1e-8)
    assert torch.allclose(olie.as_matrix(gid, X.data), X.matrix(), atol=1e-12)


@pytest.mark.parametrize("gid", GIDS)
def test_exp_log_grad_is_identity(gid):
    G = make_group(olie, gid)
    torch.manual_seed(40 + gid)
    tol = 1e-3 if gid == 4 else 1e-8
    for a in (torch.zeros(1, G.manifold_dim, dtype=dt), 0.2 * torch.randn(1, G.manifold_dim, dtype=dt)):
        J = analytic_jacobian(lambda x: G.exp(x).log(), a)
        assert torch.allclose(J, torch.eye(G.manifold_dim, dtype=dt), atol=tol)


def _gradcheck(fn, xs, atol):
    for k in range(len(xs)):
        def f(x, k=k):
            args = list(xs)
            args[k] = x
            return fn(*args)
        Ja = analytic_jacobian(f, xs[k])
        Jn = numeric_jacobian(f, xs[k], eps=1e-6)
        assert torch.allclose(Ja, Jn, atol=atol), (k, (Ja - Jn).abs().max())


@pytest.mark.parametrize("gid", GIDS)
def test_inv_log_grad(gid):
    G = make_group(olie, gid)
    torch.manual_seed(50 + gid)
    X = G.exp(0.2 * torch.randn(1, G.manifold_dim, dtype=dt))
    _gradcheck(lambda a: (G.exp(a) * X).inv().log(), [torch.zeros(1, G.manifold_dim, dtype=dt)], 1e-3 if gid == 4 else 1e-6)


@pytest.mark.parametrize("gid", GIDS)
def test_adj_and_adjT_grad(gid):
    G = make_group(olie, gid)
    torch.manual_seed(60 + gid)
    X = G.exp(0.5 * torch.randn(1, G.manifold_dim, dtype=dt))
    a0 = torch.zeros(1, G.manifold_dim, dtype=dt)
    b0 = torch.randn(1, G.manifold_dim, dtype=dt)
    _gradcheck(lambda a, b: (G.exp(a) * X).adj(b), [a0, b0], 1e-6)
    _gradcheck(lambda a, b: (G.exp(a) * X).adjT(b), [a0, b0], 1e-6)


@pytest.mark.parametrize("gid", GIDS)
def test_act_and_matrix_grad(gid):
    G = make_group(olie, gid)
    torch.manual_seed(70 + gid)
    X = G.exp(torch.randn(1, G.manifold_dim, dtype=dt))
    a0 = torch.zeros(1, G.manifold_dim, dtype=dt)
    _gradcheck(lambda a, p: (X * G.exp(a)).act(p), [a0, torch.randn(1, 3, dtype=dt)], 1e-6)
    _gradcheck(lambda a, p: (X * G.exp(a)).act(p), [a0, torch.randn(1, 4, dtype=dt)], 1e-6)
    _gradcheck(lambda a: (G.exp(a) * X).matrix(), [a0], 1e-6)


@pytest.mark.parametrize("gid", GIDS)
def test_projector_is_vec_jacobian(gid):
    """ToVec backward uses J = projector(X): d(vec(Exp(a) X))/da at a=0 must equal J[:, :K] restricted"""
    G = make_group(olie, gid)
    torch.manual_seed(80 + gid)
    X = G.exp(torch.randn(1, G.manifold_dim, dtype=dt))
    Jn = numeric_jacobian(lambda a: (G.exp(a) * X).data, torch.zeros(1, G.manifold_dim, dtype=dt), eps=1e-6)
    Pj = olie.projector(gid, X.data)[0]
    assert torch.allclose(Jn, Pj[:, :G.manifold_dim], atol=1e-6)


@pytest.mark.parametrize("gid", GIDS)
def test_jinv(gid):
    G = make_group(olie, gid)
    torch.manual_seed(90 + gid)
    X = G.exp(0.3 * torch.randn(20, G.manifold_dim, dtype=dt))
    a = torch.randn(20, G.manifold_dim, dtype=dt)
    grp = olie.GROUPS[gid]
    ref = torch.matmul(grp.left_jacobian_inverse(X.log()), a[..., None])[..., 0]
    assert torch.allclose(olie.Jinv(gid, X.data, a), ref, atol=1e-12)
    if gid != 4:   # Sim3's series jacobians are truncated (and one term is dropped in the reference)
        JJ = torch.matmul(grp.left_jacobian(X.log()), grp.left_jacobian_inverse(X.log()))
        assert torch.allclose(JJ, torch.eye(G.manifold_dim, dtype=dt).expand_as(JJ), atol=1e-8)


def test_small_angle_branches():
    """EPS=1e-6 switches (common.h:7): values just below/above the switch agree"""
    for gid in GIDS:
        G = make_group(olie, gid)
        d = torch.zeros(2, G.manifold_dim, dtype=dt)
        rot0 = 0 if gid in (1, 2) else 3
        d[0, rot0] = 0.9e-6
        d[1, rot0] = 1.1e-6
        X = G.exp(d)
        assert torch.isfinite(X.data).all()
        assert torch.allclose(X.log(), d, atol=1e-12)


def test_golden_autograd_matches_reference_python():
    """fixtures produced by the reference's devo/lietorch python on top of this oracle
    (tests/golden/make_golden.py) -- guards the harness <-> reference wrapper equivalence"""
    path = os.path.join(os.path.dirname(__file__), "golden", "lie_autograd.pt")
    gold = torch.load(path)
    names = {1: "SO3", 2: "RxSO3", 3: "SE3", 4: "Sim3"}
    for gid in GIDS:
        G = make_group(olie, gid)
        r = gold[names[gid]]
        a = r["a"].clone().requires_grad_(True)
        b = r["b"].clone().requires_grad_(True)
        p = r["p"].clone().requires_grad_(True)
        Y = G.exp(a) * G(r["X0"])
        K = G.manifold_dim
        # vec() is the identity in the forward pass; its backward multiplies by projector(X)
        vec = torch.autograd.Function  # noqa: F841
        Pj = olie.projector(gid, Y.data.detach())

        class ToVec(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x):
                return x.clone()

            @staticmethod
            def backward(ctx, g):
                return torch.matmul(g.unsqueeze(-2), Pj).squeeze(-2)

        f = (Y.inv().log() * torch.arange(1, K + 1, dtype=dt)).sum() + (Y.adjT(b) ** 2).sum() \
            + (Y.adj(b) * 0.3).sum() + (Y.act(p) ** 2).sum() + Y.matrix().sum() + (ToVec.apply(Y.data) ** 2).sum()
        ga, gb, gp = torch.autograd.grad(f, [a, b, p])
        assert torch.allclose(Y.data, r["Y"], atol=1e-12)
        assert torch.allclose(f, r["f"], atol=1e-10)
        assert torch.allclose(ga, r["ga"], atol=1e-9), (gid, (ga - r["ga"]).abs().max())
        assert torch.allclose(gb, r["gb"], atol=1e-9)
        assert torch.allclose(gp, r["gp"], atol=1e-9)
