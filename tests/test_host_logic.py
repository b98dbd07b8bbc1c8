"""CPU: host-side Python logic of the product (lietorch classes + broadcasting, projective_ops
composed path, differentiable ba.BA, Update module, scatter ops, shims) exercised WITHOUT a GPU by
binding the wrappers to the CPU oracle backend inside this test process only.  The product itself
never does this: devo_b200.lietorch_backends raises for CPU tensors."""
import importlib
import os
import sys

import pytest
import torch

from oracle import lie as olie
from oracle import pops as opops

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def host():
    """devo_b200 python layer with `lietorch_backends` swapped for the oracle (test-only)"""
    import devo_b200
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.startswith("devo_b200")}
    real = importlib.import_module("devo_b200.lietorch_backends")
    for k in [k for k in sys.modules if k.startswith("devo_b200.lietorch") or k in ("devo_b200.projective_ops", "devo_b200.ba")]:
        del sys.modules[k]
    sys.modules["devo_b200.lietorch_backends"] = olie
    devo_b200.lietorch_backends = olie
    lt = importlib.import_module("devo_b200.lietorch")
    pops = importlib.import_module("devo_b200.projective_ops")
    ba = importlib.import_module("devo_b200.ba")
    yield lt, pops, ba
    for k in [k for k in sys.modules if k.startswith("devo_b200")]:
        del sys.modules[k]
    sys.modules.update({k: v for k, v in saved.items() if v is not None})
    devo_b200.lietorch_backends = real


def test_real_backend_rejects_cpu_tensors():
    from devo_b200 import lietorch_backends as be, cuda_corr, cuda_ba
    with pytest.raises(RuntimeError):
        be.expm(3, torch.zeros(2, 6))
    with pytest.raises(RuntimeError):
        cuda_ba.neighbors(torch.zeros(3, dtype=torch.long), torch.zeros(3, dtype=torch.long))
    with pytest.raises(RuntimeError):
        cuda_corr.patchify_forward(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 2), 1)


def test_group_classes_and_broadcasting(host):
    lt, pops, ba = host
    dt = torch.float64
    torch.manual_seed(0)
    for G in (lt.SO3, lt.RxSO3, lt.SE3, lt.Sim3):
        a = 0.2 * torch.randn(2, 3, 4, G.manifold_dim, dtype=dt)
        X = G.exp(a)
        assert X.shape == (2, 3, 4) and X.tangent_shape == (2, 3, 4, G.manifold_dim)
        assert torch.allclose(X.log(), a, atol=1e-9)
        assert torch.allclose((X * X.inv()).log(), torch.zeros_like(a), atol=1e-9)
        Y = G.exp(torch.randn(1, 3, 1, G.manifold_dim, dtype=dt))         # broadcast over batch dims
        Z = Y * X
        assert Z.shape == (2, 3, 4)
        p = torch.randn(2, 3, 4, 3, dtype=dt)
        assert torch.allclose((X.inv() * (X * p)), p, atol=1e-9)
        M = X.matrix()
        assert M.shape == (2, 3, 4, 4, 4)
        ph = torch.cat([p, torch.ones_like(p[..., :1])], -1)
        assert torch.allclose(torch.matmul(M, ph[..., None])[..., 0], X * ph, atol=1e-9)
        I = G.Identity(5, dtype=dt)
        assert torch.allclose(I.log(), torch.zeros(5, G.manifold_dim, dtype=dt))
        assert torch.allclose(X.retr(torch.zeros_like(a)).data, X.data, atol=1e-12)
        assert lt.cat([X, X], 0).shape == (4, 3, 4) and lt.stack([X, X], 1).shape == (2, 2, 3, 4)
        assert X[0, 1:].shape == (2, 4) and len(X.unbind(0)) == 2
        X2 = G(X.data.clone())
        X2[0] = X[1]
        assert torch.equal(X2.data[0], X.data[1])
    T = lt.SE3.exp(0.1 * torch.randn(4, 6, dtype=dt))
    assert lt.Sim3(T).data.shape == (4, 8) and lt.SE3(lt.SO3(T)).data[:, :3].abs().sum() == 0
    prm = lt.LieGroupParameter(T)
    assert prm.shape == (4, 6) and torch.allclose(prm.retr().data, T.data)


def test_transform_composed_path_equals_oracle_and_golden(host):
    lt, pops, ba = host
    g = torch.load(os.path.join(GOLD, "transform_3x8.pt"))
    P = g["problem"]
    a = (P["patches0"], P["intrinsics"], P["ii"], P["jj"], P["kk"])
    G = lt.SE3(P["poses0"])
    assert torch.allclose(pops.transform(G, *a), g["coords"], atol=1e-12)
    assert torch.allclose(pops.transform(G, *a, tonly=True), g["coords_tonly"], atol=1e-12)
    c, v, (Ji, Jj, Jz) = pops.transform(G, *a, jacobian=True)
    assert torch.equal(v, g["valid"])
    assert torch.allclose(Ji, g["Ji"], atol=1e-10) and torch.allclose(Jj, g["Jj"], atol=1e-10) and torch.allclose(Jz, g["Jz"], atol=1e-10)
    assert torch.allclose(pops.flow_mag(G, *a, beta=0.5), g["flow_mag"], atol=1e-10)
    pc = pops.point_cloud(G, P["patches0"][:, :8], P["intrinsics"], torch.zeros(8, dtype=torch.long))
    assert torch.allclose(pc, g["point_cloud"], atol=1e-12)
    d = pops.transform(G, *a, depth=True)
    assert d.shape[-1] == 3 and torch.allclose(d[..., :2], g["coords"], atol=1e-12)


def test_differentiable_ba_equals_reference_golden(host):
    lt, pops, ba = host
    g = torch.load(os.path.join(GOLD, "ba_config1.pt"))
    P = g["problem"]
    Gs, X = lt.SE3(P["poses0"]), P["patches0"]
    for it, (pg, dg) in enumerate(g["traj"]):
        Gs, X = ba.BA(Gs, X, P["intrinsics"], P["targets"], P["weights"], 1e-4, P["ii"], P["jj"], P["kk"], P["bounds"],
                      ep=10.0, fixedp=1)
        assert torch.allclose(Gs.data, pg, atol=1e-9), it
        assert torch.allclose(X[:, :, 2, 1, 1], dg, atol=1e-9), it
    g2 = torch.load(os.path.join(GOLD, "ba_4x12.pt"))
    P = g2["problem"]
    a = (P["intrinsics"], P["targets"], P["weights"], 1e-4, P["ii"], P["jj"], P["kk"])
    G1, X1 = ba.BA(lt.SE3(P["poses0"]), P["patches0"], *a, [20, 20, 140, 100], ep=10.0, fixedp=1)
    assert torch.allclose(G1.data, g2["poses_a"], atol=1e-9) and torch.allclose(X1[:, :, 2, 1, 1], g2["depth_a"], atol=1e-9)
    G2, X2 = ba.BA(lt.SE3(P["poses0"]), P["patches0"], *a, P["bounds"], ep=100.0, fixedp=2, structure_only=True)
    assert torch.allclose(G2.data, g2["poses_b"], atol=1e-12) and torch.allclose(X2[:, :, 2, 1, 1], g2["depth_b"], atol=1e-9)
    # lmbda as a tensor, and gradients w.r.t. weights / targets / depth
    w = P["weights"].clone().requires_grad_(True)
    t = P["targets"].clone().requires_grad_(True)
    G3, X3 = ba.BA(lt.SE3(P["poses0"]), P["patches0"], P["intrinsics"], t, w, torch.tensor(1e-4, dtype=torch.float64),
                   P["ii"], P["jj"], P["kk"], [20, 20, 140, 100], ep=10.0, fixedp=1)
    assert torch.allclose(G3.data, g2["poses_a"], atol=1e-9)
    (G3.data[..., :3].sum() + X3[:, :, 2].sum()).backward()
    assert torch.isfinite(w.grad).all() and torch.isfinite(t.grad).all() and w.grad.abs().sum() > 0 and t.grad.abs().sum() > 0


def test_update_module_state_dict_keys_and_semantics():
    """same parameter names as devo/enet.py:32-99 so reference checkpoints load unchanged"""
    from devo_b200.update import Update
    up = Update(3)
    keys = set(up.state_dict().keys())
    for k in ["c1.0.weight", "c1.2.bias", "c2.0.weight", "norm.weight", "agg_kk.f.weight", "agg_kk.g.bias", "agg_kk.h.weight",
              "agg_ij.f.weight", "gru.0.weight", "gru.1.gate.0.weight", "gru.1.res.0.weight", "gru.1.res.2.bias",
              "gru.2.bias", "gru.3.gate.0.bias", "corr.0.weight", "corr.2.weight", "corr.3.weight", "corr.5.bias",
              "d.1.weight", "w.1.bias"]:
        assert k in keys, k
    assert up.corr[0].in_features == 882 and sum(p.numel() for p in up.parameters()) > 2_500_000


def test_scatter_and_shims():
    from devo_b200 import scatter
    import devo_b200
    x = torch.randn(1, 9, 3, requires_grad=True)
    idx = torch.tensor([2, 0, 0, 1, 2, 2, 0, 1, 1])
    s = scatter.scatter_sum(x, idx, dim=1, dim_size=4)
    assert s.shape == (1, 4, 3) and torch.allclose(s[0, 2], x[0, idx == 2].sum(0)) and s[0, 3].abs().sum() == 0
    w = scatter.scatter_softmax(x, idx, dim=1)
    assert torch.allclose(w[0, idx == 1], torch.softmax(x[0, idx == 1], 0), atol=1e-6)
    w.sum().backward()
    devo_b200.install_shims()
    import cuda_ba, cuda_corr, lietorch_backends, torch_scatter  # noqa: F401,E401
    assert cuda_corr.forward is devo_b200.cuda_corr.forward and hasattr(cuda_ba, "neighbors") and hasattr(torch_scatter, "scatter_softmax")
    for n in ("cuda_ba", "cuda_corr", "lietorch_backends", "torch_scatter"):
        sys.modules.pop(n, None)


def test_packed_update_weights_layout_matches_header_order():
    """host logic of the fused update operator: the parameter pack handed to devo_gru_update (include/devo_b200.h,
    devo_gru_weights_t) -- stacking order, zero padding of corr[0], refresh on parameter change.  CPU tensors suffice."""
    import torch
    from devo_b200.update import PackedUpdateWeights, Update
    torch.manual_seed(0)
    up = Update(3)
    pk = PackedUpdateWeights(up, torch.float16, 896)
    order = [up.corr[2], up.corr[5], up.c1[0], up.c1[2], up.c2[0], up.c2[2], up.agg_kk.g, up.agg_kk.f, up.agg_kk.h,
             up.agg_ij.g, up.agg_ij.f, up.agg_ij.h, up.gru[1].gate[0], up.gru[1].res[0], up.gru[1].res[2],
             up.gru[3].gate[0], up.gru[3].res[0], up.gru[3].res[2]]
    assert pk.W.shape == (18 * 384, 384) and pk.bias.shape == (19, 384) and pk.W0.shape == (384, 896)
    for k, layer in enumerate(order):
        assert torch.equal(pk.W[k * 384:(k + 1) * 384], layer.weight.detach().half())
        assert torch.equal(pk.bias[1 + k], layer.bias.detach().half())
    assert torch.equal(pk.bias[0], up.corr[0].bias.detach().half())
    assert torch.equal(pk.W0[:, :882], up.corr[0].weight.detach().half()) and (pk.W0[:, 882:] == 0).all()
    assert torch.equal(pk.ln_gamma[1], up.norm.weight.detach()) and torch.equal(pk.ln_beta[3], up.gru[2].bias.detach())
    assert torch.equal(pk.head_W[:2], up.d[1].weight.detach().half()) and torch.equal(pk.head_W[2:], up.w[1].weight.detach().half())
    w_ptr = pk.W.data_ptr()
    pk.refresh()
    assert pk.W.data_ptr() == w_ptr                          # nothing changed: no repacking
    with torch.no_grad():
        up.c1[0].weight.mul_(2.0)
    pk.refresh()
    assert torch.equal(pk.W[2 * 384:3 * 384], up.c1[0].weight.detach().half())


def test_patch_range_and_gru_flops_helpers():
    from devo_b200 import synthetic
    from devo_b200.dist import patch_range
    assert [patch_range(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    # S8: 6144 edges x (896 + 16*384) + (768 + 64) group rows x 384, all x 2*384
    assert synthetic.gru_flops(6144, 768, 64, 384, 896) == 2 * 6144 * 384 * (896 + 16 * 384) + 2 * (768 + 64) * 384 * 384
