"""Backend-parametric Lie-group harness for the tests: wraps ANY module exposing the 19
`lietorch_backends` entry points (oracle.lie on the CPU, devo_b200.lietorch_backends on the
GPU) in autograd Functions + a tiny group class, so the reference's own test properties
(devo/lietorch/run_tests.py) can be asserted against both."""
import torch


def make_group(backend, gid):
    K, N = {1: (3, 4), 2: (4, 5), 3: (6, 7), 4: (7, 8)}[gid]

    def op(fwd, bwd):
        class F(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *inputs):
                ctx.save_for_backward(*inputs)
                return fwd(gid, *[t.contiguous() for t in inputs])

            @staticmethod
            def backward(ctx, grad):
                return tuple(bwd(gid, grad.contiguous(), *[t.contiguous() for t in ctx.saved_tensors]))
        return F.apply

    Exp = op(backend.expm, backend.expm_backward)
    Log = op(backend.logm, backend.logm_backward)
    Inv = op(backend.inv, backend.inv_backward)
    Mul = op(backend.mul, backend.mul_backward)
    Adj = op(backend.adj, backend.adj_backward)
    AdjT = op(backend.adjT, backend.adjT_backward)
    Act = op(backend.act, backend.act_backward)
    Act4 = op(backend.act4, backend.act4_backward)

    class G:
        manifold_dim, embedded_dim, group_id = K, N, gid

        def __init__(self, data):
            self.data = data

        @staticmethod
        def exp(a):
            return G(Exp(a))

        def log(self):
            return Log(self.data)

        def inv(self):
            return G(Inv(self.data))

        def __mul__(self, o):
            return G(Mul(self.data, o.data))

        def adj(self, a):
            return Adj(self.data, a)

        def adjT(self, a):
            return AdjT(self.data, a)

        def act(self, p):
            return Act(self.data, p) if p.shape[-1] == 3 else Act4(self.data, p)

        def matrix(self):
            n = self.data.shape[0]
            I = torch.eye(4, dtype=self.data.dtype, device=self.data.device)
            cols = [Act4(self.data, I[c].expand(n, 4).contiguous()) for c in range(4)]
            return torch.stack(cols, dim=-1)

    return G


def numeric_jacobian(fn, x, eps=1e-6):
    """central differences of a vector function of a flat vector"""
    x = x.detach().clone()
    y0 = fn(x).reshape(-1)
    J = torch.zeros(y0.numel(), x.numel(), dtype=x.dtype, device=x.device)
    for i in range(x.numel()):
        d = torch.zeros_like(x).reshape(-1)
        d[i] = eps
        d = d.view_as(x)
        J[:, i] = (fn(x + d).reshape(-1) - fn(x - d).reshape(-1)) / (2 * eps)
    return J


def analytic_jacobian(fn, x):
    x = x.detach().clone().requires_grad_(True)
    y = fn(x).reshape(-1)
    rows = []
    for i in range(y.numel()):
        (g,) = torch.autograd.grad(y[i], x, retain_graph=True, allow_unused=True)
        rows.append(torch.zeros_like(x).reshape(-1) if g is None else g.reshape(-1))
    return torch.stack(rows, 0)
