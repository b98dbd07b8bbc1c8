"""The recurrent update operator ("ConvGRU") of DEVO -- same architecture and state_dict
keys as devo/enet.py:32-99 (`Update`) and devo/blocks.py:15-48,70-90 (GatedResidual,
SoftAgg, GradientClip), so reference checkpoints load unchanged.  Dense layers stay
PyTorch / cuBLAS (only the index ops around them are on the hot path):

  * `forward(net, inp, corr, flow, ii, jj, kk)`       reference semantics, differentiable;
        neighbours from fastba.neighbors (GPU kernel), groups from torch.unique.
  * `forward_planned(net, inp, corr, plan_kk, plan_ij, ...)`  inference path with no host
        sync at all: neighbours and groups come from device-side GraphPlans and the two
        SoftAgg reductions run as one fused segment-softmax-sum kernel each.
"""
import torch
import torch.nn as nn

from . import cuda_ba, fastba
from .scatter import scatter_softmax, scatter_sum

DIM = 384


class _ClipGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x

    @staticmethod
    def backward(ctx, g):
        g = torch.where(torch.isnan(g), torch.zeros_like(g), g)
        return g.clamp(min=-0.01, max=0.01)


class GradientClip(nn.Module):
    def forward(self, x):
        return _ClipGrad.apply(x)


class GatedResidual(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.gate = nn.Sequential(nn.Linear(dim, dim), nn.Sigmoid())
        self.res = nn.Sequential(nn.Linear(dim, dim), nn.ReLU(inplace=True), nn.Linear(dim, dim))

    def forward(self, x):
        return x + self.gate(x) * self.res(x)


class SoftAgg(nn.Module):
    """softmax-weighted aggregation over groups of edges, broadcast back to the edges"""

    def __init__(self, dim=512, expand=True):
        super().__init__()
        self.dim = dim
        self.expand = expand
        self.f = nn.Linear(dim, dim)
        self.g = nn.Linear(dim, dim)
        self.h = nn.Linear(dim, dim)

    def forward(self, x, ix):
        _, jx = torch.unique(ix, return_inverse=True)
        w = scatter_softmax(self.g(x), jx, dim=1)
        y = scatter_sum(self.f(x) * w, jx, dim=1)
        return self.h(y)[:, jx] if self.expand else self.h(y)

    def forward_planned(self, x, plan, max_groups):
        y = cuda_ba.segment_softmax_sum(self.g(x), self.f(x), plan, max_groups)
        return self.h(y)[:, plan.gid] if self.expand else self.h(y)


class Update(nn.Module):
    def __init__(self, p, dim=DIM):
        super().__init__()
        self.dim = dim
        self.c1 = nn.Sequential(nn.Linear(dim, dim), nn.ReLU(inplace=True), nn.Linear(dim, dim))
        self.c2 = nn.Sequential(nn.Linear(dim, dim), nn.ReLU(inplace=True), nn.Linear(dim, dim))
        self.norm = nn.LayerNorm(dim, eps=1e-3)
        self.agg_kk = SoftAgg(dim)
        self.agg_ij = SoftAgg(dim)
        self.gru = nn.Sequential(nn.LayerNorm(dim, eps=1e-3), GatedResidual(dim),
                                 nn.LayerNorm(dim, eps=1e-3), GatedResidual(dim))
        self.corr = nn.Sequential(nn.Linear(2 * 49 * p * p, dim), nn.ReLU(inplace=True), nn.Linear(dim, dim),
                                  nn.LayerNorm(dim, eps=1e-3), nn.ReLU(inplace=True), nn.Linear(dim, dim))
        self.d = nn.Sequential(nn.ReLU(inplace=False), nn.Linear(dim, 2), GradientClip())
        self.w = nn.Sequential(nn.ReLU(inplace=False), nn.Linear(dim, 2), GradientClip(), nn.Sigmoid())

    def _heads(self, net):
        net = self.gru(net)
        return net, (self.d(net), self.w(net), None)

    def forward(self, net, inp, corr, flow, ii, jj, kk):
        net = self.norm(net + inp + self.corr(corr))
        ix, jx = fastba.neighbors(kk, jj)
        net = net + self.c1((ix >= 0).float().reshape(1, -1, 1) * net[:, ix])
        net = net + self.c2((jx >= 0).float().reshape(1, -1, 1) * net[:, jx])
        net = net + self.agg_kk(net, kk)
        net = net + self.agg_ij(net, ii * 12345 + jj)
        return self._heads(net)

    def forward_planned(self, net, inp, corr, plan_kk, plan_ij, max_patches, max_pairs):
        """same computation; plan_kk = GraphPlan(kk, jj), plan_ij = GraphPlan(ii*12345+jj, ...)"""
        net = self.norm(net + inp + self.corr(corr))
        ix, jx = plan_kk.ix, plan_kk.jx
        net = net + self.c1((ix >= 0).to(net.dtype).reshape(1, -1, 1) * net[:, ix])
        net = net + self.c2((jx >= 0).to(net.dtype).reshape(1, -1, 1) * net[:, jx])
        net = net + self.agg_kk.forward_planned(net, plan_kk, max_patches)
        net = net + self.agg_ij.forward_planned(net, plan_ij, max_pairs)
        return self._heads(net)
