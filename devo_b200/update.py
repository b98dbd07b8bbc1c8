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
import torch.nn.functional as F

from . import cuda_ba, fastba, glue
from .scatter import scatter_softmax, scatter_sum

DIM = 384


class FrozenCast:
    """Low-precision copies of a module's Linear parameters, made once and refreshed only when a
    parameter changes.  torch.autocast re-casts every fp32 weight on every forward (about 40 cast
    kernels and 18 MB of traffic per update iteration here); running the layers by hand on the cached
    copies gives bit-identical results (same rounding of the weights, same dtype flow: Linear -> half,
    LayerNorm -> float32, element-wise ops by type promotion) without that cost."""

    def __init__(self, dtype=torch.float16):
        self.dtype = dtype
        self._cache = {}

    # Entries hold a reference to the parameter objects they were made from: id() of a dead object can be reused by a
    # new one (same _version), which must not hit the old entry.
    def get(self, p):
        hit = self._cache.get(id(p))
        if hit is None or hit[2] is not p or hit[0] != p._version or hit[1].device != p.device:
            hit = (p._version, p.detach().to(self.dtype), p)
            self._cache[id(p)] = hit
        return hit[1]

    def get_cat(self, params):
        """cached low-precision concatenation of several parameters along dim 0"""
        key = tuple(id(p) for p in params)
        ver = tuple(p._version for p in params)
        hit = self._cache.get(key)
        if (hit is None or hit[0] != ver or hit[1].device != params[0].device
                or any(a is not b for a, b in zip(hit[2], params))):
            hit = (ver, torch.cat([p.detach().to(self.dtype) for p in params], 0).contiguous(), tuple(params))
            self._cache[key] = hit
        return hit[1]

    def get_padded_t(self, w, k_padded):
        """cached low-precision W^T with the input dimension zero-padded to k_padded: [k_padded, out]"""
        key = (id(w), "padT", k_padded)
        hit = self._cache.get(key)
        if hit is None or hit[2] is not w or hit[0] != w._version or hit[1].device != w.device:
            wp = torch.zeros(w.shape[0], k_padded, dtype=self.dtype, device=w.device)
            wp[:, :w.shape[1]] = w.detach().to(self.dtype)
            hit = (w._version, wp.t(), w)
            self._cache[key] = hit
        return hit[1]

    def linear(self, layer, x):
        return F.linear(x.to(self.dtype), self.get(layer.weight), self.get(layer.bias))

    @staticmethod
    def layer_norm(layer, x):
        return F.layer_norm(x.float(), layer.normalized_shape, layer.weight, layer.bias, layer.eps)

    def linear_relu(self, layer, x2d):
        """Linear + ReLU with the ReLU in the cuBLASLt epilogue (identical result: max(.,0) commutes with rounding)"""
        return torch._addmm_activation(self.get(layer.bias), x2d, self.get(layer.weight).t(), use_gelu=False)

    def run(self, module, x):
        """evaluate an nn.Sequential / layer of the update operator with autocast semantics"""
        if isinstance(module, nn.Sequential):
            for m in module:
                x = self.run(m, x)
            return x
        if isinstance(module, nn.Linear):
            return self.linear(module, x)
        if isinstance(module, nn.LayerNorm):
            return self.layer_norm(module, x)
        if isinstance(module, nn.ReLU):
            return torch.relu(x)
        if isinstance(module, nn.Sigmoid):
            return torch.sigmoid(x)
        if isinstance(module, GradientClip):
            return x
        if isinstance(module, GatedResidual):
            return x + self.run(module.gate, x) * self.run(module.res, x)
        return module(x)


class PackedUpdateWeights:
    """The update operator's parameters in the layout devo_gru_update wants (include/devo_b200.h
    devo_gru_weights_t): all 384x384 Linear weights stacked, corr[0] zero-padded along its input dimension,
    biases stacked, the four LayerNorms stacked, both heads concatenated.  Rebuilt only when a parameter changes."""

    def __init__(self, update, dtype, corr_ld):
        self.update, self.dtype, self.corr_ld = update, dtype, corr_ld
        self._ver = None
        self.refresh()

    def _layers(self):
        u = self.update
        return [u.corr[2], u.corr[5], u.c1[0], u.c1[2], u.c2[0], u.c2[2], u.agg_kk.g, u.agg_kk.f, u.agg_kk.h,
                u.agg_ij.g, u.agg_ij.f, u.agg_ij.h, u.gru[1].gate[0], u.gru[1].res[0], u.gru[1].res[2],
                u.gru[3].gate[0], u.gru[3].res[0], u.gru[3].res[2]]

    def _version(self):
        return tuple(p._version for p in self.update.parameters()) + (next(self.update.parameters()).device,)

    def refresh(self):
        v = self._version()
        if v == self._ver:
            return self
        u, dt = self.update, self.dtype
        with torch.no_grad():
            L = self._layers()
            self.W = torch.cat([l.weight.detach().to(dt) for l in L], 0).contiguous()
            w0 = u.corr[0].weight.detach()
            self.W0 = torch.zeros(w0.shape[0], self.corr_ld, dtype=dt, device=w0.device)
            self.W0[:, :w0.shape[1]] = w0.to(dt)
            self.bias = torch.stack([u.corr[0].bias.detach().to(dt)] + [l.bias.detach().to(dt) for l in L], 0).contiguous()
            lns = [u.corr[3], u.norm, u.gru[0], u.gru[2]]
            self.ln_gamma = torch.stack([l.weight.detach().float() for l in lns], 0).contiguous()
            self.ln_beta = torch.stack([l.bias.detach().float() for l in lns], 0).contiguous()
            self.ln_eps = float(u.norm.eps)
            self.head_W = torch.cat([u.d[1].weight.detach().to(dt), u.w[1].weight.detach().to(dt)], 0).contiguous()
            self.head_b = torch.cat([u.d[1].bias.detach().to(dt), u.w[1].bias.detach().to(dt)], 0).contiguous()
        from . import _lib
        self.struct = _lib.GruWeightsStruct(self.W.data_ptr(), self.W0.data_ptr(), self.bias.data_ptr(), self.ln_gamma.data_ptr(),
                                            self.ln_beta.data_ptr(), self.ln_eps, self.head_W.data_ptr(), self.head_b.data_ptr())
        self._ver = v
        return self


def tile_local_graph(plan_kk, tile=64):
    """True when every neighbour link of the plan (previous / next edge of the same patch) stays inside the edge's own
    `tile` consecutive edges -- e.g. the patch-major all-pairs graph of enet.py:300-301 when the edges per patch divide
    the tile.  Synchronises with the host (one .item()): call it when the edge list is installed, not per update."""
    E = plan_kk.ix.numel()
    t = torch.arange(E, device=plan_kk.ix.device) // tile
    ok = ((plan_kk.ix < 0) | (plan_kk.ix // tile == t)) & ((plan_kk.jx < 0) | (plan_kk.jx // tile == t))
    return bool(ok.all().item())


class GruState:
    """The recurrent hidden state of the fused update operator: float32, in the tile layout devo_gru_update reads and
    writes in place ([tiles of 64 rows][96 float4 groups][64][4], include/devo_b200.h).  `set` / `get` convert from / to the
    reference's row-major [1,E,384]; `gather` implements `net[:, ~m]` / `torch.cat([net, zeros])` (devo.py:225-239) on the
    device without leaving the layout."""

    def __init__(self, E, device, dim=DIM):
        from . import _lib
        self.E, self.dim, self.device = int(E), dim, torch.device(device)
        self.buf = torch.zeros(max(_lib.lib().devo_gru_state_floats(self.E), 4), dtype=torch.float32, device=self.device)

    def _gather(self, src, src_layout, src_rows, idx, dst, dst_layout, dst_rows):
        from . import _lib
        _lib.check(_lib.lib().devo_gru_state_gather(src.data_ptr(), src_layout, int(src_rows), _lib.ptr(idx), dst.data_ptr(),
                                                    dst_layout, int(dst_rows), _lib.stream_ptr(self.device)), "gru_state_gather")

    def set(self, net):
        """net: [1,E,384] (any float dtype; stored as float32)"""
        net = net.detach().reshape(self.E, self.dim).to(device=self.device, dtype=torch.float32).contiguous()
        self._gather(net, 0, self.E, None, self.buf, 1, self.E)
        return self

    def get(self):
        out = torch.empty(1, self.E, self.dim, dtype=torch.float32, device=self.device)
        self._gather(self.buf, 1, self.E, None, out, 0, self.E)
        return out

    def zero_(self):
        self.buf.zero_()
        return self

    def gather(self, idx):
        """new state whose row e is row idx[e] of this one (idx[e] < 0 => zeros)"""
        idx = idx.to(device=self.device, dtype=torch.int64).contiguous()
        out = GruState(idx.numel(), self.device, self.dim)
        self._gather(self.buf, 1, self.E, idx, out.buf, 1, out.E)
        return out


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

    def forward_planned(self, x, plan, max_groups, fc=None):
        lin = (lambda layer, t: layer(t)) if fc is None else fc.linear
        y = cuda_ba.segment_softmax_sum(lin(self.g, x), lin(self.f, x), plan, max_groups)
        return lin(self.h, y)[:, plan.gid] if self.expand else lin(self.h, y)


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

    def forward_planned(self, net, inp, corr, plan_kk, plan_ij, max_patches, max_pairs, fc=None):
        """same computation; plan_kk = GraphPlan(kk, jj), plan_ij = GraphPlan(ii*12345+jj, ...).
        With `fc` (a FrozenCast) the layers run on cached low-precision weights with autocast's exact
        dtype flow and no autocast context is needed; without it, call under torch.autocast."""
        run = (lambda m, t: m(t)) if fc is None else fc.run
        net = run(self.norm, net + inp + run(self.corr, corr))
        ix, jx = plan_kk.ix, plan_kk.jx
        net = net + run(self.c1, (ix >= 0).to(net.dtype).reshape(1, -1, 1) * net[:, ix])
        net = net + run(self.c2, (jx >= 0).to(net.dtype).reshape(1, -1, 1) * net[:, jx])
        net = net + self.agg_kk.forward_planned(net, plan_kk, max_patches, fc)
        net = net + self.agg_ij.forward_planned(net, plan_ij, max_pairs, fc)
        net = run(self.gru, net)
        return net, (run(self.d, net), run(self.w, net), None)

    def forward_mma(self, net, imap16, kk, corr16, plan_kk, plan_ij, max_patches, max_pairs, packed, net_out=None,
                    workspace=None, coords=None, state=None, tile_local=False):
        """The whole forward as fused tcgen05 kernels (csrc/gru_mma.cu, devo_gru_update).
        tile_local: the caller's promise that every neighbour link (plan_kk.ix / jx) stays inside the edge's own 64-edge
             tile (see `tile_local_graph`): corr MLP + norm, c1 and c2 + the kk-aggregation layers then run as ONE launch.
        net: the hidden state [1,E,384] -- float32 (every update but the first: enet.py GatedResidual returns float32 under
             autocast) or the autocast dtype (the half zero-state of the first update; follows the half dtype flow) --
             or None when `state` (a GruState, the engine's persistent tile-layout float32 buffer) already holds it.
        imap16 [1,Np,384] context features (inp = imap16[:, kk], gathered inside), corr16 [E, corr_ld] zero-padded
        correlation rows.  Returns (new state float32 [1,E,384] (or `state` itself when given), (delta, weight, None)).
        Inference only; autocast rounding points.  With `coords` ([1,E,2,3,3] f32, the reprojection) the BA inputs are
        produced by the same launch and returned as the third element: (target f32 [1,E,2], weight f32 [1,E,2])."""
        import ctypes
        from . import _lib
        E, D = corr16.shape[0], self.dim
        dt = packed.dtype
        for t in (imap16, corr16):
            _lib.require_cuda(t)
            _lib.require_dtype(t, dt, "forward_mma input")
            _lib.require_contiguous(t=t)
        if corr16.shape[-1] != packed.corr_ld:
            raise RuntimeError("forward_mma: correlation rows must be padded to %d" % packed.corr_ld)
        dev = corr16.device
        own_state = state is None
        if own_state:
            state = GruState(E, dev)
        net16_ptr = 0
        if net is not None:
            _lib.require_cuda(net)
            if net.numel() != E * D:
                raise RuntimeError("forward_mma: the hidden state must be [1,E,%d]" % D)
            if net.dtype == torch.float32:
                state.set(net)
            else:
                _lib.require_dtype(net, dt, "forward_mma hidden state")
                net = net.contiguous()
                net16_ptr = net.data_ptr()
        delta = torch.empty(1, E, 2, dtype=dt, device=dev)
        weight = torch.empty(1, E, 2, dtype=dt, device=dev)
        L = _lib.lib()
        nbytes = L.devo_gru_workspace(E, max(max_patches, max_pairs))
        ws = workspace if workspace is not None else _lib.workspace(nbytes, dev, "gru")
        if ws.numel() < nbytes:
            raise RuntimeError("forward_mma: workspace too small")
        packed.refresh()
        io = _lib.GruIoStruct(E, D, packed.corr_ld, corr16.data_ptr(), state.buf.data_ptr(), net16_ptr, imap16.data_ptr(), kk.data_ptr(),
                              plan_kk.ix.data_ptr(), plan_kk.jx.data_ptr(),
                              plan_kk.perm.data_ptr(), plan_kk.gstart.data_ptr(), plan_kk.ngroups.data_ptr(), plan_kk.gid.data_ptr(), int(max_patches),
                              plan_ij.perm.data_ptr(), plan_ij.gstart.data_ptr(), plan_ij.ngroups.data_ptr(), plan_ij.gid.data_ptr(), int(max_pairs),
                              0 if net_out is None else net_out.data_ptr(), delta.data_ptr(), weight.data_ptr(), 0, 0, 0, int(bool(tile_local)))
        if net_out is not None:
            _lib.require_dtype(net_out, dt, "forward_mma net_out")
            _lib.require_contiguous(net_out=net_out)
        extra = None
        if coords is not None:
            _lib.require_dtype(coords, torch.float32, "coords")
            _lib.require_contiguous(coords=coords)
            if coords.numel() != E * 18:
                raise RuntimeError("forward_mma: coords must be [1,E,2,3,3]")
            extra = (torch.empty(1, E, 2, dtype=torch.float32, device=dev), torch.empty(1, E, 2, dtype=torch.float32, device=dev))
            io.coords, io.target32, io.weight32 = coords.data_ptr(), extra[0].data_ptr(), extra[1].data_ptr()
        _lib.check(L.devo_gru_update(ctypes.byref(packed.struct), ctypes.byref(io), _lib.dtype_code(corr16), ws.data_ptr(),
                                     ws.numel(), _lib.stream_ptr(dev)), "gru_update")
        return (state.get() if own_state else state), (delta, weight, extra)

    def forward_fused(self, net16, inp16, corr16, plan_kk, plan_ij, max_patches, max_pairs, fc, net_out=None):
        """forward_planned with the element-wise glue fused into hand-written kernels (devo_b200.glue) and
        ReLUs folded into GEMM epilogues.  Inputs are [1,E,*] tensors of the autocast dtype; returns
        (net float32 [1,E,dim], (delta, weight, None)); `net_out` optionally receives the new hidden state
        rounded to the autocast dtype.  Same rounding points as autocast; inference only."""
        E, D, hd = net16.shape[1], self.dim, fc.dtype
        lin, lin_relu = fc.linear, fc.linear_relu
        c2d = corr16.reshape(E, -1)
        if c2d.shape[1] == self.corr[0].in_features:
            c = lin_relu(self.corr[0], c2d)
        else:   # rows zero-padded to a GEMM-friendly K (e.g. 882 -> 896): multiply by the zero-padded weight
            c = torch._addmm_activation(fc.get(self.corr[0].bias), c2d, fc.get_padded_t(self.corr[0].weight, c2d.shape[1]), use_gelu=False)
        c = lin(self.corr[2], c)
        c = glue.layernorm_relu_half(c, self.corr[3])
        c = lin(self.corr[5], c)
        net = glue.layernorm_add3(net16.reshape(E, D), inp16.reshape(E, D), c, self.norm)       # float32 [E,D]
        for seq, idx in ((self.c1, plan_kk.ix), (self.c2, plan_kk.jx)):
            g = glue.gather_mask_cast(net, idx, hd)
            y = lin(seq[2], lin_relu(seq[0], g))
            x16 = glue.residual_add_(net, y, want_half=seq is self.c2)
        for agg, plan, mg in ((self.agg_kk, plan_kk, max_patches), (self.agg_ij, plan_ij, max_pairs)):
            y = cuda_ba.segment_softmax_sum(lin(agg.g, x16), lin(agg.f, x16), plan, mg)
            hy = lin(agg.h, y.reshape(mg, D))
            x16 = glue.residual_add_(net, hy, gid=plan.gid, want_half=agg is self.agg_kk)
        for ln, gr in ((self.gru[0], self.gru[1]), (self.gru[2], self.gru[3])):
            n32, n16 = glue.layernorm_f32(net, ln, hd)
            gate = lin(gr.gate[0], n16)
            res = lin(gr.res[2], lin_relu(gr.res[0], n16))
            net = glue.gated_residual(n32, gate, res)
        delta, weight = glue.heads(net, fc.get_cat((self.d[1].weight, self.w[1].weight)), fc.get_cat((self.d[1].bias, self.w[1].bias)))
        if net_out is not None:
            glue.relu_cast(net, hd, relu=False, out=net_out.reshape(E, D))
        return net.view(1, E, D), (delta.view(1, E, 2), weight.view(1, E, 2), None)
