"""`cuda_ba` -- drop-in for the reference's pybind module (devo/fastba/ba.cpp:152-155):
forward (in-place Gauss-Newton BA), neighbors, reproject.

`forward` mutates `poses` and `patches` in place and returns [] like the reference.
Error behaviour: the reference raises from torch::linalg::cholesky when the Schur
system is not positive definite (the caller, devo/devo.py:336-340, catches it).  With
STRICT (default) this module reads the device status word after the launch sequence
and raises RuntimeError likewise; the fused engine uses the non-blocking form.
"""
import torch

from . import _lib

STRICT = True


def _f32c(t, name, inplace=False):
    _lib.require_cuda(t)
    _lib.require_dtype(t, torch.float32, name)
    if not t.is_contiguous():
        if inplace:
            raise RuntimeError("cuda_ba: %s is updated in place and must be contiguous" % name)
        t = t.contiguous()
    return t


def _idx(t, name):
    _lib.require_cuda(t)
    _lib.require_dtype(t, torch.int64, name)
    return t.contiguous()


def prepare(E, n_free, status, workspace):
    """housekeeping of the next `forward_async(..., prepared=True)` on this workspace / status word (status and ticket
    area cleared by one small kernel -- no copy-engine node inside a captured step), on the
    CURRENT stream -- call it wherever it is off the critical path and make that stream precede the BA"""
    _lib.check(_lib.lib().devo_ba_prepare(workspace.data_ptr(), workspace.numel(), int(E), int(n_free), status.data_ptr(),
                                          _lib.stream_ptr(workspace.device)), "ba_prepare")


def forward_async(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, iterations, status=None,
                  plan=None, workspace=None, prepared=False, status_or=None):
    """enqueue the BA; returns the device status tensor (int32[1]) without synchronising.
    plan: a GraphPlan(kk, jj) of the same edge list to reuse (skips the internal sort).
    prepared: `prepare()` ran for this workspace / status since the last call and the plan is older than the kernel that
    precedes this call in the stream: the call launches nothing but its iterations; status_or: int32[1] the status is
    OR-ed into by the last launch."""
    poses = _f32c(poses, "poses", True)
    patches = _f32c(patches, "patches", True)
    intrinsics = _f32c(intrinsics, "intrinsics")
    target = _f32c(target, "target")
    weight = _f32c(weight, "weight")
    lmbda = _f32c(lmbda.reshape(-1), "lmbda")
    ii, jj, kk = _idx(ii, "ii"), _idx(jj, "jj"), _idx(kk, "kk")
    E = ii.numel()
    if target.numel() != 2 * E or weight.numel() != 2 * E or jj.numel() != E or kk.numel() != E:
        raise RuntimeError("cuda_ba.forward: target/weight/ii/jj/kk sizes disagree")
    P = patches.shape[-1]
    n_poses = poses.numel() // 7
    n_patches = patches.numel() // (3 * P * P)
    dev = poses.device
    if status is None:
        status = torch.empty(1, dtype=torch.int32, device=dev)
    L = _lib.lib()
    nbytes = L.devo_ba_workspace(E, max(int(t1) - int(t0), 0))
    ws = workspace if workspace is not None else _lib.workspace(nbytes, dev, "ba")
    if ws.numel() < nbytes:
        raise RuntimeError("cuda_ba.forward: workspace too small")
    if prepared:
        if plan is None or workspace is None:
            raise RuntimeError("cuda_ba.forward_async(prepared=True) needs the plan and the workspace prepare() was called with")
        _lib.check(L.devo_ba_forward_prepared(poses.data_ptr(), patches.data_ptr(), intrinsics.data_ptr(), target.data_ptr(),
                                              weight.data_ptr(), lmbda.data_ptr(), ii.data_ptr(), jj.data_ptr(), kk.data_ptr(),
                                              E, n_poses, n_patches, P, int(t0), int(t1), int(iterations),
                                              plan.perm.data_ptr(), plan.gstart.data_ptr(), plan.gkey.data_ptr(),
                                              plan.ngroups.data_ptr(), ws.data_ptr(), ws.numel(), status.data_ptr(),
                                              _lib.ptr(status_or), _lib.stream_ptr(dev)), "ba_forward_prepared")
        return status
    if plan is not None:
        _lib.check(L.devo_ba_forward_planned(poses.data_ptr(), patches.data_ptr(), intrinsics.data_ptr(), target.data_ptr(),
                                             weight.data_ptr(), lmbda.data_ptr(), ii.data_ptr(), jj.data_ptr(), kk.data_ptr(),
                                             E, n_poses, n_patches, P, int(t0), int(t1), int(iterations),
                                             plan.perm.data_ptr(), plan.gstart.data_ptr(), plan.gkey.data_ptr(),
                                             plan.ngroups.data_ptr(), ws.data_ptr(), ws.numel(), status.data_ptr(),
                                             _lib.stream_ptr(dev)), "ba_forward_planned")
        return status
    _lib.check(L.devo_ba_forward(poses.data_ptr(), patches.data_ptr(), intrinsics.data_ptr(), target.data_ptr(),
                                 weight.data_ptr(), lmbda.data_ptr(), ii.data_ptr(), jj.data_ptr(), kk.data_ptr(),
                                 E, n_poses, n_patches, P, int(t0), int(t1), int(iterations), ws.data_ptr(),
                                 ws.numel(), status.data_ptr(), _lib.stream_ptr(dev)), "ba_forward")
    return status


def forward(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, iterations):
    status = forward_async(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, iterations)
    if STRICT:
        code = int(status.item())
        if code > 0:
            raise RuntimeError("cuda_ba.forward: Schur complement not positive definite at iteration %d "
                               "(linalg.cholesky would have raised)" % (code - 1))
        if code < 0:
            raise RuntimeError("cuda_ba.forward: capacity error %d (a patch has more edges than one batch holds)" % code)
    return []


class ShardedBA:
    """One rank's side of an edge-sharded Gauss-Newton BA: this rank holds only the edges it owns -- all
    edges of a patch live on one rank (devo_b200.dist.shard_edges_by_patch) -- and the full, replicated
    `poses`/`patches`.  Per iteration: `accumulate(itr)` (local edges -> `self.system`, fp64 partial of the reduced
    system [S|y] + a status word) -> ONE all-reduce(sum) of `self.system` over the ranks -> `solve(itr)` (identical on
    every rank: damping, LDL^T, pose retraction).  `finish()` applies the last depth update to the local patches.
    Nothing synchronises with the host."""

    def __init__(self, poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, status=None):
        self.poses = _f32c(poses, "poses", True)
        self.patches = _f32c(patches, "patches", True)
        self.intrinsics = _f32c(intrinsics, "intrinsics")
        self.target = _f32c(target, "target")
        self.weight = _f32c(weight, "weight")
        self.lmbda = _f32c(lmbda.reshape(-1), "lmbda")
        self.ii, self.jj, self.kk = _idx(ii, "ii"), _idx(jj, "jj"), _idx(kk, "kk")
        E = self.E = self.ii.numel()
        if self.target.numel() != 2 * E or self.weight.numel() != 2 * E or self.jj.numel() != E or self.kk.numel() != E:
            raise RuntimeError("cuda_ba.ShardedBA: target/weight/ii/jj/kk sizes disagree")
        self.t0, self.t1 = int(t0), int(t1)
        if self.t1 - self.t0 <= 0:
            raise RuntimeError("cuda_ba.ShardedBA: needs a free pose (structure-only BA has nothing to exchange)")
        self.P = self.patches.shape[-1]
        self.n_poses = self.poses.numel() // 7
        self.n_patches = self.patches.numel() // (3 * self.P * self.P)
        dev = self.poses.device
        L = _lib.lib()
        self.status = status if status is not None else torch.zeros(1, dtype=torch.int32, device=dev)
        self._ws = torch.empty(max(L.devo_ba_workspace(E, self.t1 - self.t0), 256), dtype=torch.uint8, device=dev)
        self.system = torch.zeros(L.devo_ba_system_doubles(self.t1 - self.t0), dtype=torch.float64, device=dev)

    def _acc(self, itr, flags):
        L = _lib.lib()
        _lib.check(L.devo_ba_sharded_accumulate(
            self.poses.data_ptr(), self.patches.data_ptr(), self.intrinsics.data_ptr(), self.target.data_ptr(),
            self.weight.data_ptr(), self.lmbda.data_ptr(), self.ii.data_ptr(), self.jj.data_ptr(), self.kk.data_ptr(),
            self.E, self.n_poses, self.n_patches, self.P, self.t0, self.t1, int(itr), int(flags), self.system.data_ptr(),
            self._ws.data_ptr(), self._ws.numel(), self.status.data_ptr(), _lib.stream_ptr(self.poses.device)),
            "ba_sharded_accumulate")

    def accumulate(self, itr):
        self._acc(itr, 2 | (1 if itr > 0 else 0) | (4 if itr == 0 else 0))
        return self.system

    def solve(self, itr):
        _lib.check(_lib.lib().devo_ba_sharded_solve(
            self.poses.data_ptr(), self.system.data_ptr(), self.E, self.n_poses, self.t0, self.t1, int(itr),
            self._ws.data_ptr(), self._ws.numel(), self.status.data_ptr(), _lib.stream_ptr(self.poses.device)),
            "ba_sharded_solve")

    def finish(self, iterations):
        self._acc(iterations, 1)

    # ---- all-reduce fused into the solve over NVLink peer memory (no NCCL call on the data path) --------------------
    def enable_peer(self, group=None):
        """Put the partial system into torch symmetric memory (peer-mapped on every rank of `group`): [2][nsys] doubles,
        double-buffered by iteration parity, + two 64-bit epoch flags.  Collective: every rank must call it."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        nsys = _lib.lib().devo_ba_system_doubles(self.t1 - self.t0)
        dev = self.poses.device
        self._nsys = nsys
        self._sym = symm_mem.empty(2 * nsys + 2, dtype=torch.float64, device=dev)
        self._sym.zero_()
        self._hdl = symm_mem.rendezvous(self._sym, group if group is not None else dist.group.WORLD)
        self._epoch = 0
        torch.cuda.synchronize(dev)
        dist.barrier(group)                      # every buffer is zeroed (flags = 0) before anybody polls a peer
        return self

    def accumulate_peer(self, itr):
        self._epoch += 1
        par = self._epoch & 1
        self.system = self._sym[par * self._nsys:(par + 1) * self._nsys]
        return self.accumulate(itr)

    def solve_peer(self, itr):
        """flag handshake + rank-ordered peer reduction + solve in ONE kernel (devo_ba_sharded_solve_peer)"""
        h = self._hdl
        _lib.check(_lib.lib().devo_ba_sharded_solve_peer(
            self.poses.data_ptr(), int(h.buffer_ptrs_dev), int(h.world_size), int(h.rank), int(self._epoch), self.E,
            self.n_poses, self.t0, self.t1, int(itr), self._ws.data_ptr(), self._ws.numel(), self.status.data_ptr(),
            _lib.stream_ptr(self.poses.device)), "ba_sharded_solve_peer")


def forward_sharded(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, iterations, group=None,
                    status=None, peer=False):
    """fastba.BA for ONE frame graph split over the ranks of `group` (north_star: "a single NCCL all-reduce of the
    pose-block Hessian"): see ShardedBA.  The all-reduce runs on NCCL over NVLink (fp64 sum, 7.2 KB at 7 free poses);
    poses stay bitwise replicated; only the depths of the local patches change (dist.gather_patch_depths exchanges
    them when the caller needs all of them).  Returns the device status tensor."""
    import torch.distributed as dist
    if int(t1) - int(t0) <= 0:      # structure-only BA: depths are rank-local, nothing to exchange
        return forward_async(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, iterations, status=status)
    ba = ShardedBA(poses, patches, intrinsics, target, weight, lmbda, ii, jj, kk, t0, t1, status=status)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if peer and multi:          # peer=True: no NCCL on the data path -- the reduction happens inside the solve kernel
        ba.enable_peer(group)
        for itr in range(int(iterations)):
            ba.accumulate_peer(itr)
            ba.solve_peer(itr)
        if iterations > 0:
            ba.finish(int(iterations))
        torch.cuda.synchronize(ba.poses.device)
        dist.barrier(group)     # the symmetric buffer is released below: no peer may still be reading it
        return ba.status
    for itr in range(int(iterations)):
        ba.accumulate(itr)
        if multi:
            dist.all_reduce(ba.system, op=dist.ReduceOp.SUM, group=group)
        ba.solve(itr)
    if iterations > 0:
        ba.finish(int(iterations))
    return ba.status


def neighbors(ii, jj):
    """-> [ix, jx] int64 CUDA tensors (previous / next edge of the same ii in jj order, -1 at the ends)"""
    ii, jj = _idx(ii, "ii"), _idx(jj, "jj")
    E = ii.numel()
    if jj.numel() != E:
        raise RuntimeError("cuda_ba.neighbors: ii and jj must have the same length")
    ix = torch.empty(E, dtype=torch.int64, device=ii.device)
    jx = torch.empty(E, dtype=torch.int64, device=ii.device)
    if E == 0:
        return [ix, jx]
    L = _lib.lib()
    ws = _lib.workspace(L.devo_graph_plan_workspace(E), ii.device, "plan")
    _lib.check(L.devo_neighbors(ii.data_ptr(), jj.data_ptr(), ix.data_ptr(), jx.data_ptr(), E, ws.data_ptr(),
                                ws.numel(), _lib.stream_ptr(ii.device)), "neighbors")
    return [ix, jx]


def reproject(poses, patches, intrinsics, ii, jj, kk):
    """-> coords [1,E,2,P,P] float32"""
    poses = _f32c(poses, "poses")
    patches = _f32c(patches, "patches")
    intrinsics = _f32c(intrinsics, "intrinsics")
    ii, jj, kk = _idx(ii, "ii"), _idx(jj, "jj"), _idx(kk, "kk")
    E = ii.numel()
    P = patches.shape[-1]
    coords = torch.empty(1, E, 2, P, P, dtype=torch.float32, device=poses.device)
    _lib.check(_lib.lib().devo_reproject(poses.data_ptr(), patches.data_ptr(), intrinsics.data_ptr(), ii.data_ptr(),
                                         jj.data_ptr(), kk.data_ptr(), coords.data_ptr(), E, P,
                                         _lib.stream_ptr(poses.device)), "reproject")
    return coords


class GraphPlan:
    """device-side analysis of an edge list grouped by `ka` and ordered by `kb`
    (replaces torch.unique / fastba.neighbors host round trips; see include/devo_b200.h).
    `capacity`: allocate the outputs for up to that many edges, so that `rebind` can follow an edge list that grows and
    shrinks (DEVO.append_factors / remove_factors) without reallocating; the public tensors are views of the first E."""

    def __init__(self, ka, kb, max_ka=-1, max_kb=-1, want_neighbors=True, capacity=None):
        ka, kb = _idx(ka, "ka"), _idx(kb, "kb")
        dev = ka.device
        cap = max(int(capacity or 0), ka.numel(), 1)
        self.capacity = cap
        self._perm = torch.empty(cap, dtype=torch.int32, device=dev)
        self._gid = torch.empty(cap, dtype=torch.int32, device=dev)
        self._gstart = torch.empty(cap + 1, dtype=torch.int32, device=dev)
        self._gkey = torch.empty(cap, dtype=torch.int64, device=dev)
        self.ngroups = torch.zeros(1, dtype=torch.int32, device=dev)
        self._ix = torch.empty(cap, dtype=torch.int64, device=dev) if want_neighbors else None
        self._jx = torch.empty(cap, dtype=torch.int64, device=dev) if want_neighbors else None
        self._ws = torch.empty(max(_lib.lib().devo_graph_plan_workspace(cap), 256), dtype=torch.uint8, device=dev)
        self.rebind(ka, kb, max_ka, max_kb)

    def rebind(self, ka, kb, max_ka=None, max_kb=None):
        """analyse a (new) edge list of at most `capacity` edges"""
        ka, kb = _idx(ka, "ka"), _idx(kb, "kb")
        E = ka.numel()
        if E > self.capacity or kb.numel() != E:
            raise RuntimeError("GraphPlan.rebind: %d edges exceed the capacity %d (or ka/kb lengths differ)" % (E, self.capacity))
        self.E = E
        self._ka, self._kb = ka, kb
        if max_ka is not None:
            self._max = (int(max_ka), int(-1 if max_kb is None else max_kb))
        self.perm, self.gid, self.gkey = self._perm[:E], self._gid[:E], self._gkey[:E]
        self.gstart = self._gstart[:E + 1]
        self.ix = None if self._ix is None else self._ix[:E]
        self.jx = None if self._jx is None else self._jx[:E]
        return self.update()

    def update(self):
        """recompute in place (same buffers: CUDA-graph friendly)"""
        if self.E == 0:
            self.ngroups.zero_()
            return self
        _lib.check(_lib.lib().devo_graph_plan(self._ka.data_ptr(), self._kb.data_ptr(), self.E, self._max[0],
                                              self._max[1], self._perm.data_ptr(), self._gid.data_ptr(),
                                              self._gstart.data_ptr(), self._gkey.data_ptr(), self.ngroups.data_ptr(),
                                              _lib.ptr(self._ix), _lib.ptr(self._jx), self._ws.data_ptr(),
                                              self._ws.numel(), _lib.stream_ptr(self._ka.device)), "graph_plan")
        return self


def segment_softmax_sum(g, f, plan, max_groups):
    """y[grp] = sum_r f[r] * softmax_over_group(g)[r]  for x of shape [1,E,dim] -> [1,max_groups,dim]"""
    _lib.require_cuda(g, f)
    g = g.contiguous()
    f = f.contiguous()
    dim = g.shape[-1]
    n_rows = g.numel() // dim
    y = torch.empty(1, max_groups, dim, dtype=g.dtype, device=g.device)
    _lib.check(_lib.lib().devo_segment_softmax_sum(g.data_ptr(), f.data_ptr(), plan.perm.data_ptr(),
                                                   plan.gstart.data_ptr(), plan.ngroups.data_ptr(), max_groups,
                                                   y.data_ptr(), _lib.dtype_code(g), n_rows, dim,
                                                   _lib.stream_ptr(g.device)), "segment_softmax_sum")
    return y
