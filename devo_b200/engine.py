"""UpdateOperator -- the B200-native form of one DEVO update iteration
(the body of DEVO.update, devo/devo.py:308-338):

    reproject -> correlation lookup (all pyramid levels) -> context gather -> Update (GRU)
    -> target/weight -> fastba.BA(iterations)

All state lives in fixed device buffers (HBM layout below), every op is enqueued on one
stream with no host synchronisation, so the whole iteration is captured once in a CUDA graph
and replayed.  Feature pyramids and patch features are held PIXEL-MAJOR (channels last) --
the layout the TMA-staged lookup kernel wants; `ingest_frame` converts a planar frame once,
when it enters the ring buffer (replaces devo/devo.py:523-527 + pyramidify, utils.py:70-79).

  poses      f32 [1, n_frames, 7]        patches  f32 [1, n_patches, 3, 3, 3]
  intrinsics f32 [1, n_frames, 4]        state    f32 tile layout (GruState; the hidden state, [1, E, 384] logically)
  imap       f16 [1, n_patches, 384]     gmap_pm  f16 [n_patches, 9, C]
  levels_pm  f16 [n_frames, H/s, W/s, C] for s in levels
  ii, jj, kk i64 [E]
"""
import torch

from . import _lib, cuda_ba, cuda_corr, projective_ops as pops
from .update import FrozenCast, GruState, PackedUpdateWeights


class UpdateOperator:
    def __init__(self, update, n_frames, patches_per_frame, n_edges, H, W, C=128, dim=384, levels=(1, 4),
                 device="cuda", feat_dtype=torch.float16, ba_iterations=2, t0=1, t1=None, fused_gru=True, gru="mma"):
        self.update = update
        self.Nf, self.M, self.E = n_frames, patches_per_frame, n_edges
        self.Np = n_frames * patches_per_frame
        self.H, self.W, self.C, self.dim = H, W, C, dim
        self.levels = tuple(levels)
        self.device = torch.device(device)
        self.feat_dtype = feat_dtype
        self.ba_iterations = ba_iterations
        self.fused_gru = fused_gru
        # "mma": fused tcgen05 kernels (csrc/gru_mma.cu); "cublas": cuBLAS Linears + glue kernels (forward_fused)
        self.gru_mode = gru if (fused_gru and dim == 384) else "cublas"
        self.t0 = t0
        self.t1 = n_frames if t1 is None else t1
        dev, f32, i64 = self.device, torch.float32, torch.int64
        # geometry + edge list live in ONE contiguous arena (256-byte aligned views), so that a host caller can refresh
        # all of it with a single H2D copy of `state_arena` (layout: `state_layout`, name -> (offset, shape, dtype))
        spec = [("poses", (1, self.Nf, 7), f32), ("patches", (1, self.Np, 3, 3, 3), f32), ("intrinsics", (1, self.Nf, 4), f32),
                ("ii", (self.E,), i64), ("jj", (self.E,), i64), ("kk", (self.E,), i64)]
        self.state_layout, off = {}, 0
        for name, shape, dt in spec:
            n = 1
            for d in shape:
                n *= d
            self.state_layout[name] = (off, shape, dt)
            off += (n * torch.empty((), dtype=dt).element_size() + 255) // 256 * 256
        self.state_arena = torch.zeros(off, dtype=torch.uint8, device=dev)
        for name, (o, shape, dt) in self.state_layout.items():
            n = 1
            for d in shape:
                n *= d
            nbytes = n * torch.empty((), dtype=dt).element_size()
            setattr(self, name, self.state_arena[o:o + nbytes].view(dt).view(shape))
        self.poses[..., 6] = 1.0
        self.pair_key = torch.zeros(self.E, dtype=i64, device=dev)
        # the recurrent hidden state: float32 like the reference's (devo.py:232-233,308-316), held in the tile layout the
        # fused update operator reads and writes in place; the cuBLAS comparison paths keep a row-major copy
        self.state = GruState(self.E, dev, dim) if self.gru_mode == "mma" else None
        self.net = None if self.gru_mode == "mma" else torch.zeros(1, self.E, dim, dtype=torch.float32, device=dev)
        self.imap = torch.zeros(1, self.Np, dim, dtype=feat_dtype, device=dev)
        self.gmap_pm = torch.zeros(self.Np, 9, C, dtype=feat_dtype, device=dev)
        self.levels_pm = [torch.zeros(self.Nf, H // s, W // s, C, dtype=feat_dtype, device=dev) for s in self.levels]
        # correlation features, rows zero-padded from 441*L to a multiple of 64 so the first Linear of the corr MLP
        # (K = 882) runs as an aligned tensor-core GEMM; the lookup kernel never touches the padding
        self.corr_k = 441 * len(self.levels)
        self.corr_ld = (self.corr_k + 63) // 64 * 64
        self.corr_buf = torch.zeros(self.E, self.corr_ld, dtype=feat_dtype, device=dev)
        self.lmbda = torch.as_tensor([1e-4], dtype=f32, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.status_sticky = torch.zeros(1, dtype=torch.int32, device=dev)
        self.plan_kk = None
        self.plan_ij = None
        self.zeros_e = torch.zeros(self.E, dtype=i64, device=dev)
        self.fc = FrozenCast(feat_dtype)
        self.packed = PackedUpdateWeights(update, feat_dtype, self.corr_ld) if (fused_gru and dim == 384 and gru == "mma") else None
        self._gru_ws = (torch.empty(_lib.lib().devo_gru_workspace(self.E, max(self.Np, self.Nf * self.Nf)), dtype=torch.uint8, device=dev)
                        if self.gru_mode == "mma" else None)
        self._side = torch.cuda.Stream(device=dev)
        self._side2 = torch.cuda.Stream(device=dev)
        self._side3 = torch.cuda.Stream(device=dev)
        self._side4 = torch.cuda.Stream(device=dev)
        self._ingest_pending = False
        self._ba_ws = torch.empty(_lib.lib().devo_ba_workspace(self.E, max(self.t1 - self.t0, 0)), dtype=torch.uint8, device=dev)
        self._graph = None
        self._pristine = None
        self.tile_local = False
        self.delta = None
        self.weight = None
        self.coords = None

    # ---- state -------------------------------------------------------------------------------
    def set_graph(self, ii, jj, kk):
        """install the edge list (patch kk observed from frame ii in frame jj)"""
        self.ii.copy_(ii)
        self.jj.copy_(jj)
        self.kk.copy_(kk)
        self.refresh_pair_key()
        if self.plan_kk is None:
            self.plan_kk = cuda_ba.GraphPlan(self.kk, self.jj, self.Np, self.Nf)
            self.plan_ij = cuda_ba.GraphPlan(self.pair_key, self.zeros_e, self.Nf * self.Nf, 1, want_neighbors=False)
        else:
            self.plan_kk.update()
        # does every neighbour link stay inside its 64-edge tile (a patch-major list, e.g. the all-pairs graph of
        # enet.py:300-301)?  Then the update operator runs its first three programs as one launch.  Checked HERE, once per
        # installed edge list (one host synchronisation), never per update.
        from .update import tile_local_graph
        self.tile_local = self.gru_mode == "mma" and tile_local_graph(self.plan_kk)

    def refresh_pair_key(self, same_graph=False, overlap=False):
        """(re)compute the frame-pair key of SoftAgg's second grouping from ii / jj -- after set_graph, or after the caller
        refreshed `state_arena` (which carries the edge list).  `same_graph`: the caller vouches that the refreshed edge
        list is the one set_graph installed; otherwise what set_graph verified about it (`tile_local`) is dropped.
        The reference uses ii * 12345 + jj (enet.py:96); any key that orders the pairs the same way gives the same groups,
        and ii * Nf + jj needs 6 bits instead of 17: half the radix passes of the plan, whose width the engine fixes
        through the bound Nf * Nf it passes to GraphPlan."""
        if overlap:
            # the key's only reader is the frame-pair plan, which the next iteration builds on `_side2`: computed there, the
            # key is off the critical path (reprojection and lookup do not wait for it)
            self._side2.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self._side2):
                torch.add(self.jj, self.ii, alpha=self.Nf, out=self.pair_key)
        else:
            torch.add(self.jj, self.ii, alpha=self.Nf, out=self.pair_key)
        if not same_graph:
            self.tile_local = False

    def ingest_frame(self, idx, fmap, gmap_patches=None, imap_patches=None, overlap=False, only_levels=None):
        """fmap [C,H,W] planar features of frame `idx` -> all pixel-major pyramid levels;
        gmap_patches [M,C,3,3], imap_patches [M,dim] -> patch feature buffers.
        overlap=True: the packing runs on a side stream; the next iteration joins it right before the correlation
        lookup (its first consumer), so it overlaps the graph analysis and the reprojection."""
        if overlap:
            # two side streams: the full-resolution level on one, the pooled levels + patch features on the other.  The
            # lookup is the first consumer of both, and 20 us of packing in a row on ONE stream were longer than the reset
            # + reprojection they are meant to hide behind (12 us): the ingest sat on the critical path of the step.
            cur = torch.cuda.current_stream(self.device)
            self._side3.wait_stream(cur)
            self._side4.wait_stream(cur)
            fused2 = len(self.levels) == 2 and self.levels[0] == 1       # both levels from one read of the frame
            with torch.cuda.stream(self._side3):
                self.ingest_frame(idx, fmap, None, None, overlap=False, only_levels=((0, 1) if fused2 else (0,)))
            with torch.cuda.stream(self._side4):
                self.ingest_frame(idx, fmap, gmap_patches, imap_patches, overlap=False,
                                  only_levels=(() if fused2 else tuple(range(1, len(self.levels)))))
            for t in (fmap, gmap_patches, imap_patches):
                if t is not None and t.is_cuda:          # the caller may drop its inputs right away: the caching allocator must
                    t.record_stream(self._side3)         # not hand their memory out while the packing kernels still read it
                    t.record_stream(self._side4)
            self._ingest_pending = True
            return
        f = fmap.reshape(1, self.C, self.H, self.W).to(self.feat_dtype)
        want = [l for l in range(len(self.levels)) if only_levels is None or l in only_levels]
        if want == [0, 1] and len(self.levels) == 2 and self.levels[0] == 1 and cuda_corr.pack_pixel_major2(
                f, self.levels[1], self.levels_pm[0][idx:idx + 1], self.levels_pm[1][idx:idx + 1]):
            want = []                              # the [1, s] pyramid of DEVO: one kernel, one read of the frame
        for l in want:                             # packed straight into the ring-buffer slot (no staging copy)
            cuda_corr.pack_pixel_major(f, self.levels[l], out=self.levels_pm[l][idx:idx + 1])
        if gmap_patches is not None:
            cuda_corr.pack_gmap(gmap_patches.to(self.feat_dtype), out=self.gmap_pm[idx * self.M:(idx + 1) * self.M])
        if imap_patches is not None:
            src, dst = imap_patches.to(self.feat_dtype), self.imap[0, idx * self.M:(idx + 1) * self.M]
            if src.is_cuda and src.is_contiguous() and src.shape == dst.shape:
                _lib.copy_(dst, src)               # a kernel: no copy-engine node inside a captured step
            else:
                dst.copy_(src)

    def set_net(self, net):
        """install the recurrent hidden state ([1,E,dim], any float dtype; kept as float32)"""
        if self.state is not None:
            self.state.set(net)
        else:
            self.net.copy_(net)

    def get_net(self):
        """the recurrent hidden state as a row-major float32 [1,E,dim] tensor"""
        return self.state.get() if self.state is not None else self.net

    def snapshot_geometry(self):
        # poses and patches are the first two blocks of the arena: one contiguous snapshot, restored by ONE copy
        o, shape, dt = self.state_layout["patches"]
        self._geom_bytes = o + self.patches.numel() * self.patches.element_size()
        self._pristine = self.state_arena[:self._geom_bytes].clone()

    def pristine_geometry(self):
        """(poses, patches) views of the snapshot taken by snapshot_geometry()"""
        po, ps, pd = self.state_layout["poses"]
        xo, xs, xd = self.state_layout["patches"]
        return (self._pristine[po:po + self.poses.numel() * 4].view(pd).view(ps),
                self._pristine[xo:xo + self.patches.numel() * 4].view(xd).view(xs))

    # ---- one iteration -----------------------------------------------------------------------
    def _iteration(self, reset_geometry=False, marks=None):
        """`marks`: optional list of 5 CUDA events recorded on the current stream at the stage boundaries (start, after the
        reprojection, after the lookup, after the update operator, after fastba) -- bench.py times the stages inside the
        captured step with them (events created with external=True become event-record nodes of the graph)."""
        def mark(i):
            if marks is not None:
                marks[i].record(torch.cuda.current_stream(self.device))
        mark(0)
        if reset_geometry and self._pristine is not None:
            _lib.copy_(self.state_arena[:self._geom_bytes], self._pristine)
        # (0) graph analysis on the device (neighbours, patch groups, frame-pair groups) on a side stream:
        #     it only depends on ii/jj/kk, so it overlaps the reprojection and the correlation lookup
        cur = torch.cuda.current_stream(self.device)
        self._side.wait_stream(cur)
        self._side2.wait_stream(cur)
        with torch.cuda.stream(self._side):
            self.plan_kk.update()
            # the two memsets of the BA call, here instead of between the update operator and the first Gauss-Newton launch
            cuda_ba.prepare(self.E, self.t1 - self.t0, self.status, self._ba_ws)
        with torch.cuda.stream(self._side2):
            self.plan_ij.update()
        # (1) reproject: [1,E,2,3,3]
        coords = pops.transform_fused(self.poses, self.patches, self.intrinsics, self.ii, self.jj, self.kk, layout=1)
        mark(1)
        # (2) correlation lookup over all levels, output already in the GRU's [E, 882] layout
        if self._ingest_pending:                    # an overlapped frame ingest: the lookup is its first consumer
            cur.wait_stream(self._side3)
            cur.wait_stream(self._side4)
            self._ingest_pending = False
        cuda_corr.lookup_fused(self.gmap_pm, self.levels_pm, self.levels, coords[0], self.kk, self.jj, out=self.corr_buf)
        corr = self.corr_buf if self.fused_gru else self.corr_buf[:, :self.corr_k]
        mark(2)
        cur.wait_stream(self._side)
        cur.wait_stream(self._side2)
        # (3) GRU: cached fp16 weights, autocast-identical dtype flow, no host sync
        target = None
        if self.gru_mode == "mma":      # target / weight for BA come out of the heads epilogue of the same launch
            _, (delta, weight16, (target, weight)) = self.update.forward_mma(
                None, self.imap, self.kk, self.corr_buf, self.plan_kk, self.plan_ij, self.Np, self.Nf * self.Nf,
                self.packed, workspace=self._gru_ws, coords=coords, state=self.state, tile_local=self.tile_local)
        elif self.fused_gru:
            ctx = self.imap[:, self.kk]
            net, (delta, weight, _) = self.update.forward_fused(
                self.net.to(self.feat_dtype), ctx, corr.view(1, self.E, -1), self.plan_kk, self.plan_ij, self.Np, self.Nf * self.Nf,
                self.fc)
            self.net.copy_(net)
        else:
            ctx = self.imap[:, self.kk]
            net, (delta, weight, _) = self.update.forward_planned(
                self.net, ctx, corr.reshape(1, self.E, -1), self.plan_kk, self.plan_ij, self.Np, self.Nf * self.Nf, self.fc)
            self.net.copy_(net)
        mark(3)
        # (4) BA targets and in-place Gauss-Newton (reuses the kk/jj plan: one sort serves neighbours,
        #     SoftAgg and the Schur grouping)
        if target is None:
            target = coords[:, :, :, 1, 1] + delta.float()
            weight = weight.float()
        cuda_ba.forward_async(self.poses, self.patches, self.intrinsics, target, weight, self.lmbda,
                              self.ii, self.jj, self.kk, self.t0, self.t1, self.ba_iterations, status=self.status,
                              plan=self.plan_kk, workspace=self._ba_ws, prepared=True, status_or=self.status_sticky)
        # (status_sticky: non-zero once any iteration failed; OR-ed in by the last launch of the BA)
        mark(4)
        self.coords, self.delta, self.weight = coords, delta, weight

    @torch.no_grad()
    def step(self, reset_geometry=False):
        self._iteration(reset_geometry)

    @torch.no_grad()
    def capture(self, reset_geometry=False, warmup=3):
        """capture one iteration into a CUDA graph (after `warmup` eager iterations)"""
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._iteration(reset_geometry)
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._iteration(reset_geometry)
        self._graph = g
        return g

    def replay(self):
        if self._ingest_pending:                         # an overlapped ingest issued OUTSIDE the captured graph: join it here
            torch.cuda.current_stream(self.device).wait_stream(self._side3)
            torch.cuda.current_stream(self.device).wait_stream(self._side4)
            self._ingest_pending = False
        self._graph.replay()
