"""PatchGraphVO -- the per-frame loop of DEVO (devo/devo.py:21-554: ring-buffer ingest, motion model, edge bookkeeping,
update, keyframing) around the fused B200 update operator.

What differs from running the reference's `DEVO` class on the shims (which also works, tests/test_gpu_reference_callers.py):

  * frame ingest (devo.py:523-527): the new frame's matching features go straight into PIXEL-MAJOR pyramid ring buffers
    (`devo_pyramid_pack`, one launch per level, avg-pool fused) and `[patch][9][C]` patch features -- the layout the
    TMA/tcgen05 lookup reads; the reference's `kk % (M*mem)` / `jj % mem` ring indexing is kept (devo.py:213-214)
  * update (devo.py:308-338): fused reprojection (1 launch), fused multi-level lookup (1), device-side graph analysis (2,
    no host round trip: the reference's `neighbors` copies the edge list to the host and sorts there), the fused update
    operator (7 launches) on a float32 tile-layout hidden state, fastba (3) -- no host synchronisation inside `update()`
  * edge bookkeeping (devo.py:225-239): `append_factors` / `remove_factors` gather the hidden state on the device
    (`devo_gru_state_gather`) instead of `torch.cat` / boolean indexing of a [1,E,384] tensor
  * keyframing (devo.py:258-306) needs one host decision per frame (the edge count changes), like the reference; the
    flow magnitude itself is three fused reprojections

The frame front end (event-voxel encoders + patch selector, devo/enet.py:103-200) is outside the update-operator hot
path: `patchify(image)` is any callable returning the per-frame features (the reference's own `eVONet.patchify` wrapped by
`frontend_from_reference_network`, or a synthetic generator for benchmarks).
"""
import torch

from . import cuda_ba, cuda_corr, projective_ops as pops  # noqa: I001
from .lietorch import SE3
from .update import GruState, PackedUpdateWeights


class VOConfig:
    """the fields of devo/config.py + config/default.yaml the loop reads"""
    PATCHES_PER_FRAME = 96
    BUFFER_SIZE = 4096
    REMOVAL_WINDOW = 22
    OPTIMIZATION_WINDOW = 10
    PATCH_LIFETIME = 13
    KEYFRAME_INDEX = 4
    KEYFRAME_THRESH = 15.0
    MOTION_MODEL = "DAMPED_LINEAR"
    MOTION_DAMPING = 0.5
    MIXED_PRECISION = True
    NORM = "std"

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    @classmethod
    def from_cfg(cls, cfg):
        return cls(**{k: getattr(cfg, k) for k in dir(cls) if k.isupper() and hasattr(cfg, k)})


def frontend_from_reference_network(network, cfg):
    """wrap the reference's eVONet.patchify (devo/enet.py:122-200) as the front end of PatchGraphVO"""
    def patchify(image):
        with torch.autocast("cuda", enabled=bool(cfg.MIXED_PRECISION)):
            fmap, gmap, imap, patches, _, clr = network.patchify(
                image[None, None], patches_per_image=cfg.PATCHES_PER_FRAME, return_color=True,
                scorer_eval_mode=cfg.SCORER_EVAL_MODE, scorer_eval_use_grid=cfg.SCORER_EVAL_USE_GRID)
        return dict(fmap=fmap[0, 0], gmap=gmap[0], imap=imap[0, :, :, 0, 0], patches=patches[0], clr=clr)
    return patchify


def frontend_from_encoders(fnet, inet, cfg, select=None):
    """front end of PatchGraphVO on this library's fused patch gather (devo_b200/frontend.py): `fnet` / `inet` are the two
    encoders (any callables image [1,1,B,H,W] -> [1,1,C,H/4,W/4]; the reference divides their outputs by 4, enet.py:125-126),
    `select(image, n_patches)` -> centres [1,M,2] (default: the reference's RANDOM selection)."""
    from .frontend import gather_patches

    def patchify(image):
        with torch.autocast("cuda", enabled=bool(cfg.MIXED_PRECISION)):
            fmap = fnet(image[None, None]) / 4.0
            imap = inet(image[None, None]) / 4.0
        fmap, imap = fmap[0], imap[0].to(fmap.dtype)
        h, w = fmap.shape[-2:]
        M = cfg.PATCHES_PER_FRAME
        if select is None:
            x = torch.randint(1, w - 1, size=[1, M], device=fmap.device)
            y = torch.randint(1, h - 1, size=[1, M], device=fmap.device)
            coords = torch.stack([x, y], dim=-1).float()
        else:
            coords = select(image, M)
        g, gpm, im, pt = gather_patches(fmap, imap, coords, None, 3, planar=False, pixel_major=True)
        return dict(fmap=fmap[0], gmap=None, gmap_pm=gpm, imap=im, patches=pt, clr=None)
    return patchify


class PatchGraphVO:
    def __init__(self, cfg, update, patchify, ht=480, wd=640, dim_inet=384, dim_fnet=128, P=3, RES=4.0, mem=32,
                 levels=(1, 4), device="cuda", edge_capacity=65536):
        self.cfg = cfg if isinstance(cfg, VOConfig) else VOConfig.from_cfg(cfg)
        self.update_module = update.to(device).eval()
        self.patchify = patchify
        self.device = dev = torch.device(device)
        self.P, self.RES, self.mem, self.levels = P, RES, mem, tuple(levels)
        self.dim, self.C = dim_inet, dim_fnet
        self.M, self.N = self.cfg.PATCHES_PER_FRAME, self.cfg.BUFFER_SIZE
        self.dt = torch.float16 if self.cfg.MIXED_PRECISION else torch.float16      # the fused path is 16-bit only
        self.n = self.m = self.counter = 0
        self.is_initialized = False
        self.tlist, self.delta = [], {}
        self.H4, self.W4 = int(ht // RES), int(wd // RES)
        f32, i64 = torch.float32, torch.int64
        M, N = self.M, self.N
        self.tstamps_ = torch.zeros(N, dtype=i64, device=dev)
        self.poses_ = torch.zeros(N, 7, dtype=f32, device=dev)
        self.poses_[:, 6] = 1.0
        self.patches_ = torch.zeros(N, M, 3, P, P, dtype=f32, device=dev)
        self.intrinsics_ = torch.zeros(N, 4, dtype=f32, device=dev)
        self.index_ = torch.zeros(N, M, dtype=i64, device=dev)
        # ring buffers, in the layouts the lookup kernel reads
        self.imap_ = torch.zeros(mem * M, dim_inet, dtype=self.dt, device=dev)
        self.gmap_pm = torch.zeros(mem * M, P * P, dim_fnet, dtype=self.dt, device=dev)
        self.levels_pm = [torch.zeros(mem, self.H4 // s, self.W4 // s, dim_fnet, dtype=self.dt, device=dev) for s in self.levels]
        self.ii = torch.zeros(0, dtype=i64, device=dev)
        self.jj = torch.zeros(0, dtype=i64, device=dev)
        self.kk = torch.zeros(0, dtype=i64, device=dev)
        self.state = GruState(0, dev, dim_inet)
        self.first_update = True                   # the very first update sees a half zero-state (devo.py:84)
        self._graph_epoch, self._plan_epoch = 0, -1  # bumped by append / remove_factors; the plans are rebuilt when they differ
        self.corr_ld = (441 * len(self.levels) + 63) // 64 * 64
        self.cap = int(edge_capacity)
        self.corr_buf = torch.zeros(self.cap, self.corr_ld, dtype=self.dt, device=dev)
        self.packed = PackedUpdateWeights(self.update_module, self.dt, self.corr_ld)
        zero1 = torch.zeros(1, dtype=i64, device=dev)
        self.plan_kk = cuda_ba.GraphPlan(zero1, zero1, N * M, N, capacity=self.cap)
        self.plan_ij = cuda_ba.GraphPlan(zero1, zero1, -1, 1, want_neighbors=False, capacity=self.cap)
        self.lmbda = torch.as_tensor([1e-4], dtype=f32, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.n_updates = 0
        self._Id = SE3.Identity(1, device=dev)

    # ---- views like the reference's properties (devo.py:151-181)
    @property
    def poses(self):
        return self.poses_.view(1, self.N, 7)

    @property
    def patches(self):
        return self.patches_.view(1, self.N * self.M, 3, self.P, self.P)

    @property
    def intrinsics(self):
        return self.intrinsics_.view(1, self.N, 4)

    @property
    def ix(self):
        return self.index_.view(-1)

    @property
    def imap(self):
        return self.imap_.view(1, self.mem * self.M, self.dim)

    # ---- edge bookkeeping (devo.py:225-239)
    def append_factors(self, ii, jj):
        E0 = self.ii.numel()
        self.jj = torch.cat([self.jj, jj])
        self.kk = torch.cat([self.kk, ii])
        self.ii = torch.cat([self.ii, self.ix[ii]])
        if self.ii.numel() > self.cap:
            raise RuntimeError("PatchGraphVO: edge capacity %d exceeded" % self.cap)
        idx = torch.cat([torch.arange(E0, device=self.device), torch.full((ii.numel(),), -1, dtype=torch.int64, device=self.device)])
        self.state = self.state.gather(idx)                    # new edges start from a zero hidden state
        self._graph_epoch += 1

    def remove_factors(self, m):
        keep = torch.nonzero(~m).view(-1)
        self.ii, self.jj, self.kk = self.ii[keep], self.jj[keep], self.kk[keep]
        self.state = self.state.gather(keep)
        self._graph_epoch += 1

    # ---- the unit of work (devo.py:210-223, 308-338)
    def _lookup(self, coords, kk, jj, E):
        buf = self.corr_buf[:E]
        cuda_corr.lookup_fused(self.gmap_pm, self.levels_pm, self.levels, coords[0], kk % (self.M * self.mem), jj % self.mem, out=buf)
        return buf

    def _run_update(self, ii, jj, kk, state, net16, plan_kk, plan_ij, max_patches, max_pairs):
        E = ii.numel()
        coords = pops.transform_fused(self.poses, self.patches, self.intrinsics, ii, jj, kk, layout=1)
        corr = self._lookup(coords, kk, jj, E)
        _, (delta, weight16, (target, weight)) = self.update_module.forward_mma(
            net16, self.imap_.view(1, -1, self.dim), kk % (self.M * self.mem), corr, plan_kk, plan_ij, max_patches, max_pairs,
            self.packed, coords=coords, state=state)
        return coords, delta, target, weight

    @torch.no_grad()
    def update(self):
        E = self.ii.numel()
        if self._plan_epoch != self._graph_epoch:              # the edge list only changes in append / remove_factors: the 12
            self.plan_kk.rebind(self.kk, self.jj)              # updates of the initialisation (devo.py:536-538) share one analysis
            self.plan_ij.rebind(self.ii * 12345 + self.jj, torch.zeros_like(self.ii))
            self._plan_epoch = self._graph_epoch
        net16 = torch.zeros(1, E, self.dim, dtype=self.dt, device=self.device) if self.first_update else None
        self.first_update = False
        n_act = self.n
        coords, delta, target, weight = self._run_update(self.ii, self.jj, self.kk, self.state, net16, self.plan_kk, self.plan_ij,
                                                         max(self.m, 1), max(min(E, n_act * n_act), 1))
        t0 = self.n - self.cfg.OPTIMIZATION_WINDOW if self.is_initialized else 1
        t0 = max(t0, 1)
        cuda_ba.forward_async(self.poses, self.patches, self.intrinsics, target, weight, self.lmbda, self.ii, self.jj, self.kk,
                              t0, self.n, 2, status=self.status, plan=self.plan_kk)
        self.n_updates += 1

    @torch.no_grad()
    def motion_probe(self):
        """devo.py:241-256: median predicted flow of the newest active patches into the frame that just arrived (its
        features are already in the ring), from a zero hidden state"""
        kk = torch.arange(self.m - self.M, self.m, device=self.device)
        jj = self.n * torch.ones_like(kk)
        ii = self.ix[kk]
        E = kk.numel()
        plan_kk = cuda_ba.GraphPlan(kk, jj, self.N * self.M, self.N)
        plan_ij = cuda_ba.GraphPlan(ii * 12345 + jj, torch.zeros_like(ii), -1, 1, want_neighbors=False)
        net16 = torch.zeros(1, E, self.dim, dtype=self.dt, device=self.device)
        _, delta, _, _ = self._run_update(ii, jj, kk, GruState(E, self.device, self.dim), net16, plan_kk, plan_ij, E, E)
        return torch.quantile(delta.norm(dim=-1).float(), 0.5)

    def motionmag(self, i, j):
        k = (self.ii == i) & (self.jj == j)
        flow = pops.flow_mag(SE3(self.poses), self.patches, self.intrinsics, self.ii[k], self.jj[k], self.kk[k], beta=0.5)
        return flow.mean().item()

    @torch.no_grad()
    def keyframe(self):
        """devo.py:258-306"""
        cfg, mem = self.cfg, self.mem
        i = self.n - cfg.KEYFRAME_INDEX - 1
        j = self.n - cfg.KEYFRAME_INDEX + 1
        m = self.motionmag(i, j) + self.motionmag(j, i)
        if m / 2 < cfg.KEYFRAME_THRESH:
            k = self.n - cfg.KEYFRAME_INDEX
            t0 = self.tstamps_[k - 1].item()
            t1 = self.tstamps_[k].item()
            dP = SE3(self.poses_[k]) * SE3(self.poses_[k - 1]).inv()
            self.delta[t1] = (t0, dP)
            self.remove_factors((self.ii == k) | (self.jj == k))
            self.kk[self.ii > k] -= self.M
            self.ii[self.ii > k] -= 1
            self.jj[self.jj > k] -= 1
            n = self.n
            for buf in (self.tstamps_, self.poses_, self.patches_, self.intrinsics_):
                buf[k:n - 1] = buf[k + 1:n].clone()
            for f in range(k, n - 1):                        # ring slots: (f % mem) <- ((f+1) % mem)
                a, b = f % mem, (f + 1) % mem
                self.imap_[a * self.M:(a + 1) * self.M] = self.imap_[b * self.M:(b + 1) * self.M]
                self.gmap_pm[a * self.M:(a + 1) * self.M] = self.gmap_pm[b * self.M:(b + 1) * self.M]
                for lv in self.levels_pm:
                    lv[a] = lv[b]
            self.n -= 1
            self.m -= self.M
        self.remove_factors(self.ix[self.kk] < self.n - cfg.REMOVAL_WINDOW)

    # ---- edges of a new frame (devo.py:360-380)
    def _edges_forw(self):
        r = self.cfg.PATCH_LIFETIME
        t0, t1 = self.M * max(self.n - r, 0), self.M * max(self.n - 1, 0)
        kk, jj = torch.meshgrid(torch.arange(t0, t1, device=self.device), torch.arange(self.n - 1, self.n, device=self.device), indexing="ij")
        return kk.reshape(-1), jj.reshape(-1)

    def _edges_back(self):
        r = self.cfg.PATCH_LIFETIME
        t0, t1 = self.M * max(self.n - 1, 0), self.M * max(self.n, 0)
        kk, jj = torch.meshgrid(torch.arange(t0, t1, device=self.device), torch.arange(max(self.n - r, 0), self.n, device=self.device), indexing="ij")
        return kk.reshape(-1), jj.reshape(-1)

    # ---- one frame (devo.py:382-554)
    @torch.no_grad()
    def __call__(self, tstamp, image, intrinsics, scale=1.0):
        if (self.n + 1) >= self.N:
            raise RuntimeError("PatchGraphVO: keyframe buffer too small")
        if torch.is_tensor(image):                        # voxel normalisation of the frame (devo.py:419-452), on the device
            from . import voxel
            image = voxel.normalize_frame(image[None, None].float(), self.cfg.NORM)
            if image is None:
                return                                    # 'rescale' with an empty polarity: the reference skips the frame
            image = image[0, 0]
        fe = self.patchify(image)
        n, M, mem = self.n, self.M, self.mem
        self.tlist.append(tstamp)
        self.tstamps_[n] = self.counter
        self.intrinsics_[n] = intrinsics / self.RES
        self.index_[n + 1] = n + 1
        if n > 1:
            if self.cfg.MOTION_MODEL == "DAMPED_LINEAR":
                P1, P2 = SE3(self.poses_[n - 1]), SE3(self.poses_[n - 2])
                xi = self.cfg.MOTION_DAMPING * (P1 * P2.inv()).log()
                self.poses_[n] = (SE3.exp(xi) * P1).data
            else:
                self.poses_[n] = self.poses_[n - 1]
        patches = fe["patches"].to(torch.float32).view(1, M, 3, self.P, self.P).clone()
        patches[:, :, 2] = torch.rand_like(patches[:, :, 2, 0, 0, None, None])      # devo.py:511
        if self.is_initialized:
            patches[:, :, 2] = torch.median(self.patches_[n - 3:n, :, 2])
        self.patches_[n] = patches[0]
        # ---- ring-buffer ingest in the lookup's layouts (replaces devo.py:523-527 + avg_pool2d)
        slot = n % mem
        self.imap_[slot * M:(slot + 1) * M] = fe["imap"].reshape(M, self.dim).to(self.dt)
        if fe.get("gmap_pm") is not None:             # a front end on devo_b200.frontend hands the pixel-major patches over directly
            self.gmap_pm[slot * M:(slot + 1) * M] = fe["gmap_pm"].reshape(M, self.P * self.P, self.C).to(self.dt)
        else:
            cuda_corr.pack_gmap(fe["gmap"].reshape(M, self.C, self.P, self.P).to(self.dt), out=self.gmap_pm[slot * M:(slot + 1) * M])
        fmap = fe["fmap"].reshape(1, self.C, self.H4, self.W4).to(self.dt)
        if not (len(self.levels) == 2 and self.levels[0] == 1 and cuda_corr.pack_pixel_major2(
                fmap, self.levels[1], self.levels_pm[0][slot:slot + 1], self.levels_pm[1][slot:slot + 1])):
            for l, s in enumerate(self.levels):
                cuda_corr.pack_pixel_major(fmap, s, out=self.levels_pm[l][slot:slot + 1])
        self.counter += 1
        if self.n > 0 and not self.is_initialized:
            thres = 2.0 if scale == 1.0 else scale ** 2
            if self.motion_probe() < thres:
                self.delta[self.counter - 1] = (self.counter - 2, self._Id[0])
                return
        self.n += 1
        self.m += M
        self.append_factors(*self._edges_forw())
        self.append_factors(*self._edges_back())
        if self.n == 8 and not self.is_initialized:
            self.is_initialized = True
            for _ in range(12):
                self.update()
        elif self.is_initialized:
            self.update()
            self.keyframe()

    # ---- trajectory (devo.py:183-208)
    def _get_pose(self, t, traj):
        if t in traj:
            return SE3(traj[t])
        t0, dP = self.delta[t]
        return dP * self._get_pose(t0, traj)

    def terminate(self):
        import numpy as np
        from . import lietorch
        traj = {self.tstamps_[i].item(): self.poses_[i] for i in range(self.n)}
        poses = lietorch.stack([self._get_pose(t, traj) for t in range(self.counter)], dim=0)
        return poses.inv().data.cpu().numpy(), np.array(self.tlist, dtype=np.float64)
