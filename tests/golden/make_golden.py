"""Generate golden vectors by running the REFERENCE's own Python code
(/root/reference/devo/{lietorch/*.py, projective_ops.py, ba.py}) in this container,
on top of the oracle's restatement of the missing native/third-party pieces
(tests/golden/ref_import.py).  The outputs pin oracle/pops.py and oracle/ba.py -- and the
product's projective_ops / ba / lietorch wrappers on the GPU box, where /root/reference
does not exist -- against the reference's Python logic.

    python tests/golden/make_golden.py        # writes tests/golden/*.pt (fp64, small)
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_import  # noqa: E402
from problems import ba_problem  # noqa: E402


def main():
    lt, pops, ba = ref_import.load()
    torch.manual_seed(7)
    dt = torch.float64

    # ---- (1) transform (+ Jacobians, tonly, valid) on a 3-frame problem
    P = ba_problem(n_frames=3, patches_per_frame=8, seed=11, init="perturbed")
    G = lt.SE3(P["poses0"])
    args = (P["patches0"], P["intrinsics"], P["ii"], P["jj"], P["kk"])
    c = pops.transform(G, *args)
    ct = pops.transform(G, *args, tonly=True)
    cv, v = pops.transform(G, *args, valid=True)
    cj, vj, (Ji, Jj, Jz) = pops.transform(G, *args, jacobian=True)
    fm = pops.flow_mag(G, *args, beta=0.5)
    pc = pops.point_cloud(G, P["patches0"][:, :8], P["intrinsics"], torch.zeros(8, dtype=torch.long))
    torch.save(dict(problem=P, coords=c, coords_tonly=ct, valid=v, Ji=Ji, Jj=Jj, Jz=Jz, flow_mag=fm, point_cloud=pc),
               os.path.join(HERE, "transform_3x8.pt"))

    # ---- (2) ba.py Gauss-Newton, BASELINE.json config 1 (2 frames x 32 patches), 2 and 10 steps
    P = ba_problem(n_frames=2, patches_per_frame=32, seed=1234)
    Gs, X = lt.SE3(P["poses0"].clone()), P["patches0"].clone()
    traj = []
    for it in range(10):
        Gs, X = ba.BA(Gs, X, P["intrinsics"], P["targets"], P["weights"], 1e-4, P["ii"], P["jj"], P["kk"],
                      P["bounds"], ep=10.0, fixedp=1)
        traj.append((Gs.data.clone(), X[:, :, 2, 1, 1].clone()))
    torch.save(dict(problem=P, traj=traj), os.path.join(HERE, "ba_config1.pt"))

    # structure-only + 4-frame problem with bounds active
    P = ba_problem(n_frames=4, patches_per_frame=12, seed=5, noise=2.0)
    Gs, X = lt.SE3(P["poses0"].clone()), P["patches0"].clone()
    Gs1, X1 = ba.BA(Gs, X, P["intrinsics"], P["targets"], P["weights"], 1e-4, P["ii"], P["jj"], P["kk"],
                    [20, 20, 140, 100], ep=10.0, fixedp=1)
    Gs2, X2 = ba.BA(Gs, X, P["intrinsics"], P["targets"], P["weights"], 1e-4, P["ii"], P["jj"], P["kk"],
                    P["bounds"], ep=100.0, fixedp=2, structure_only=True)
    torch.save(dict(problem=P, poses_a=Gs1.data, depth_a=X1[:, :, 2, 1, 1], poses_b=Gs2.data, depth_b=X2[:, :, 2, 1, 1]),
               os.path.join(HERE, "ba_4x12.pt"))

    # ---- (3) gradients through the reference's autograd wrappers (tangent-space convention)
    out = {}
    for Grp in (lt.SO3, lt.RxSO3, lt.SE3, lt.Sim3):
        K = Grp.manifold_dim
        a = (0.3 * torch.randn(5, K, dtype=dt)).requires_grad_(True)
        b = torch.randn(5, K, dtype=dt).requires_grad_(True)
        p = torch.randn(5, 3, dtype=dt).requires_grad_(True)
        X0 = Grp.exp(0.5 * torch.randn(5, K, dtype=dt))
        Y = Grp.exp(a) * X0
        f = (Y.inv().log() * torch.arange(1, K + 1, dtype=dt)).sum() + (Y.adjT(b) ** 2).sum() \
            + (Y.adj(b) * 0.3).sum() + (Y.act(p) ** 2).sum() + Y.matrix().sum() + (Y.vec() ** 2).sum()
        ga, gb, gp = torch.autograd.grad(f, [a, b, p])
        out[Grp.group_name] = dict(a=a.detach(), b=b.detach(), p=p.detach(), X0=X0.data, f=f.detach(),
                                   ga=ga, gb=gb, gp=gp, Y=Y.data.detach())
    torch.save(out, os.path.join(HERE, "lie_autograd.pt"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
