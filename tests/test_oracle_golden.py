"""CPU: oracle/pops.py and oracle/ba.py against golden vectors produced by the reference's own
Python code (tests/golden/make_golden.py), and live against /root/reference when present."""
import os
import sys

import pytest
import torch

from oracle import ba as oba
from oracle import pops as opops

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_transform_golden():
    g = torch.load(os.path.join(GOLD, "transform_3x8.pt"))
    P = g["problem"]
    a = (P["poses0"], P["patches0"], P["intrinsics"], P["ii"], P["jj"], P["kk"])
    assert torch.equal(opops.transform(*a), g["coords"])
    assert torch.equal(opops.transform(*a, tonly=True), g["coords_tonly"])
    c, v = opops.transform(*a, valid=True)
    assert torch.equal(v, g["valid"])
    c, v, (Ji, Jj, Jz) = opops.transform(*a, jacobian=True)
    assert torch.allclose(Ji, g["Ji"], atol=1e-12) and torch.allclose(Jj, g["Jj"], atol=1e-12)
    assert torch.allclose(Jz, g["Jz"], atol=1e-12)
    assert torch.allclose(opops.flow_mag(*a, beta=0.5), g["flow_mag"], atol=1e-12)
    pc = opops.point_cloud(P["poses0"], P["patches0"][:, :8], P["intrinsics"], torch.zeros(8, dtype=torch.long))
    assert torch.allclose(pc, g["point_cloud"], atol=1e-12)


def test_ba_config1_golden():
    """BASELINE.json config 1: ba.py Gauss-Newton, 2 frames x 32 patches, ep=10, fixedp=1"""
    g = torch.load(os.path.join(GOLD, "ba_config1.pt"))
    P = g["problem"]
    poses, patches = P["poses0"].clone(), P["patches0"].clone()
    for it, (pg, dg) in enumerate(g["traj"]):
        poses, patches = oba.ba_step(poses, patches, P["intrinsics"], P["targets"], P["weights"], 1e-4,
                                     P["ii"], P["jj"], P["kk"], P["bounds"], ep=10.0, fixedp=1)
        assert torch.allclose(poses, pg, atol=1e-10), it
        assert torch.allclose(patches[:, :, 2, 1, 1], dg, atol=1e-10), it
    # Gauss-Newton actually converged towards the noisy targets
    c = opops.transform(poses, patches, P["intrinsics"], P["ii"], P["jj"], P["kk"])
    assert (c[..., 1, 1, :] - P["targets"]).norm(dim=-1).mean() < 1.0


def test_ba_bounds_and_structure_only_golden():
    g = torch.load(os.path.join(GOLD, "ba_4x12.pt"))
    P = g["problem"]
    a = (P["intrinsics"], P["targets"], P["weights"], 1e-4, P["ii"], P["jj"], P["kk"])
    p1, x1 = oba.ba_step(P["poses0"], P["patches0"], *a, [20, 20, 140, 100], ep=10.0, fixedp=1)
    assert torch.allclose(p1, g["poses_a"], atol=1e-10) and torch.allclose(x1[:, :, 2, 1, 1], g["depth_a"], atol=1e-10)
    p2, x2 = oba.ba_step(P["poses0"], P["patches0"], *a, P["bounds"], ep=100.0, fixedp=2, structure_only=True)
    assert torch.allclose(p2, g["poses_b"], atol=1e-12) and torch.allclose(x2[:, :, 2, 1, 1], g["depth_b"], atol=1e-10)


@pytest.mark.skipif(not os.path.isdir("/root/reference/devo"), reason="reference checkout not present")
def test_live_against_reference_python():
    sys.path.insert(0, GOLD)
    import ref_import
    from problems import ba_problem
    lt, rpops, rba = ref_import.load()
    P = ba_problem(n_frames=5, patches_per_frame=10, seed=99, init="perturbed")
    a = (P["patches0"], P["intrinsics"], P["ii"], P["jj"], P["kk"])
    c1, v1, J1 = opops.transform(P["poses0"], *a, jacobian=True)
    c2, v2, J2 = rpops.transform(lt.SE3(P["poses0"]), *a, jacobian=True)
    assert torch.equal(c1, c2) and torch.equal(v1, v2)
    for x, y in zip(J1, J2):
        assert torch.allclose(x, y, atol=1e-12)
    p1, x1 = oba.ba_step(P["poses0"], P["patches0"], P["intrinsics"], P["targets"], P["weights"], 1e-4,
                         P["ii"], P["jj"], P["kk"], P["bounds"], ep=10.0, fixedp=1)
    p2, x2 = rba.BA(lt.SE3(P["poses0"]), P["patches0"], P["intrinsics"], P["targets"], P["weights"], 1e-4,
                    P["ii"], P["jj"], P["kk"], P["bounds"], ep=10.0, fixedp=1)
    assert torch.allclose(p1, p2.data, atol=1e-11) and torch.allclose(x1, x2, atol=1e-11)
