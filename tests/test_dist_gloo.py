"""CPU, world_size 2 over gloo: the N>1 plumbing of bench.py (sequence sharding, max-over-ranks timing,
aggregate throughput) and the reference arm's "rank 0 only" rule."""
import json
import os
import subprocess
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from devo_b200 import dist as d
    r, w, _ = d.init("gloo")
    seqs = d.shard_sequences(5, r, w)
    t = d.max_over_ranks(0.010 * (rank + 1))                 # rank 1 is slower
    thr = d.aggregate_throughput(100, 0.010 * (rank + 1))
    q.put((rank, seqs, t, thr))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    assert out[0][1] == [0, 1, 2] and out[1][1] == [3, 4]            # balanced, disjoint, complete
    assert abs(out[0][2] - 0.020) < 1e-9 and abs(out[1][2] - 0.020) < 1e-9   # max over ranks on both
    assert abs(out[0][3] - 2 * 100 / 0.020) < 1e-6


def test_shard_properties():
    from devo_b200.dist import shard_sequences
    for n in (0, 1, 7, 8, 64):
        for w in (1, 2, 4, 8):
            parts = [shard_sequences(n, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""              # other ranks exit 0 without work
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "iterations/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


# ---- edge-sharded BA: the Schur complement distributes over patch-owning ranks ----------------------
def _schur_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from devo_b200 import dist as d
    from oracle import fastba as ofba
    from problems import ba_problem
    d.init("gloo")
    P = ba_problem(n_frames=4, patches_per_frame=10, seed=11, init="perturbed")
    Np = 40
    sel = d.shard_edges_by_patch(P["kk"], Np, rank, world)
    lm = torch.tensor([1e-4], dtype=torch.float64)
    _, _, st, s = ofba.ba(P["poses0"], P["patches0"][0], P["intrinsics"], P["targets"][:, sel], P["weights"][:, sel], lm,
                          P["ii"][sel], P["jj"][sel], P["kk"][sel], 1, 4, 1, return_system=True)
    Q = 1.0 / (s["C"] + lm)
    S = s["B"] - (s["E"] * Q[None]) @ s["E"].t()           # this rank's undamped partial
    y = s["v"] - (s["E"] * Q[None]) @ s["u"]
    buf = torch.cat([S.reshape(-1), y])
    dist.all_reduce(buf)                                    # the ONE collective of an iteration
    q.put((rank, sel.numel(), buf))
    dist.barrier()
    dist.destroy_process_group()


def test_schur_complement_distributes_over_patch_shards():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import fastba as ofba
    from problems import ba_problem
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29100 + (os.getpid() % 300)
    ps = [ctx.Process(target=_schur_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted([q.get(timeout=180) for _ in range(2)], key=lambda t: t[0])
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    P = ba_problem(n_frames=4, patches_per_frame=10, seed=11, init="perturbed")
    lm = torch.tensor([1e-4], dtype=torch.float64)
    _, _, _, s = ofba.ba(P["poses0"], P["patches0"][0], P["intrinsics"], P["targets"], P["weights"], lm,
                         P["ii"], P["jj"], P["kk"], 1, 4, 1, return_system=True)
    n = s["S"].shape[0]
    S_full = s["S"] - torch.eye(n, dtype=torch.float64) * ((s["S"].diagonal() - 1.0) / 1.0001 * 1e-4 + 1.0)   # undo the damping
    assert out[0][1] + out[1][1] == P["ii"].numel()                       # edges partitioned, none lost
    assert torch.equal(out[0][2], out[1][2])                              # every rank holds the same reduced system
    got_S, got_y = out[0][2][:n * n].view(n, n), out[0][2][n * n:]
    assert (got_S - S_full).abs().max() <= 1e-9 * S_full.abs().max()
    assert (got_y - s["y"]).abs().max() <= 1e-9 * s["y"].abs().max()


def test_patch_sharding_properties():
    from devo_b200.dist import patch_owner, patch_range, shard_edges_by_patch
    kk = torch.randint(0, 37, (500,))
    for w in (1, 2, 3, 8):
        own = patch_owner(37, w)
        parts = [shard_edges_by_patch(kk, 37, r, w) for r in range(w)]
        assert torch.equal(torch.sort(torch.cat(parts)).values, torch.arange(500))
        for r in range(w):
            lo, hi = patch_range(37, r, w)
            assert (own[lo:hi] == r).all() and (own[kk[parts[r]]] == r).all()
            assert (parts[r][1:] > parts[r][:-1]).all()                  # edge order preserved inside a shard
