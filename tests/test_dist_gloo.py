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
