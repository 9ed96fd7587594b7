"""CPU-only, world_size 2 over gloo: the N>1 host logic (partition by signal, max-over-ranks timing)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_signals_partition(jw):
    from jwave_pro_b200.sharding import shard_signals
    for total in (0, 1, 7, 512, 4096, 8192, 1000003):
        for world in (1, 2, 3, 4, 8):
            got = [shard_signals(total, world, r) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == total
            for (s0, c0), (s1, _) in zip(got, got[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in got) - min(c for _, c in got) <= 1


def test_two_rank_gloo_run():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29613", os.path.join(ROOT, "tests", "_gloo_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["world"] == 2 and d["t_max"] == 2.0
    assert d["shards"]["4096"] == [0, 2048] and d["shards"]["7"] == [0, 3]
