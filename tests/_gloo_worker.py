"""Worker for tests/test_sharding_gloo.py: world_size-2 gloo run of the N>1 host logic used by bench.py."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from jwave_pro_b200.sharding import reduce_max, reduce_sum, shard_signals  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    out = {}
    for total in (4096, 8192, 7, 1, 0, 513):
        start, count = shard_signals(total, world, rank)
        owned = torch.zeros(max(total, 1), dtype=torch.int64)
        owned[start:start + count] += 1
        dist.all_reduce(owned, op=dist.ReduceOp.SUM)
        assert total == 0 or bool((owned[:total] == 1).all()), "every signal owned exactly once"
        out[str(total)] = [start, count]
        assert reduce_sum([count])[0] == total
    # per-rank "device time": the reported step time is the max over ranks, the value the sum of units / that time
    t_local = 1.0 + rank
    t_max = reduce_max([t_local])[0]
    assert t_max == float(world)
    dist.barrier()
    if rank == 0:
        print(json.dumps({"world": world, "shards": out, "t_max": t_max}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
