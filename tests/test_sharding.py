"""CPU: read sharding over ranks (world_size 2, gloo) -- the only multi-GPU logic of the path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sloika_b200 import sharding


def test_partition_is_a_balanced_cover():
    rng = np.random.default_rng(0)
    lengths = rng.integers(1000, 120000, size=101)
    for world in (1, 2, 4, 8):
        shards = sharding.partition_reads(lengths, world)
        assert sorted(i for s in shards for i in s) == list(range(len(lengths)))
        loads = [int(lengths[s].sum()) for s in shards]
        assert max(loads) - min(loads) <= lengths.max()
    assert sharding.partition_reads([], 2) == [[], []]
    assert sharding.partition_reads([5], 4) == [[0], [], [], []]
    covered = [sharding.shard_batch(1024, r, 8) for r in range(8)]
    assert covered[0] == (0, 128) and covered[-1] == (896, 1024)
    spans = [sharding.shard_batch(10, r, 4) for r in range(4)]
    assert spans == [(0, 3), (3, 6), (6, 8), (8, 10)]


def _worker(rank, world, port, lengths, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        shards = sharding.partition_reads(lengths, world)
        mine = shards[rank]
        # stand-in for the per-rank basecall: result depends only on the read, not on the rank
        local = [('read{}'.format(i), -float(lengths[i]), [i, i + 1], int(lengths[i])) for i in mine]
        full = sharding.gather_results(mine, local, len(lengths))
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)          # the bench's max-over-ranks timing reduce
        ok = all(r is not None and r[0] == 'read{}'.format(i) for i, r in enumerate(full)) and t.item() == world
        with open(os.path.join(out_dir, 'rank{}'.format(rank)), 'w') as fh:
            fh.write('ok' if ok else 'bad')
    finally:
        dist.destroy_process_group()


def test_gather_over_two_gloo_ranks(tmp_path):
    lengths = [50000, 1200, 90000, 30000, 31000, 700, 64000]
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, lengths, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / 'rank0').read_text() == 'ok' and (tmp_path / 'rank1').read_text() == 'ok'
