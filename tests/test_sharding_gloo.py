"""CPU: the N>1 host logic (edit sharding + the single result gather) with a world-size-2 gloo group."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusionhandles_b200.batch import gather_records, shard_edits


def test_shard_edits_partition():
    for n, w in ((256, 1), (256, 2), (256, 8), (10, 4), (3, 8)):
        shards = [shard_edits(n, r, w) for r in range(w)]
        assert sorted(e for s in shards for e in s) == list(range(n))
        assert all(e % w == r for r, s in enumerate(shards) for e in s)
    with pytest.raises(ValueError):
        shard_edits(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 12
    mine = shard_edits(n, rank, world)
    rec = torch.tensor([[e, e * e] for e in mine], dtype=torch.int64)      # per-edit record computed by the owner
    out = gather_records(rec, dst=0)
    if rank == 0:
        q.put(out.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_gather_records_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out == [[e, e * e] for e in range(12)]
