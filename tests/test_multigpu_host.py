"""Host-side logic of the multi-GPU path on CPU: world_size-2 (and 3) gloo runs of the sharding +
final gather (treatppmx_b200/multigpu.py).  The data path itself has no collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from treatppmx_b200 import multigpu


def test_shard_range_partitions_everything():
    for n in (1, 2, 7, 8, 100, 1024):
        for world in (1, 2, 3, 4, 8):
            blocks = [multigpu.shard_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_shard_chains_follow_their_dataset():
    cd = np.repeat(np.arange(10), 4)
    seen = []
    for r in range(3):
        lo, hi, idx, local = multigpu.shard_chains(cd, 3, r, 10)
        assert ((cd[idx] >= lo) & (cd[idx] < hi)).all() and (local == cd[idx] - lo).all()
        seen += idx.tolist()
    assert sorted(seen) == list(range(40))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = multigpu.shard_range(n_total, world, rank)
    # per-dataset "summaries" whose content encodes the dataset index
    ids = torch.arange(lo, hi, dtype=torch.float64)
    local = {"psm": ids[:, None, None, None] * torch.ones(1, 2, 3, 3, dtype=torch.float64),
             "best_arm": torch.arange(lo, hi, dtype=torch.int32)[:, None].repeat(1, 5)}
    out = multigpu.gather_summaries(local, n_total)
    ok = (out["psm"].shape == (n_total, 2, 3, 3) and
          torch.equal(out["psm"][:, 0, 0, 0], torch.arange(n_total, dtype=torch.float64)) and
          torch.equal(out["best_arm"][:, 0], torch.arange(n_total, dtype=torch.int32)))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 7), (2, 8), (3, 4)])
def test_gather_summaries_gloo(world, n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]
