"""Host-side logic of the N > 1 path on CPU: unit sharding and the out-of-band exchange of the NCCL unique id over a
world_size-2 gloo group (the same code path torchrun drives on the GPU box; only the backend differs)."""
import os
import socket

import numpy as np
import pytest

from thylacine_b200.distributed import exchange_unique_id, rank_world, shard_range


@pytest.mark.parametrize("total,world", [(0, 1), (1, 2), (7, 2), (4096, 8), (65536, 8), (1_000_003, 8), (5, 8)])
def test_shard_range_partitions_every_unit_once(total, world):
    blocks = [shard_range(total, r, world) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == total
    for (lo0, hi0), (lo1, hi1) in zip(blocks, blocks[1:]):
        assert hi0 == lo1 and lo0 <= hi0
    sizes = [hi - lo for lo, hi in blocks]
    assert max(sizes) - min(sizes) <= 1 and sum(sizes) == total


def test_shard_range_rejects_bad_arguments():
    for args in [(10, 2, 2), (10, -1, 2), (10, 0, 0), (-1, 0, 1)]:
        with pytest.raises(ValueError):
            shard_range(*args)


def test_rank_world_defaults(monkeypatch):
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    assert rank_world() == (0, 1, 0)
    monkeypatch.setenv("RANK", "3")
    monkeypatch.setenv("WORLD_SIZE", "8")
    monkeypatch.setenv("LOCAL_RANK", "3")
    assert rank_world() == (3, 8, 3)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        calls = []

        def make_id():
            calls.append(rank)
            return bytes(range(128))

        uid = exchange_unique_id(make_id, rank, world)
        assert uid == bytes(range(128)) and calls == ([0] if rank == 0 else [])
        # the bench's aggregation: every rank times its own shard, the job's time is the max over ranks,
        # throughput = units of ALL ranks / that time
        lo, hi = shard_range(1001, rank, world)
        local_ms = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)
        dist.all_reduce(local_ms, op=dist.ReduceOp.MAX)
        units = torch.tensor([float(hi - lo)], dtype=torch.float64)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
        # moments merge: (count, sum x, sum x x^T) add across ranks
        x = np.random.default_rng(rank).standard_normal((50 + rank, 3))
        acc = torch.from_numpy(np.concatenate([[x.shape[0]], x.sum(0), (x.T @ x).ravel()]))
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        np.save(os.path.join(out_dir, f"r{rank}.npy"), np.concatenate([local_ms.numpy(), units.numpy(), acc.numpy()]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_exchange_and_aggregation(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(r0, r1)
    assert r0[0] == 15.0 and r0[1] == 1001.0
    xs = np.concatenate([np.random.default_rng(r).standard_normal((50 + r, 3)) for r in range(2)])
    assert r0[2] == xs.shape[0]
    assert np.allclose(r0[3:6], xs.sum(0)) and np.allclose(r0[6:], (xs.T @ xs).ravel())
