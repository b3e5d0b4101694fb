"""CPU, world_size 2 over gloo: the N>1 host logic — block sharding of pairs and the single result
all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nopesac_b200 import dist as nd


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 64, 511, 512):
        for w in (1, 2, 4, 8):
            ranges = [nd.shard_range(n, r, w) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = nd.shard_sizes(n, w)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, num_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(num_pairs * 16, dtype=torch.float32).reshape(num_pairs, 16)   # "the single-rank result"
        lo, hi = nd.shard_range(num_pairs, rank, world)
        got = nd.gather_results(full[lo:hi].clone(), num_pairs)
        q.put((rank, bool(torch.equal(got, full))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_pairs", [8, 7])
def test_two_rank_gather_equals_single_rank_concat(num_pairs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
