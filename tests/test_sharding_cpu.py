"""CPU (gloo, world_size 2): batch sharding and the final gather -- the only multi-rank logic of the path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mod_extraction_b200.sharding import all_gather_rendered, max_over_ranks, shard_range, shard_sizes


def test_shard_ranges_cover_batch():
    for n in (0, 1, 7, 512, 4096, 4097):
        for w in (1, 2, 3, 8):
            ranges = [shard_range(n, r, w) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(w - 1))
            assert max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1
    assert shard_range(4096, 3, 8) == (1536, 2048)        # BASELINE config 4: 512 examples per GPU


def _worker(rank, world, port, n):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
        lo, hi = shard_range(n, rank, world)
        got = all_gather_rendered(full[lo:hi].clone(), n)
        assert torch.equal(got, full), (rank, got)
        assert max_over_ranks(float(rank + 1), "cpu") == float(world)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 7])
def test_all_gather_world2_gloo(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, n), nprocs=2, join=True)


def test_host_step_chunk_schedule_covers_the_batch():
    """render_host's chunk schedule: contiguous, complete, a short first and last chunk for batches of several chunks."""
    from mod_extraction_b200.render import InterwovenRenderer
    for B, chunk in [(1, 512), (26, 7), (23, 6), (512, 512), (1100, 512), (4096, 512), (4096, 256), (4097, 384), (5, 1)]:
        edges = InterwovenRenderer._chunk_edges(B, chunk)
        assert edges[0][0] == 0 and edges[-1][1] == B
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        assert all(0 < hi - lo <= chunk + max(1, chunk // 4) for lo, hi in edges)
        if B > 2 * chunk:
            short = max(1, chunk // 4)
            assert edges[0][1] - edges[0][0] == short and edges[-1][1] - edges[-1][0] <= short
