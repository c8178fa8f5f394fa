"""Host-side multi-GPU logic on CPU: world_size 2, gloo backend."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rwkv_lm_ext_b200.dist import gather_rows, max_over_ranks, shard_batch, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        idx = torch.arange(n * 5).view(n, 5)
        mine = shard_batch(idx)
        emb = mine.float() * 2 + 1                       # stand-in for the per-rank embedding forward
        full = gather_rows(emb, n)
        assert torch.equal(full, idx.float() * 2 + 1)
        slow = max_over_ranks(10.0 + rank)
        assert slow == 10.0 + world - 1
        if rank == 0:
            torch.save(full, out)
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_shard_and_gather(tmp_path):
    out = str(tmp_path / "full.pt")
    mp.spawn(_worker, args=(2, _free_port(), 7, out), nprocs=2, join=True)
    assert torch.load(out).shape == (7, 5)
