"""-m "not gpu": the N>1 path on CPU -- clip sharding is a partition, and a world_size-2 gloo
run reassembles per-rank results in clip order (inference needs no other communication)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from music_synthesis_b200.sharding import clip_shard, gather_clips


@pytest.mark.parametrize("n,w", [(256, 8), (256, 1), (5, 2), (3, 8), (0, 4), (257, 4)])
def test_clip_shard_is_a_balanced_partition(n, w):
    spans = [clip_shard(n, r, w) for r in range(w)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and a <= b
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        clip_shard(4, 4, 4)


def _worker(rank, world, port, n):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = clip_shard(n, rank, world)
        # stand-in for the per-clip waveform: a deterministic function of the clip index, so
        # every rank can check the reassembled batch without sharing inputs
        local = torch.stack([torch.full((1, 6), float(i)) + torch.arange(6.0) for i in range(lo, hi)]) \
            if hi > lo else torch.zeros((0, 1, 6))
        full = gather_clips(local, n)
        ref = torch.stack([torch.full((1, 6), float(i)) + torch.arange(6.0) for i in range(n)])
        assert full.shape == ref.shape and torch.equal(full, ref)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [5, 8])
def test_two_rank_gloo_gather_preserves_clip_order(n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n), nprocs=2, join=True)


def _grad_sync_worker(rank, world, port):
    """data-parallel gradient exchange of the trainers (train/train.py:_sync_grads): sum over
    ranks then 1/world -- rank r holds gradient r+1 everywhere, every rank must end with 1.5"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from music_synthesis_b200.train.train import _sync_grads
        net = torch.nn.Sequential(torch.nn.Conv1d(2, 3, 3), torch.nn.Conv1d(3, 1, 1))
        for p in net.parameters():
            p.grad = torch.full_like(p, float(rank + 1))
        _sync_grads(torch.optim.SGD(net.parameters(), lr=0.1), net)
        for p in net.parameters():
            assert torch.allclose(p.grad, torch.full_like(p, 1.5))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gradient_average():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_grad_sync_worker, args=(2, port), nprocs=2, join=True)
