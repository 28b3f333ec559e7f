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


def test_generate_chunk_plan_covers_the_batch_with_small_edge_chunks():
    """MelGanGenerator.generate(): the first and the last chunk (whose copies cannot overlap any
    kernel) are small, the middle is split evenly; every plan is a partition of [0, B)."""
    from music_synthesis_b200.generator.full import MelGanGenerator as G
    for B, chunk, edge in ((256, 120, 8), (256, 64, 16), (100, 64, 16), (40, 64, 16), (5, 120, 8),
                           (64, 64, 0), (1000, 120, 8), (17, 4, 2)):
        plan = G._chunk_plan(B, chunk, edge)
        assert plan[0][0] == 0 and plan[-1][1] == B
        assert all(a[1] == b[0] for a, b in zip(plan, plan[1:]))
        assert all(0 < hi - lo <= max(chunk, edge) for lo, hi in plan)
    assert G._chunk_plan(256, 120, 8) == [(0, 8), (8, 128), (128, 248), (248, 256)]


def test_phase_split_tap_offsets_of_the_fused_upsampling_stage():
    """The algebra behind csrc/upstack.cu on the host (numpy): with even and odd rows of a tile in
    separate blocks (E[m] = x[2m], O[m] = x[2m+1]) a k3 conv tap of ODD dilation d reads
        E block:  -d -> O[m - (d+1)/2],  0 -> E[m],  +d -> O[m + (d-1)/2]
        O block:  -d -> E[m - (d-1)/2],  0 -> O[m],  +d -> E[m + (d+1)/2]
    and the k4 s2 p1 transposed conv is  out[2u] = x[u] W1 + x[u-1] W3,
    out[2u+1] = x[u+1] W0 + x[u] W2."""
    import numpy as np
    rs = np.random.RandomState(0)
    L = 64
    x = rs.standard_normal(L)
    w = rs.standard_normal(3)
    E, O = x[0::2], x[1::2]
    get = lambda a, i: a[i] if 0 <= i < len(a) else 0.0
    for d in (1, 3, 9):
        ref = np.array([w[0] * get(x, t - d) + w[1] * x[t] + w[2] * get(x, t + d) for t in range(L)])
        ye = np.array([w[0] * get(O, m - (d + 1) // 2) + w[1] * E[m] + w[2] * get(O, m + (d - 1) // 2)
                       for m in range(L // 2)])
        yo = np.array([w[0] * get(E, m - (d - 1) // 2) + w[1] * O[m] + w[2] * get(E, m + (d + 1) // 2)
                       for m in range(L // 2)])
        assert np.allclose(ye, ref[0::2]) and np.allclose(yo, ref[1::2])
    # transposed conv: torch semantics out[t] = sum_q x[q] W[t + 1 - 2q]
    wt = rs.standard_normal(4)
    u = rs.standard_normal(L // 2)
    ref = np.zeros(L)
    for q in range(L // 2):
        for k in range(4):
            t = 2 * q - 1 + k
            if 0 <= t < L:
                ref[t] += u[q] * wt[k]
    even = np.array([get(u, m) * wt[1] + get(u, m - 1) * wt[3] for m in range(L // 2)])
    odd = np.array([get(u, m + 1) * wt[0] + get(u, m) * wt[2] for m in range(L // 2)])
    assert np.allclose(even, ref[0::2]) and np.allclose(odd, ref[1::2])
