"""world_size-2 gloo test of the N>1 host logic: contiguous row sharding,
global ids, gather + merge of the per-rank top-k lists, data-parallel split of sequences.
The local search of each rank is the oracle here (there is no GPU on this box); on the GPU
box the same ShardedSearch drives css_index_search_device."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, d, k, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from claude_semantic_search_b200.sharded import ShardedSearch, shard_bounds
    from oracle import search_oracle as so
    rng = np.random.default_rng(0)
    x = so.normalize_rows(rng.standard_normal((n, d), dtype=np.float32))
    q = so.normalize_rows(rng.standard_normal((5, d), dtype=np.float32))
    x[n - 3] = x[1]   # a tie across shards: the lower global id must win
    q[0] = x[1]
    lo, hi = shard_bounds(n, world, rank)
    shard = x[lo:hi]
    ss = ShardedSearch(None, id_offset=lo, local_search=lambda qq, kk: so.flat_search(shard, qq, kk))
    D, I = ss.search_host(q, k)
    if rank == 0:
        np.savez(out, D=D, I=I)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_search_two_ranks_gloo(tmp_path):
    import torch.multiprocessing as mp
    from oracle import search_oracle as so
    n, d, k = 1001, 64, 10
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), n, d, k, out), nprocs=2, join=True)
    got = np.load(out)
    rng = np.random.default_rng(0)
    x = so.normalize_rows(rng.standard_normal((n, d), dtype=np.float32))
    q = so.normalize_rows(rng.standard_normal((5, d), dtype=np.float32))
    x[n - 3] = x[1]
    q[0] = x[1]
    D, I = so.flat_search(x, q, k)
    np.testing.assert_array_equal(got["I"], I)
    np.testing.assert_allclose(got["D"], D, atol=1e-6)
    assert list(got["I"][0][:2]) == [1, n - 3]


def test_shard_bounds_cover_everything():
    from claude_semantic_search_b200.sharded import shard_bounds
    for n in (0, 1, 7, 8, 100, 1_000_003):
        for w in (1, 2, 3, 8):
            parts = [shard_bounds(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_merge_topk_host_holes_and_ties():
    from claude_semantic_search_b200.sharded import merge_topk_host
    D = np.array([[[0.9, 0.5, -3.4e38]], [[0.9, 0.7, 0.1]]], np.float32)
    I = np.array([[[4, 9, -1]], [[2, 11, 12]]], np.int64)
    Dm, Im = merge_topk_host(D, I, 4)
    assert Im.tolist() == [[2, 4, 11, 9]]
    np.testing.assert_allclose(Dm, [[0.9, 0.9, 0.7, 0.5]])
