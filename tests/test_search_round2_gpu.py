"""Round-2 GPU parity of the flat index: the rigorous bf16 bound of the two-phase scan (adversarial rounding),
clustered corpora, orphan-aware reference mode, batched row kills, k > 128, growth without copies, two host
threads on two streams, and the multi-shard index (block-cyclic ids + in-kernel result exchange) -- the
latter on ONE GPU by listing the device several times, on distinct GPUs when the box has them.
Every call goes through the C ABI; the checker is the oracle (oracle/flat_ip.c)."""
import os
import threading

import numpy as np
import pytest

from oracle import search_oracle as so

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def native():
    from claude_semantic_search_b200 import _native
    assert _native.device_count() >= 1, "no sm_100 device: the GPU suite must not pass silently"
    return _native


def _check(D_ref, I_ref, D, I, tol=TOL):
    ok, why = so.compare_topk(D_ref, I_ref, D, I, tol=tol)
    assert ok, why


def test_two_phase_bound_survives_adversarial_rounding(native):
    """round-to-nearest with bf16's 8-bit significand moves a value by up to 2^-8 relative, twice what
    the round-1 bound (2^-9) allowed.  Twelve rows whose every component sits just BELOW a rounding midpoint lose
    2^-8 of their score in the shadow copy; forty decoys whose components sit just ABOVE midpoints gain 2^-8.  In
    fp32 the twelve win by 7e-4; in bf16 they trail by 6e-3 -- outside 2 x the old bound, inside the tracked
    ||x - bf16(x)|| bound.  They must come back, from the batch-1 and from the batched path."""
    d, n = 768, 120_000
    rng = np.random.default_rng(5)
    x = so.normalize_rows(rng.standard_normal((n, d), dtype=np.float32))
    q = np.ones((1, d), np.float32) / np.float32(np.sqrt(d))
    u8 = np.float32(2.0 ** -8)
    delta = np.float32(0.02) * u8
    s5 = np.float32(2.0 ** -5)
    adv = np.full((12, d), s5 * (np.float32(1) + 3 * u8 - delta), np.float32)       # rounds DOWN to 1 + 2 u8
    decoys = np.full((40, d), s5 * (np.float32(1) + 3 * u8 + delta), np.float32)    # rounds UP to 1 + 4 u8
    for j in range(40):                                                             # 10 %: 1 + u8 + delta -> 1 + 2 u8
        decoys[j, rng.choice(d, size=77, replace=False)] = s5 * (np.float32(1) + u8 + delta)
    where_adv = rng.choice(n, size=12, replace=False)
    where_dec = rng.choice(np.setdiff1d(np.arange(n), where_adv), size=40, replace=False)
    x[where_adv] = adv
    x[where_dec] = decoys
    s32 = x[np.concatenate([where_adv, where_dec])].astype(np.float64) @ q[0].astype(np.float64)
    assert s32[:12].min() - s32[12:].max() > 5e-4                   # fp32: the twelve are the top-12
    import torch
    xb = torch.from_numpy(x[np.concatenate([where_adv, where_dec])]).to(torch.bfloat16).to(torch.float64).numpy()
    sb = xb @ q[0].astype(np.float64)
    assert sb[12:].min() - sb[:12].max() > 2 * 1.1 / 512            # bf16: they trail by more than twice the old bound
    idx = native.Index(d)
    idx.add(x, normalize=False)
    st = idx.scan_stats()
    assert 0.0033 < st["max_bf16_error_norm"] < 0.0036, st          # = 2^-8 ||x|| of the planted rows, tracked at add
    for k in (1, 10, 12, 32):
        D, I = idx.search(q, k)
        Dr, Ir = so.flat_search_c(x, q, k)
        _check(Dr, Ir, D, I, tol=1e-5)
    D, I = idx.search(q, 12)
    assert set(I[0].tolist()) == set(where_adv.tolist())            # all twelve, ahead of every decoy
    qb = np.repeat(q, 16, axis=0)                                   # tensor-core path: the query is rounded as well
    Db, Ib = idx.search(qb, 12)
    np.testing.assert_array_equal(Ib[3], I[0])
    np.testing.assert_array_equal(Db[3], D[0])
    idx.close()


@pytest.mark.parametrize("order", ["session", "shuffled"])
def test_clustered_corpus_exact_and_mostly_proven(native, order):
    """a clustered corpus (2000 caps, members adjacent in row order for order='session').  The
    result is exact either way; the two-phase scan must also PROVE nearly all of it (fallback rate < 5 % here,
    measured at 1 M rows in bench.py)."""
    from bench_data import clustered_numpy
    n = 300_000
    x, q, _ = clustered_numpy(n, n_clusters=600, order=order, seed=3, n_queries=48)
    idx = native.Index(768)
    idx.add(x, normalize=False)
    s0 = idx.scan_stats()
    for k in (10, 32):
        Dr, Ir = so.flat_search_c(x, q, k)
        for i in range(0, 48, 8):                                # nq = 8 per call: the streaming scan
            D, I = idx.search(q[i:i + 8], k)
            _check(Dr[i:i + 8], Ir[i:i + 8], D, I)
    s1 = idx.scan_stats()
    asked = s1["two_phase_queries"] - s0["two_phase_queries"]
    unproven = s1["unproven_queries"] - s0["unproven_queries"]
    assert asked == 96, s1
    assert unproven <= 4, f"{unproven} of {asked} clustered queries fell back to the fp32 sweep"
    # the batched path on the same data
    Db, Ib = idx.search(q, 10)
    Dr, Ir = so.flat_search_c(x, q, 10)
    _check(Dr, Ir, Db, Ib)
    idx.close()


def test_near_duplicate_corpus_switches_to_fp32_sweep(native):
    """A corpus whose rows all sit within the bf16 bound of each other (random-init encoder outputs: pairwise cosine
    ~0.99) cannot be proven from int8 or bf16 scores; results stay exact and after 64 such queries per tier the index
    stops paying for phase 1 (adaptive bypass: int8 tier off after 64 queries, bf16 tier after the next 64)."""
    rng = np.random.default_rng(9)
    n, d = 80_000, 768
    base = so.normalize_rows(rng.standard_normal((1, d), dtype=np.float32))
    x = so.normalize_rows(base + 0.02 * rng.standard_normal((n, d), dtype=np.float32) / np.sqrt(d) * 3)
    q = so.normalize_rows(base + 0.02 * rng.standard_normal((192, d), dtype=np.float32) / np.sqrt(d) * 3)
    idx = native.Index(d)
    idx.add(x, normalize=False)
    Dr, Ir = so.flat_search_c(x, q, 10)
    for i in range(192):
        D, I = idx.search(q[i:i + 1], 10)
        _check(Dr[i:i + 1], Ir[i:i + 1], D, I, tol=1e-5)
    st = idx.scan_stats()
    if os.environ.get("CSS_SCAN_BF16", "1") != "0" and os.environ.get("CSS_SCAN_ADAPTIVE", "1") != "0":
        assert st["unproven_queries"] >= 60 and st["bypassed"], st
        assert st["two_phase_queries"] < 192 and st["last_tier"] == 0, st   # the tail of the loop skipped phase 1
    idx.close()


def test_reference_mode_window_includes_orphans(native, tmp_path):
    """filter_mode='reference' reproduces R = (global top-max_results INCLUDING orphaned rows) walked with
    the orphan / filter checks (src/storage.py:432-490).  After deletions the orphans must still occupy their slots
    of the window -- css_filter.ignore_alive must reach the scan."""
    from claude_semantic_search_b200 import Chunk, HybridStorage, SearchConfig, StorageConfig
    rng = np.random.default_rng(2)
    n, d = 3000, 768
    emb = rng.standard_normal((n, d)).astype(np.float32)
    qv = rng.standard_normal(d).astype(np.float32)
    chunks = [Chunk(id=f"c{i}", text=f"t{i}", metadata=dict(session_id=f"s{i % 7}", project_name=f"/p/{i % 3}",
                    file_path=f"/f/{i % 50}", chunk_type="qa", timestamp="2024-01-01T00:00:00+00:00",
                    has_code=bool(i % 2), has_tools=False, message_count=1, char_count=5, word_count=1),
                    embedding=emb[i]) for i in range(n)]
    st = HybridStorage(StorageConfig(data_dir=str(tmp_path), use_gpu=True, auto_save=False, filter_mode="reference"))
    st.initialize()
    st.add_chunks(chunks)
    x = so.normalize_rows(emb)
    qn = so.normalize_query(qv)
    order = np.lexsort((np.arange(n), -(x @ qn[0]).astype(np.float64)))
    # orphan 30 of the global top-40 (delete through the API: alive bits cleared)
    dead = set(int(i) for i in order[:40][rng.permutation(40)[:30]])
    for i in sorted(dead):
        assert st.delete_chunk(f"c{i}")
    cfg = SearchConfig(top_k=10, max_results=40)
    got = [r.chunk_id for r in st.search(qv, cfg, filters={"has_code": True})]
    # the reference: walk the top-40 INCLUDING orphans, skip orphans, apply the filter, stop at top_k
    want = [f"c{int(i)}" for i in order[:40] if int(i) not in dead and (int(i) % 2 == 1)][:10]
    assert got == want, (got, want)
    assert len(want) < 10                                       # the truncation is what the mode is for
    # prefilter mode returns the full best-10 of the surviving, matching rows; R is its prefix
    st.config.filter_mode = "prefilter"
    full = [r.chunk_id for r in st.search(qv, cfg, filters={"has_code": True})]
    assert len(full) == 10 and full[:len(want)] == want
    st.close()


def test_set_alive_ids_and_large_k(native):
    rng = np.random.default_rng(4)
    n, d = 20_000, 768
    x = so.normalize_rows(rng.standard_normal((n, d), dtype=np.float32))
    q = so.normalize_rows(rng.standard_normal((2, d), dtype=np.float32))
    idx = native.Index(d)
    idx.add(x)
    dead = rng.choice(n, size=700, replace=False)
    idx.set_alive_ids(np.concatenate([dead, [-5, n + 3]]), False)        # out-of-range ids are ignored
    alive = np.ones(n, bool)
    alive[dead] = False
    words, n_pass = idx.filter_mask(None)
    np.testing.assert_array_equal(words, so.pack_mask(alive))
    assert n_pass == n - 700
    idx.set_alive_ids(dead[:100], True)
    alive[dead[:100]] = True
    # k > CSS_MAX_K (the reference's max_results is a free field): exact, in order, holes at the end
    for k in (129, 300, 1000):
        D, I = idx.search(q, k)
        Dr, Ir = so.flat_search_c(x, q, k, mask_words=so.pack_mask(alive))
        _check(Dr, Ir, D, I)
    small = native.Index(d)
    small.add(x[:150])
    D, I = small.search(q[:1], 200)
    assert (I[0, :150] >= 0).all() and (I[0, 150:] == -1).all() and len(set(I[0, :150].tolist())) == 150
    small.close()
    idx.close()


def test_growth_maps_memory_without_moving_rows(native):
    """Appending past the capacity maps more physical memory behind the same addresses (CUDA VMM): earlier rows
    are untouched, ids stay dense, searches see all rows."""
    rng = np.random.default_rng(8)
    d = 768
    idx = native.Index(d)
    parts = [so.normalize_rows(rng.standard_normal((m, d), dtype=np.float32)) for m in (1, 900, 5000, 33, 70_000, 7)]
    first = 0
    for p in parts:
        assert idx.add(p) == first
        first += p.shape[0]
    x = np.concatenate(parts)
    assert idx.ntotal == x.shape[0] and idx.capacity >= idx.ntotal
    np.testing.assert_array_equal(idx.get_rows(0, 6000), x[:6000])
    q = so.normalize_rows(rng.standard_normal((3, d), dtype=np.float32))
    D, I = idx.search(q, 10)
    Dr, Ir = so.flat_search_c(x, q, 10)
    _check(Dr, Ir, D, I)
    idx.reset()
    assert idx.ntotal == 0
    idx.add(parts[2])
    D, I = idx.search(q, 5)
    Dr, Ir = so.flat_search_c(parts[2], q, 5)
    _check(Dr, Ir, D, I)
    idx.close()


def test_two_threads_two_streams(native):
    """css_index_search_device from two host threads on two streams -- per-stream scratch, no
    shared partial lists / tickets / overflow counters."""
    import torch
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(12)
    n, d = 200_000, 768
    x = so.normalize_rows(rng.standard_normal((n, d), dtype=np.float32))
    q = so.normalize_rows(rng.standard_normal((64, d), dtype=np.float32))
    idx = native.Index(d)
    idx.add(x)
    Dr, Ir = so.flat_search_c(x, q, 10)
    qd = torch.from_numpy(q).to(dev)
    out = {}

    def worker(tid):
        s = torch.cuda.Stream(dev)
        D = torch.empty((64, 10), device=dev, dtype=torch.float32)
        I = torch.empty((64, 10), device=dev, dtype=torch.int64)
        with torch.cuda.stream(s):
            for rep in range(6):
                for i in range(tid, 64, 2):          # batch-1 two-phase scans, interleaved with the other thread
                    idx.search_device(qd[i].data_ptr(), 1, 10, D[i].data_ptr(), I[i].data_ptr(), 0, 0, s.cuda_stream)
                lo = tid * 32                        # and a tensor-core batch of its own half
                idx.search_device(qd[lo].data_ptr(), 32, 10, D[lo].data_ptr(), I[lo].data_ptr(), 0, 0, s.cuda_stream)
                for i in range(1 - tid, 64, 2):
                    if lo <= i < lo + 32:
                        continue
                    idx.search_device(qd[i].data_ptr(), 1, 10, D[i].data_ptr(), I[i].data_ptr(), 0, 0, s.cuda_stream)
        s.synchronize()
        out[tid] = (D.cpu().numpy(), I.cpu().numpy())

    ts = [threading.Thread(target=worker, args=(t,)) for t in (0, 1)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for tid in (0, 1):
        _check(Dr, Ir, out[tid][0], out[tid][1])
    idx.close()


def _devices_for_shards(native, n_shards):
    have = native.device_count()
    return [i % have for i in range(n_shards)]


@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_multi_shard_index_matches_single(native, n_shards, tmp_path):
    """css_index_create_sharded: block-cyclic rows over n shards, in-kernel exchange of the top-k lists.  Runs on one
    GPU (the device listed n times) as well as on n distinct GPUs; ids, scores, filters, alive bits, persistence
    and compaction must match the oracle exactly as the single-device index does."""
    devs = _devices_for_shards(native, n_shards)
    rng = np.random.default_rng(20 + n_shards)
    d, n = 768, 9 * 4096 * n_shards // 2 + 1234           # several blocks per shard + a ragged tail
    x = so.normalize_rows(rng.standard_normal((n, d), dtype=np.float32))
    q = so.normalize_rows(rng.standard_normal((40, d), dtype=np.float32))
    x[n - 2] = x[5]
    q[0] = x[5]                                            # exact duplicates on different shards: ids decide
    idx = native.Index(d, devices=devs)
    assert idx.add(x[:10_000]) == 0
    assert idx.add(x[10_000:10_001]) == 10_000
    assert idx.add(x[10_001:]) == 10_001
    assert idx.ntotal == n
    np.testing.assert_array_equal(idx.get_rows(4090, 8200), x[4090:4090 + 8200])
    Dr, Ir = so.flat_search_c(x, q, 10)
    for i in range(6):                                     # batch-1: two-phase scan + exchange
        D, I = idx.search(q[i:i + 1], 10)
        _check(Dr[i:i + 1], Ir[i:i + 1], D, I)
    assert idx.search(q[:1], 10)[1][0][:2].tolist() == [5, n - 2]
    D, I = idx.search(q[:9], 10)                           # nq = 9: one grid, nine exchanges
    _check(Dr[:9], Ir[:9], D, I)
    D, I = idx.search(q, 10)                               # tensor-core batches + host merge (shards < 65536 rows: scan)
    _check(Dr, Ir, D, I)
    D, I = idx.search(q[:3], 100)                          # k = 100: the fp32 sweep + exchange
    Dr100, Ir100 = so.flat_search_c(x, q[:3], 100)
    _check(Dr100, Ir100, D, I)
    # filters: clause on a column + explicit row mask + dead rows
    col = rng.integers(0, 50, size=n).astype(np.int32)
    idx.set_column(3, col)
    dead = rng.choice(n, size=500, replace=False)
    idx.set_alive_ids(dead, False)
    alive = np.ones(n, bool)
    alive[dead] = False
    rowmask = rng.random(n) < 0.6
    flt = native.Filter().add_range(3, 10, 29).set_row_mask(so.pack_mask(rowmask))
    want = alive & rowmask & (col >= 10) & (col <= 29)
    words, n_pass = idx.filter_mask(flt)
    np.testing.assert_array_equal(words, so.pack_mask(want))
    assert n_pass == int(want.sum())
    D, I = idx.search(q[:5], 10, flt)
    Drm, Irm = so.flat_search_c(x, q[:5], 10, mask_words=so.pack_mask(want))
    _check(Drm, Irm, D, I)
    # persistence: the file is a plain faiss IndexFlatIP file whatever the number of shards
    path = tmp_path / "sharded.faiss"
    idx.save(path)
    single = native.Index(d)
    single.load(path)
    assert single.ntotal == n
    np.testing.assert_array_equal(single.get_rows(0, n), x)
    single.close()
    again = native.Index(d, devices=devs)
    again.load(path)
    D, I = again.search(q[:4], 10)
    _check(Dr[:4], Ir[:4], D, I)
    # compaction across shards
    keep = np.sort(rng.choice(n, size=n // 3, replace=False))
    again.compact(keep)
    assert again.ntotal == keep.shape[0]
    np.testing.assert_array_equal(again.get_rows(0, keep.shape[0]), x[keep])
    D, I = again.search(q[:4], 10)
    Drk, Irk = so.flat_search_c(x[keep], q[:4], 10)
    _check(Drk, Irk, D, I)
    again.close()
    idx.close()


def test_hybrid_storage_over_several_shards(native, tmp_path):
    """The drop-in surface over a multi-device index: StorageConfig.devices, nothing else changes."""
    from claude_semantic_search_b200 import Chunk, HybridStorage, SearchConfig, StorageConfig
    devs = _devices_for_shards(native, 4)
    rng = np.random.default_rng(31)
    n, d = 9000, 768
    emb = rng.standard_normal((n, d)).astype(np.float32)
    chunks = [Chunk(id=f"c{i}", text=f"t{i}", metadata=dict(session_id=f"s{i % 7}", project_name=f"/home/u/proj{i % 4}",
                    file_path=f"/f/{i % 50}", chunk_type="qa", timestamp=f"2024-02-{1 + i % 28:02d}T00:00:00+00:00",
                    has_code=bool(i % 2), has_tools=False, message_count=1, char_count=5, word_count=1),
                    embedding=emb[i]) for i in range(n)]
    st = HybridStorage(StorageConfig(data_dir=str(tmp_path), use_gpu=True, devices=devs))
    st.initialize()
    st.add_chunks(chunks[:5000])
    st.add_chunks(chunks[5000:])
    assert st.faiss_index._native.ntotal == n
    x = so.normalize_rows(emb)
    for qi in (17, 4242, 8999):
        qv = emb[qi] + 0.1 * rng.standard_normal(d).astype(np.float32)
        res = st.search(qv, SearchConfig(top_k=10), filters={"project_name": "PROJ2", "has_code": False})
        s = x @ so.normalize_query(qv)[0]
        ok = np.array([(i % 4 == 2) and (i % 2 == 0) for i in range(n)])
        order = [int(i) for i in np.lexsort((np.arange(n), -s.astype(np.float64))) if ok[i]][:10]
        assert [r.chunk_id for r in res] == [f"c{i}" for i in order]
        assert abs(res[0].similarity - float(s[order[0]])) < 1e-4
    assert st.remove_chunks_for_file("/f/7") == n // 50
    res = st.search(emb[7], SearchConfig(top_k=3))
    assert all(int(r.chunk_id[1:]) % 50 != 7 for r in res)
    st.close()
    # reopen: the index file written by the sharded handle loads back (O(N) load, then O(1) initialize)
    st2 = HybridStorage(StorageConfig(data_dir=str(tmp_path), use_gpu=True, devices=devs))
    st2.initialize()
    assert st2.faiss_index._native.ntotal == n
    res2 = st2.search(emb[8], SearchConfig(top_k=1))
    assert res2[0].chunk_id == "c8"
    st2.close()


IPC_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["CSS_ROOT"])
from claude_semantic_search_b200 import _native
from claude_semantic_search_b200.sharded import ShardedSearch, shard_bounds
from oracle import search_oracle as so
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ndev = torch.cuda.device_count()
dev = torch.device("cuda", rank % ndev)
torch.cuda.set_device(dev)
dist.init_process_group("gloo")                 # plumbing only: the result exchange is CUDA IPC + NVLink / local memory
torch.cuda.set_stream(torch.cuda.Stream(dev))
n, d, k = 90_001, 768, 10
rng = np.random.default_rng(0)
x = so.normalize_rows(rng.standard_normal((n, d), dtype=np.float32))
q = so.normalize_rows(rng.standard_normal((12, d), dtype=np.float32))
x[n - 3] = x[1]; q[0] = x[1]
lo, hi = shard_bounds(n, world, rank)
idx = _native.Index(d, device=dev.index)
idx.add(x[lo:hi])
ss = ShardedSearch(idx, id_offset=lo)
qd = torch.from_numpy(q).to(dev)
Dr, Ir = so.flat_search_c(x, q, k)
for rep in range(3):                             # epochs advance in lockstep on every rank
    for i in range(4):
        D1, I1 = ss.search_device(qd[i:i + 1], k)
        torch.cuda.synchronize(dev)
        ok, why = so.compare_topk(Dr[i:i + 1], Ir[i:i + 1], D1.cpu().numpy(), I1.cpu().numpy())
        assert ok, f"rank {rank} q{i}: {why}"
D8, I8 = ss.search_device(qd[:8], k)             # eight queries, one launch, eight exchanges
torch.cuda.synchronize(dev)
ok, why = so.compare_topk(Dr[:8], Ir[:8], D8.cpu().numpy(), I8.cpu().numpy())
assert ok, f"rank {rank} nq8: {why}"
assert I8.cpu().numpy()[0][:2].tolist() == [1, n - 3]
# a filtered search at N > 1: every rank evaluates the filter on its own shard
col = (np.arange(n) % 10).astype(np.int32)
idx.set_column(2, col[lo:hi])
flt = _native.Filter().add_range(2, 3, 5)
mptr, _ = idx.filter_mask_device(flt, torch.cuda.current_stream(dev).cuda_stream)
Df, If = ss.search_device(qd[:4], k, mask_ptr=mptr)
torch.cuda.synchronize(dev)
want = (col >= 3) & (col <= 5)
Drf, Irf = so.flat_search_c(x, q[:4], k, mask_words=so.pack_mask(want))
ok, why = so.compare_topk(Drf, Irf, Df.cpu().numpy(), If.cpu().numpy())
assert ok, f"rank {rank} filtered: {why}"
# one query in HOST memory per call (css_index_search_exchange: result through mapped memory + completion flag)
torch.cuda.synchronize(dev)
for i in range(4):
    Dh, Ih = idx.search_exchange_host(ss._exchange(), q[i], k, lo)
    ok, why = so.compare_topk(Dr[i:i + 1], Ir[i:i + 1], Dh, Ih)
    assert ok, f"rank {rank} host q{i}: {why}"
Dh, Ih = idx.search_exchange_host(ss._exchange(), q[1], k, lo, mask_ptr=mptr)
ok, why = so.compare_topk(Drf[1:2], Irf[1:2], Dh, Ih)
assert ok, f"rank {rank} host filtered: {why}"
dist.barrier()
ss.close()
idx.close()
dist.destroy_process_group()
print(f"rank {rank} ok")
"""


@pytest.mark.parametrize("world", [2, 4])
def test_in_kernel_exchange_between_processes(native, tmp_path, world):
    """The torchrun layout (one process per shard) with the fused exchange: CUDA IPC mappings between the ranks, lists
    stored into the peers' memory by the scan kernel, merged in the kernel.  With fewer GPUs than ranks the ranks
    share a device (the protocol is the same; only the wire differs)."""
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    script = tmp_path / "ipc_worker.py"
    script.write_text(IPC_WORKER)
    env = dict(os.environ, CSS_ROOT=str(root))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29540 + world), str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for rk in range(world):
        assert f"rank {rk} ok" in r.stdout


def test_repeated_filter_reuses_mask_until_the_index_changes(native):
    """An unchanged filter over an unchanged index is not evaluated again (its mask is still on the device); any
    change of rows, columns or alive bits invalidates it."""
    rng = np.random.default_rng(41)
    n, d = 50_000, 768
    x = so.normalize_rows(rng.standard_normal((n, d), dtype=np.float32))
    q = so.normalize_rows(rng.standard_normal((3, d), dtype=np.float32))
    col = rng.integers(0, 20, size=n).astype(np.int32)
    idx = native.Index(d)
    idx.add(x)
    idx.set_column(1, col)
    launches = native.kernel_launch_count

    def filt(lo, hi, allowed):
        return native.Filter().add_range(1, lo, hi).add_set(1, allowed, 20)

    def expect(lo, hi, allowed, alive=None):
        m = (col >= lo) & (col <= hi) & np.isin(col, allowed)
        return m if alive is None else (m & alive)

    f1 = filt(3, 12, [3, 4, 5, 9, 12, 15])
    l0 = launches()
    D1, I1 = idx.search(q[:1], 10, f1)
    first = launches() - l0
    l0 = launches()
    D2, I2 = idx.search(q[:1], 10, filt(3, 12, [3, 4, 5, 9, 12, 15]))     # an equal filter, rebuilt
    again = launches() - l0
    assert again == first - 1, (first, again)                               # the filter kernel was skipped
    np.testing.assert_array_equal(I1, I2)
    Dr, Ir = so.flat_search_c(x, q[:1], 10, mask_words=so.pack_mask(expect(3, 12, [3, 4, 5, 9, 12, 15])))
    _check(Dr, Ir, D2, I2)
    D3, I3 = idx.search(q[:1], 10, filt(3, 12, [3, 4, 5, 9, 12]))          # one bit of the set differs: evaluated
    Dr, Ir = so.flat_search_c(x, q[:1], 10, mask_words=so.pack_mask(expect(3, 12, [3, 4, 5, 9, 12])))
    _check(Dr, Ir, D3, I3)
    # index changes: column rewrite, kill, append -- each invalidates the cached mask
    col[:] = (col + 7) % 20
    idx.set_column(1, col)
    D4, I4 = idx.search(q[:1], 10, filt(3, 12, [3, 4, 5, 9, 12]))
    Dr, Ir = so.flat_search_c(x, q[:1], 10, mask_words=so.pack_mask(expect(3, 12, [3, 4, 5, 9, 12])))
    _check(Dr, Ir, D4, I4)
    alive = np.ones(n, bool)
    alive[I4[0][:5]] = False
    idx.set_alive_ids(I4[0][:5], False)
    D5, I5 = idx.search(q[:1], 10, filt(3, 12, [3, 4, 5, 9, 12]))
    Dr, Ir = so.flat_search_c(x, q[:1], 10, mask_words=so.pack_mask(expect(3, 12, [3, 4, 5, 9, 12], alive)))
    _check(Dr, Ir, D5, I5)
    extra = so.normalize_rows(q[:1] + 0.01 * rng.standard_normal((4, d), dtype=np.float32))
    idx.add(extra)
    idx.set_column(1, np.full(4, 4, np.int32), start=n)
    D6, I6 = idx.search(q[:1], 10, filt(3, 12, [3, 4, 5, 9, 12]))
    assert set(I6[0][:4].tolist()) == {n, n + 1, n + 2, n + 3}
    idx.close()
