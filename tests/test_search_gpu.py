"""GPU parity of the flat index (half B of the hot path) against the oracle.

Every call goes through the C ABI (ctypes -> libcss_b200.so).  Bars:
  * scores within 1e-4 absolute of the fp32 oracle (north_star);
  * ids identical except across score gaps below that tolerance;
  * filter masks bit-exact.
"""
import os
import struct

import numpy as np
import pytest

from oracle import search_oracle as so
from oracle.make_golden import build_inputs

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-4  # north_star: scores within 1e-4 absolute


@pytest.fixture(scope="module")
def native():
    from claude_semantic_search_b200 import _native
    assert _native.device_count() >= 1, "no sm_100 device: the GPU suite must not pass silently"
    return _native


def _check(D_ref, I_ref, D, I, tol=TOL):
    ok, why = so.compare_topk(D_ref, I_ref, D, I, tol=tol)
    assert ok, why


# ------------------------------------------------------------------ golden
def test_golden_small_ip_l2_masked(native):
    g = np.load(os.path.join(GOLDEN, "search_small.npz"))
    k = int(g["k"])
    idx = native.Index(64, native.METRIC_INNER_PRODUCT)
    idx.add(g["x"])
    D, I = idx.search(g["q"], k)
    np.testing.assert_array_equal(I, g["I"])
    np.testing.assert_allclose(D, g["D"], atol=1e-5)
    flt = native.Filter().set_row_mask(so.pack_mask(g["mask"]))
    Dm, Im = idx.search(g["q"], k, flt)
    np.testing.assert_array_equal(Im, g["I_masked"])
    np.testing.assert_allclose(Dm, g["D_masked"], atol=1e-5)
    idx.close()
    l2 = native.Index(64, native.METRIC_L2)
    l2.add(g["x"])
    Dl, Il = l2.search(g["q"], k)
    _check(g["D_l2"], g["I_l2"], Dl, Il, tol=1e-5)
    l2.close()


def test_golden_768(native):
    g = np.load(os.path.join(GOLDEN, "search_768.npz"))
    x, q, mask = build_inputs(int(g["seed"]), int(g["n"]), int(g["d"]), int(g["nq"]))
    idx = native.Index(768)
    idx.add(x)
    D, I = idx.search(q, 10)
    _check(g["D"], g["I"], D, I)
    # planted exact duplicates: ties come out in ascending-id order
    assert list(I[0][:3]) == [7, 100, int(g["n"]) - 148]
    D100, I100 = idx.search(q, 100)
    _check(g["D100"], g["I100"], D100, I100)
    Dm, Im = idx.search(q, 10, native.Filter().set_row_mask(so.pack_mask(mask)))
    _check(g["D_masked"], g["I_masked"], Dm, Im)
    idx.close()


# ------------------------------------------------------------- edge cases
def test_empty_and_underfull(native):
    idx = native.Index(768)
    D, I = idx.search(np.ones((1, 768), np.float32), 10)
    assert (I == -1).all() and (D == -np.finfo(np.float32).max).all()
    rng = np.random.default_rng(0)
    x = so.normalize_rows(rng.standard_normal((3, 768), dtype=np.float32))
    idx.add(x)
    D, I = idx.search(x[:1], 10)
    assert list(I[0][:3]) == list(so.flat_search(x, x[:1], 3)[1][0])
    assert (I[0][3:] == -1).all()
    idx.close()


@pytest.mark.parametrize("n", [1, 7, 8, 9, 31, 33, 255, 1000, 4097])
@pytest.mark.parametrize("d", [768, 100])
def test_ragged_sizes(native, n, d):
    rng = np.random.default_rng(n * 1000 + d)
    x = rng.standard_normal((n, d), dtype=np.float32)
    q = rng.standard_normal((3, d), dtype=np.float32)
    k = 5
    idx = native.Index(d)
    idx.add(x)
    D, I = idx.search(q, k)
    Dr, Ir = so.flat_search(x, q, k)
    _check(Dr, Ir, D, I, tol=2e-4 * np.sqrt(d))  # un-normalised rows: scores are O(sqrt(d))
    idx.close()


@pytest.mark.parametrize("k", [1, 10, 32, 33, 64, 100, 128])
def test_all_k(native, k):
    rng = np.random.default_rng(k)
    x = so.normalize_rows(rng.standard_normal((20000, 768), dtype=np.float32))
    q = so.normalize_rows(rng.standard_normal((2, 768), dtype=np.float32))
    idx = native.Index(768)
    idx.add(x)
    D, I = idx.search(q, k)
    Dr, Ir = so.flat_search_c(x, q, k)
    _check(Dr, Ir, D, I)
    assert (np.diff(D, axis=1) <= 0).all()  # sorted best first
    idx.close()


def test_add_normalizes_like_reference(native):
    # src/storage.py:347-350: x / (||x|| + 1e-8)
    rng = np.random.default_rng(5)
    raw = (rng.standard_normal((1000, 768)) * 3).astype(np.float32)
    idx = native.Index(768)
    idx.add(raw, normalize=True)
    got = idx.get_rows(0, 1000)
    np.testing.assert_allclose(got, so.normalize_rows(raw), rtol=0, atol=1e-7)
    idx.close()


def test_incremental_add_growth_and_ids(native):
    rng = np.random.default_rng(6)
    idx = native.Index(768)
    parts = [so.normalize_rows(rng.standard_normal((n, 768), dtype=np.float32)) for n in (5, 1500, 37, 9000)]
    first = [idx.add(p) for p in parts]
    assert first == [0, 5, 1505, 1542]
    x = np.concatenate(parts)
    assert idx.ntotal == x.shape[0]
    q = x[[3, 700, 1520, 10000]]
    D, I = idx.search(q, 4)
    assert list(I[:, 0]) == [3, 700, 1520, 10000]
    Dr, Ir = so.flat_search(x, q, 4)
    _check(Dr, Ir, D, I)
    idx.close()


# ---------------------------------------------------- medium vs C oracle
def test_scan_200k_vs_oracle(native):
    rng = np.random.default_rng(42)
    x = so.normalize_rows(rng.standard_normal((200_000, 768), dtype=np.float32))
    q = so.normalize_rows(rng.standard_normal((6, 768), dtype=np.float32))
    idx = native.Index(768)
    idx.add(x)
    for k in (10, 100):
        D, I = idx.search(q, k)
        Dr, Ir = so.flat_search_c(x, q, k)
        _check(Dr, Ir, D, I)
    # 5 % mask
    mask = rng.random(x.shape[0]) < 0.05
    D, I = idx.search(q, 10, native.Filter().set_row_mask(so.pack_mask(mask)))
    Dr, Ir = so.flat_search_c(x, q, 10, mask_words=so.pack_mask(mask))
    _check(Dr, Ir, D, I)
    assert mask[I].all()
    idx.close()


def test_two_phase_scan_proof_and_fallback(native):
    """Batch-1 / small-nq searches of a 768-d inner-product index stream the bf16 shadow rows first and prove the
    exact top-k from the 64 / 128 best of that pass (rescore768_kernel); what cannot be proven -- more rows within
    the rounding bound of the k-th score than the list holds -- is re-run by the fp32 scan on the device.  Either
    way the answer is the exact one, with the fp32 scan's scores (bit-identical to the batched path)."""
    rng = np.random.default_rng(123)
    n, d = 150_000, 768
    x = so.normalize_rows(rng.standard_normal((n, d), dtype=np.float32))
    q = so.normalize_rows(rng.standard_normal((15, d), dtype=np.float32))
    # query 0: 400 near-duplicates of the query, all inside the bf16 rounding bound of each other -> fallback
    near = rng.choice(n, size=400, replace=False)
    x[near] = so.normalize_rows(q[0] + 1e-4 * rng.standard_normal((400, d), dtype=np.float32))
    # query 1: 30 exact duplicates straddle the k-th place -> ids decide, proven from the list (30 < 64)
    dup = rng.choice(np.setdiff1d(np.arange(n), near), size=30, replace=False)
    x[dup] = so.normalize_rows(q[1] + 0.5 * so.normalize_rows(rng.standard_normal((1, d), dtype=np.float32)))[0]
    # query 2: scores 1e-3 apart around the k-th place (inside 2 eps: order must come from the fp32 re-score)
    lad = rng.choice(np.setdiff1d(np.arange(n), np.concatenate([near, dup])), size=24, replace=False)
    u = so.normalize_rows(rng.standard_normal((1, d), dtype=np.float32))[0]
    u -= (u @ q[2]) * q[2]
    u /= np.linalg.norm(u)
    for j, r in enumerate(lad):
        c = 0.9 - 1e-3 * j
        x[r] = c * q[2] + np.sqrt(1 - c * c) * u
    idx = native.Index(d)
    idx.add(x, normalize=False)
    for k in (1, 10, 32, 33, 64):
        D, I = idx.search(q, k)
        Dr, Ir = so.flat_search_c(x, q, k)
        _check(Dr, Ir, D, I)
    D, I = idx.search(q[:3], 10)
    assert set(I[0]) <= set(near.tolist()) and set(I[1]) <= set(dup.tolist())
    np.testing.assert_array_equal(I[1], np.sort(dup)[:10])          # ties: ascending id
    np.testing.assert_array_equal(I[2], lad[:10])                    # the ladder, in order
    # same scores, bit for bit, as the tensor-core batched path (both re-score with the fp32 scan's arithmetic)
    qb = np.concatenate([q, so.normalize_rows(rng.standard_normal((17, d), dtype=np.float32))])
    Db, Ib = idx.search(qb, 10)
    for i in range(15):
        D1, I1 = idx.search(q[i:i + 1], 10)
        np.testing.assert_array_equal(I1[0], Ib[i])
        np.testing.assert_array_equal(D1[0], Db[i])
    # orphaned rows (alive bits) stay on the two-phase path: kill the current winners and 2000 random rows
    alive = np.ones(n, np.uint8)
    alive[I[:3, :5].ravel()] = 0
    alive[rng.choice(n, size=2000, replace=False)] = 0
    idx.set_alive(alive)
    Da, Ia = idx.search(q, 10)
    Dr, Ir = so.flat_search_c(x, q, 10, mask_words=so.pack_mask(alive.astype(bool)))
    _check(Dr, Ir, Da, Ia)
    assert alive[Ia].all()
    idx.close()
    # un-normalised rows (norms 0.1 .. 30): the bound scales with the largest row norm
    y = x[:60_000] * rng.uniform(0.1, 30.0, size=(60_000, 1)).astype(np.float32)
    idy = native.Index(d)
    idy.add(y, normalize=False)
    D, I = idy.search(q[3:9], 10)
    Dr, Ir = so.flat_search_c(y, q[3:9], 10)
    _check(Dr, Ir, D, I, tol=TOL * 30)
    idy.close()


def test_batched_queries_vs_oracle(native):
    """nq >= 16 over >= 65536 rows takes the batched (tensor-core) entry."""
    rng = np.random.default_rng(43)
    x = so.normalize_rows(rng.standard_normal((100_000, 768), dtype=np.float32))
    q = so.normalize_rows(rng.standard_normal((200, 768), dtype=np.float32))
    x[5000] = x[17]          # exact duplicate rows -> tie broken by id
    q[0] = x[17]
    idx = native.Index(768)
    idx.add(x)
    for k in (10, 100):
        D, I = idx.search(q, k)
        Dr, Ir = so.flat_search_c(x, q, k)
        _check(Dr, Ir, D, I)
    assert list(I[0][:2]) == [17, 5000]
    mask = rng.random(x.shape[0]) < 0.05
    D, I = idx.search(q, 10, native.Filter().set_row_mask(so.pack_mask(mask)))
    Dr, Ir = so.flat_search_c(x, q, 10, mask_words=so.pack_mask(mask))
    _check(Dr, Ir, D, I)
    assert mask[I[I >= 0]].all()
    # batch-1 and batched paths return the same scores for the same query
    D1, I1 = idx.search(q[:1], 10)
    Db, Ib = idx.search(q, 10)
    _check(D1, I1, Db[:1], Ib[:1], tol=1e-6)
    idx.close()


# ----------------------------------------------------------------- filters
def _meta_rows(n, rng):
    projects = [f"/home/u/proj-{i}" for i in range(40)] + ["My-Awesome-Project", "my-awesome-project-2"]
    rows = []
    for i in range(n):
        day = int(rng.integers(0, 700))
        rows.append(dict(
            session_id=f"s{int(rng.integers(0, 500))}",
            project_name=projects[int(rng.integers(0, len(projects)))] if rng.random() > 0.02 else None,
            file_path=f"/f/{i % 97}.jsonl", chunk_type=["qa_pair", "code_block", "tool_usage"][i % 3],
            timestamp=f"2023-{1 + (day // 28) % 12:02d}-{1 + day % 28:02d}T{i % 24:02d}:00:00+00:00",
            has_code=bool(rng.random() < 0.4), has_tools=bool(rng.random() < 0.2),
            message_count=int(rng.integers(1, 9)), char_count=int(rng.integers(10, 4000)),
            word_count=int(rng.integers(1, 800))))
    return rows


FILTER_CASES = [
    {"has_code": True},
    {"project_name": "awesome"},
    {"project_name": "PROJ-1"},
    {"timestamp": {"gte": "2023-03-01", "lte": "2023-06-15T12:00:00+00:00"}},
    {"word_count": {"gt": 100, "lt": 300}},
    {"chunk_type": ["qa_pair", "tool_usage"]},
    {"session_id": "s7"},
    {"timestamp": {"gte": "2023-02"}, "project_name": "proj", "has_code": True},
    {"project_name": "no-such"},
    {"related_to": "ignored-key"},
]


@pytest.mark.parametrize("n", [1, 31, 32, 33, 5000])
def test_filter_mask_bit_exact(native, n):
    from claude_semantic_search_b200.filters import ColumnStore
    rng = np.random.default_rng(n)
    rows = _meta_rows(n, rng)
    x = so.normalize_rows(rng.standard_normal((n, 64), dtype=np.float32))
    idx = native.Index(64)
    idx.add(x)
    cols = ColumnStore()
    cols.append_rows(rows)
    cols.sync(idx)
    for f in FILTER_CASES:
        want = so.filter_mask(rows, f)
        flt = cols.compile(f)
        words, n_pass = idx.filter_mask(flt)
        np.testing.assert_array_equal(words, so.pack_mask(want), err_msg=str(f))
        assert n_pass == int(want.sum())
    # dead rows never pass
    alive = np.ones(n, np.uint8)
    alive[::3] = 0
    idx.set_alive(alive)
    words, n_pass = idx.filter_mask(cols.compile({"has_code": True}))
    want = so.filter_mask(rows, {"has_code": True}) & alive.astype(bool)
    np.testing.assert_array_equal(words, so.pack_mask(want))
    idx.close()


def test_filtered_search_prefix_property(native):
    """Reference result R (post-filter of the global top-100) is a prefix of the
    device prefilter result P."""
    from claude_semantic_search_b200.filters import ColumnStore
    rng = np.random.default_rng(9)
    n = 30000
    rows = _meta_rows(n, rng)
    x = so.normalize_rows(rng.standard_normal((n, 768), dtype=np.float32))
    idx = native.Index(768)
    idx.add(x)
    cols = ColumnStore()
    cols.append_rows(rows)
    cols.sync(idx)
    f = {"timestamp": {"gte": "2023-02"}, "project_name": "proj-1", "has_code": True}
    for qi in range(4):
        q = rng.standard_normal(768).astype(np.float32)
        R = so.storage_search(x, rows, q, top_k=10, filters=f)
        P = so.prefilter_search(x, rows, q, top_k=10, filters=f)
        D, I = idx.search(so.normalize_query(q).reshape(1, -1), 10, cols.compile(f))
        _check(np.array([[s for _, s in P]], np.float32), np.array([[i for i, _ in P]]),
               D[:, :len(P)], I[:, :len(P)])
        assert [i for i, _ in R] == list(I[0][:len(R)])
    idx.close()


# ------------------------------------------------------------ persistence
def test_faiss_file_roundtrip(native, tmp_path):
    rng = np.random.default_rng(11)
    x = so.normalize_rows(rng.standard_normal((777, 768), dtype=np.float32))
    idx = native.Index(768)
    idx.add(x)
    p = tmp_path / "embeddings.faiss"
    idx.save(p)
    raw = p.read_bytes()
    # faiss IndexFlatIP layout: fourcc, d, ntotal, 2 dummies, is_trained, metric, count, data
    assert raw[:4] == b"IxFI"
    d, ntotal = struct.unpack_from("<iq", raw, 4)
    assert (d, ntotal) == (768, 777)
    assert raw[32] == 1 and struct.unpack_from("<i", raw, 33)[0] == 0
    assert struct.unpack_from("<Q", raw, 37)[0] == 777 * 768
    np.testing.assert_array_equal(np.frombuffer(raw, np.float32, offset=45).reshape(777, 768), x)
    idx2 = native.Index(768)
    idx2.load(p)
    assert idx2.ntotal == 777
    np.testing.assert_array_equal(idx2.get_rows(0, 777), x)
    D, I = idx2.search(x[:2], 5)
    Dr, Ir = so.flat_search(x, x[:2], 5)
    _check(Dr, Ir, D, I)
    idx.close()
    idx2.close()


# ------------------------------------------------------------ compaction
@pytest.mark.parametrize("n,keep_frac", [(1, 1.0), (1000, 0.5), (70_000, 0.3), (70_000, 1.0), (5000, 0.0)])
def test_compact_on_device(native, n, keep_frac):
    """css_index_compact (HybridStorage.optimize): the kept rows, their columns and alive bits move
    to ids 0..n_keep-1 in order; searches and filters afterwards equal the oracle on the kept rows."""
    rng = np.random.default_rng(n + int(keep_frac * 10))
    d = 768
    x = so.normalize_rows(rng.standard_normal((n, d), dtype=np.float32))
    col = rng.integers(0, 50, size=n).astype(np.int32)
    alive = (rng.random(n) < 0.9).astype(np.uint8)
    idx = native.Index(d)
    idx.add(x)
    idx.set_column(2, col)
    idx.set_alive(alive)
    keep = np.flatnonzero(rng.random(n) < keep_frac).astype(np.int64)
    idx.compact(keep)
    assert idx.ntotal == keep.shape[0]
    if keep.shape[0]:
        np.testing.assert_array_equal(idx.get_rows(0, keep.shape[0]), x[keep])
        xk, ak = x[keep], alive[keep].astype(bool)
        q = so.normalize_rows(rng.standard_normal((20, d), dtype=np.float32))
        for qq in (q[:2], q):                      # streaming scan and batched path
            D, I = idx.search(qq, 10)              # dead rows never come back
            Dr, Ir = so.flat_search(xk, qq, 10, mask=ak)
            _check(Dr, Ir, D, I)
        flt = native.Filter().add_range(2, 10, 30)
        words, n_pass = idx.filter_mask(flt)
        want = ak & (col[keep] >= 10) & (col[keep] <= 30)
        np.testing.assert_array_equal(words, so.pack_mask(want))
        assert n_pass == int(want.sum())
        # appending after a compaction continues the dense id sequence
        first = idx.add(x[:3])
        assert first == keep.shape[0] and idx.ntotal == keep.shape[0] + min(3, n)
    else:
        D, I = idx.search(x[:1], 5)
        assert (I == -1).all()
    with pytest.raises(native.NativeError):
        idx.compact([0, 0])                        # not strictly ascending
    idx.close()


# -------------------------------------------- full-size property (config 2)
def test_one_million_rows_planted_needles(native):
    """BASELINE config 2 size.  The CPU oracle checks 4 queries in full; the rest
    is verified by planted needles whose exact top-10 is known a priori."""
    n, d = 1_000_000, 768
    rng = np.random.default_rng(42)
    idx = native.Index(d)
    idx.reserve(n)
    x = np.empty((n, d), np.float32)
    for s in range(0, n, 100_000):
        blk = so.normalize_rows(rng.standard_normal((100_000, d), dtype=np.float32))
        x[s:s + 100_000] = blk
    q = so.normalize_rows(np.random.default_rng(43).standard_normal((64, d), dtype=np.float32))
    # needles: for query j, rows ids[j, r] = normalise(q + sigma_r * noise): decreasing similarity
    ids = rng.choice(n, size=(64, 10), replace=False)
    for j in range(64):
        for r in range(10):
            v = q[j] + (0.05 + 0.05 * r) * so.normalize_rows(rng.standard_normal((1, d), dtype=np.float32))[0]
            x[ids[j, r]] = v / np.linalg.norm(v)
    idx.add(x)
    D, I = idx.search(q[:8], 10)                 # scan path
    np.testing.assert_array_equal(I, ids[:8])
    Dr, Ir = so.flat_search_c(x, q[:4], 10)
    _check(Dr, Ir, D[:4], I[:4])
    Db, Ib = idx.search(q, 10)                   # batched path
    np.testing.assert_array_equal(Ib, ids)
    assert np.abs(Db[:8] - D).max() < 1e-6
    # idempotence: same call, same answer (ticket counters reset correctly)
    D2, I2 = idx.search(q[:8], 10)
    np.testing.assert_array_equal(I2, I)
    np.testing.assert_array_equal(D2, D)
    idx.close()


def test_batched_overflow_falls_back_to_scan(native):
    """More near-ties than a candidate list holds: the query must be answered by the exact
    scan, never truncated (CSS_ERR_OVERFLOW is not an acceptable silent outcome)."""
    rng = np.random.default_rng(77)
    n, d = 80_000, 768
    x = so.normalize_rows(rng.standard_normal((n, d), dtype=np.float32))
    base = x[11].copy()
    dup = rng.choice(n, size=9000, replace=False)
    x[dup] = base                       # 9000 exact duplicates: ids decide the order
    q = so.normalize_rows(rng.standard_normal((40, d), dtype=np.float32))
    q[3] = base                         # this query sees 9000 rows at score 1.0
    idx = native.Index(d)
    idx.add(x)
    D, I = idx.search(q, 10)
    Dr, Ir = so.flat_search_c(x, q, 10)
    _check(Dr, Ir, D, I)
    np.testing.assert_array_equal(I[3], np.sort(np.union1d(dup, [11]))[:10])
    idx.close()


def test_batched_ragged_nq_and_unnormalized(native):
    rng = np.random.default_rng(78)
    n, d = 70_001, 768                   # ragged last tile
    x = (rng.standard_normal((n, d)) * rng.uniform(0.2, 3.0, size=(n, 1))).astype(np.float32)
    q = (rng.standard_normal((130, d)) * 2).astype(np.float32)   # 2 row blocks, ragged
    idx = native.Index(d)
    idx.add(x)
    D, I = idx.search(q, 7)
    Dr, Ir = so.flat_search_c(x, q, 7)
    _check(Dr, Ir, D, I, tol=2e-3)       # scores are O(100): 1e-4 absolute does not apply to raw IP
    idx.close()
