"""CPU-only: host half of the device filter (column codecs + clause compiler)
against the row-wise oracle restating src/storage.py:508-543."""
import numpy as np
import pytest

from claude_semantic_search_b200 import _native
from claude_semantic_search_b200.filters import COLUMN_INDEX, NULL, ColumnStore
from oracle import search_oracle as so


def eval_clauses_host(store: ColumnStore, flt) -> np.ndarray:
    """Evaluate a compiled native Filter with numpy, the way the S4 kernel does."""
    n = store.n
    ok = np.ones(n, dtype=bool)
    if flt is None:
        return ok
    for col, kind, lo, hi, bits, nbits in flt._clauses:
        v = store.codecs[col].codes[:n].astype(np.int64)
        notnull = v != NULL
        if kind == _native.CLAUSE_RANGE:
            ok &= notnull & (v >= lo) & (v <= hi)
        else:
            inr = notnull & (v >= 0) & (v < nbits)
            hit = np.zeros(n, dtype=bool)
            vv = v[inr]
            hit[inr] = ((bits[vv >> 5] >> (vv & 31).astype(np.uint32)) & 1).astype(bool)
            ok &= hit
    if flt.row_mask is not None:
        ok &= np.unpackbits(flt.row_mask.view(np.uint8), bitorder="little")[:n].astype(bool)
    return ok


def make_rows(n, seed=0):
    rng = np.random.default_rng(seed)
    projects = ["-Users-a-dev-Alpha", "-Users-a-dev-beta-service", "ALPHA-tools", "gamma", "Delta_alpha", None]
    rows = []
    for i in range(n):
        day = int(rng.integers(0, 730))
        ts = (np.datetime64("2023-01-01T00:00:00") + np.timedelta64(day, "D")
              + np.timedelta64(int(rng.integers(0, 86400)), "s"))
        rows.append(dict(
            session_id=f"sess-{int(rng.integers(0, 40))}",
            project_name=projects[int(rng.integers(0, len(projects)))],
            file_path=f"/p/{int(rng.integers(0, 25))}.jsonl",
            chunk_type=["qa_pair", "code_block", "tool_usage", "context_segment"][int(rng.integers(0, 4))],
            timestamp=str(ts) + "+00:00",
            has_code=int(rng.random() < 0.4),
            has_tools=int(rng.random() < 0.3),
            message_count=int(rng.integers(1, 9)),
            char_count=int(rng.integers(50, 2000)),
            word_count=int(rng.integers(5, 400)),
        ))
    return rows


FILTERS = [
    {"project_name": "alpha"},
    {"project_name": "ALPHA", "has_code": True},
    {"has_code": True},
    {"has_code": False, "has_tools": True},
    {"session_id": "sess-3"},
    {"timestamp": {"gte": "2023-06-01T00:00:00+00:00", "lte": "2023-12-31T23:59:59+00:00"}},
    {"timestamp": {"gt": "2024-01-01", "lt": "2024-03"}},
    {"timestamp": {"gte": "2030"}},
    {"word_count": {"gte": 5, "lt": 100}},
    {"word_count": {"gt": 10.5, "lte": 300.0}},
    {"chunk_type": ["qa_pair", "code_block"]},
    {"message_count": [1, 2, 3]},
    {"chunk_type": "tool_usage", "word_count": {"gte": 50}, "project_name": "a"},
    {"related_to": "chunk_1", "same_session": True},          # ignored keys (MCP)
    {"timestamp": {"gte": "2023-03-01T00:00:00+00:00", "lte": "2023-09-01T23:59:59+00:00"},
     "project_name": "dev", "has_code": True},
    {"session_id": {"gte": "sess-2", "lt": "sess-3"}},
    {"char_count": 100},
    {"has_code": 1},
    {"project_name": ["gamma", None]},
]


@pytest.mark.parametrize("filters", FILTERS)
def test_compiled_filter_matches_row_wise_oracle(filters):
    rows = make_rows(1500)
    # rows with NULL project under a *string equality/substring* filter are rejected by the
    # reference (None != value); ranges over NULL raise there, so keep NULLs out of ranged cols
    store = ColumnStore()
    store.append_rows(rows[:700])
    store.append_rows(rows[700:])       # incremental append path
    flt = store.compile(filters)
    got = eval_clauses_host(store, flt)
    want = so.filter_mask(rows, filters)
    np.testing.assert_array_equal(got, want)


def test_ordered_codes_preserve_string_order_under_random_inserts():
    rng = np.random.default_rng(5)
    store = ColumnStore()
    vals = [f"2024-{int(rng.integers(1, 13)):02d}-{int(rng.integers(1, 29)):02d}T{int(rng.integers(0, 24)):02d}:00:00"
            for _ in range(3000)]
    vals += ["2024-01-15T10:00:00", "2024-01-15T10:00:00Z", "2024-01-15T10:00:00.5+00:00", "2024-01-15"]
    for v in vals:                      # one by one: exercises gap splitting and rebalancing
        store.append_rows([{"timestamp": v}])
    codec = store.codecs[COLUMN_INDEX["timestamp"]]
    codes = codec.codes[:store.n]
    order_by_code = np.argsort(codes, kind="stable")
    sorted_vals = [vals[i] for i in order_by_code]
    assert sorted_vals == sorted(vals)
    # equal strings share a code
    assert len(set(zip(vals, codes.tolist()))) == len(set(vals))


def test_bulk_and_incremental_ordered_agree():
    rows = make_rows(5000, seed=9)
    a, b = ColumnStore(), ColumnStore()
    a.append_rows(rows)                                   # bulk path (>= 2048)
    for s in range(0, 5000, 500):
        b.append_rows(rows[s:s + 500])                    # incremental path
    f = {"timestamp": {"gte": "2023-06-01", "lte": "2024-02-01"}}
    np.testing.assert_array_equal(eval_clauses_host(a, a.compile(f)), eval_clauses_host(b, b.compile(f)))
    np.testing.assert_array_equal(eval_clauses_host(a, a.compile(f)), so.filter_mask(rows, f))


def test_batch_append_of_plain_columns_equals_per_value_path():
    """ColumnCodec._bulk_plain (int / dict columns, batches of >= 64 rows) assigns exactly the codes of the
    per-value path, falls back when a value does not fit (None in an int column, unhashable, out of range)."""
    rows = make_rows(3000, seed=4)
    rows[17]["message_count"] = None          # int column with a null: per-value path for that batch
    rows[400]["project_name"] = ["a", "list"]  # unhashable in a dict column: exotic row
    rows[900]["char_count"] = 2 ** 40          # out of int32 range: exotic row
    rows[1200]["has_code"] = np.bool_(True)
    a, b = ColumnStore(), ColumnStore()
    for s in range(0, 3000, 300):
        a.append_rows(rows[s:s + 300])                    # batch path where applicable
    for s in range(0, 3000, 50):
        b.append_rows(rows[s:s + 50])                     # below the batch threshold: per value
    for ca, cb in zip(a.codecs, b.codecs):
        if ca.kind == "ordered":
            continue                                      # order-preserving codes depend on insertion history
        np.testing.assert_array_equal(ca.codes[:ca.n], cb.codes[:cb.n], err_msg=ca.name)
        assert ca.values == cb.values and ca.exotic == cb.exotic, ca.name
    for f in ({"has_code": True}, {"project_name": "alpha"}, {"message_count": rows[0]["message_count"]}, {"char_count": {"lt": 500}},
              {"has_code": True, "timestamp": {"gte": "2023-06-01"}}):
        want = so.filter_mask(rows, f)
        for st in (a, b):
            c = st.compile(f)
            got = eval_clauses_host(st, c)
            np.testing.assert_array_equal(got, want, err_msg=str(f))


def test_range_on_string_column_with_wrong_bound_type_raises_like_reference():
    store = ColumnStore()
    store.append_rows([{"timestamp": "2024-01-01"}])
    with pytest.raises(TypeError):
        store.compile({"timestamp": {"gte": 5}})
    with pytest.raises(TypeError):
        so.matches_filters({"timestamp": "2024-01-01"}, {"timestamp": {"gte": 5}})
