"""Host logic of HybridStorage (src/storage.py's SQLite side: add_chunks rows, id maps, the result assembly of
search, deletions) with the device index replaced by a numpy stand-in -- the CUDA index itself is covered by the
-m gpu suite; this file keeps the Python around it honest on a machine without a GPU."""
import numpy as np
import pytest

from claude_semantic_search_b200 import Chunk, HybridStorage, SearchConfig, StorageConfig
from claude_semantic_search_b200 import hybrid_storage as hs


class _FakeNative:
    """Exact inner-product top-k over the stored rows, honouring the alive bytes; filters are ignored."""

    def __init__(self, dim):
        self.dim = dim
        self.x = np.zeros((0, dim), np.float32)
        self.alive = np.zeros(0, bool)
        self.calls = 0
        self.alive_calls = 0

    @property
    def ntotal(self):
        return self.x.shape[0]

    def add(self, x, normalize=False):
        x = np.asarray(x, np.float32)
        if normalize:
            x = x / (np.linalg.norm(x, axis=1, keepdims=True) + 1e-8)
        first = self.ntotal
        self.x = np.concatenate([self.x, x])
        self.alive = np.concatenate([self.alive, np.ones(x.shape[0], bool)])
        return first

    def set_alive(self, alive, start=0):
        a = np.asarray(alive).astype(bool)
        self.alive[start:start + a.shape[0]] = a

    def set_alive_ids(self, ids, alive):
        self.alive_calls += 1
        self.alive[np.asarray(ids, np.int64)] = bool(alive)

    def search(self, q, k, flt=None):
        self.calls += 1
        s = (self.x @ np.asarray(q, np.float32).reshape(-1)).astype(np.float32)
        s[~self.alive] = -np.inf
        order = np.lexsort((np.arange(s.shape[0]), -s))[:k]
        D = np.full((1, k), -np.finfo(np.float32).max, np.float32)
        I = np.full((1, k), -1, np.int64)
        live = [i for i in order if np.isfinite(s[i])]
        D[0, :len(live)] = s[live]
        I[0, :len(live)] = live
        return D, I

    def __getattr__(self, name):   # set_column, close, ...: nothing to do on the host
        return lambda *a, **k: None


class _FakeIndex:
    def __init__(self, dim, *a, **k):
        self._native = _FakeNative(dim)

    @property
    def ntotal(self):
        return self._native.ntotal


@pytest.fixture
def storage(tmp_path, monkeypatch):
    monkeypatch.setattr(hs.faiss_compat, "IndexFlatIP", _FakeIndex)
    monkeypatch.setattr(hs.faiss_compat, "IndexFlatL2", _FakeIndex)
    st = HybridStorage(StorageConfig(data_dir=str(tmp_path), use_gpu=True, auto_save=False))
    st.initialize()
    yield st
    st.db.close()


def _chunks(n, rng, as_array):
    emb = rng.standard_normal((n, 768)).astype(np.float32)
    return [Chunk(id=f"c{i:04d}", text=f"text {i}", metadata=dict(
        session_id=f"s{i % 5}", project_name=f"/p/{i % 3}", file_path=f"/f/{i % 10}.jsonl", chunk_type="qa_pair",
        timestamp=f"2024-01-{1 + i % 28:02d}T10:00:00+00:00", has_code=bool(i & 1), has_tools=False, message_count=2,
        char_count=10 + i, word_count=2), embedding=emb[i] if as_array else emb[i].tolist()) for i in range(n)], emb


@pytest.mark.parametrize("as_array", [False, True])
def test_add_search_delete_roundtrip(storage, as_array):
    rng = np.random.default_rng(5)
    chunks, emb = _chunks(300, rng, as_array)
    storage.add_chunks(chunks)
    assert storage.faiss_index.ntotal == 300 and storage.total_chunks == 300
    assert storage.chunk_id_to_faiss_id["c0007"] == 7 and storage.faiss_id_to_chunk_id[299] == "c0299"
    x = emb / (np.linalg.norm(emb, axis=1, keepdims=True) + 1e-8)
    q = emb[42] * 3.0
    res = storage.search(q, SearchConfig(top_k=10))
    want = np.argsort(-(x @ (q / (np.linalg.norm(q) + 1e-8))), kind="stable")[:10]
    assert [r.chunk_id for r in res] == [f"c{i:04d}" for i in want]
    assert res[0].chunk_id == "c0042" and abs(res[0].similarity - 1.0) < 1e-5
    assert res[0].text == "text 42" and res[0].metadata["char_count"] == 52 and res[0].chunk.id == "c0042"
    assert storage.faiss_index._native.calls == 1                      # one device search per query
    # projections of the result: text / metadata are only fetched when asked for
    bare = storage.search(q, SearchConfig(top_k=3, include_text=False, include_metadata=False))
    assert [r.chunk_id for r in bare] == [r.chunk_id for r in res[:3]]
    assert all(r.text is None and r.chunk is None for r in bare)
    only_md = storage.search(q, SearchConfig(top_k=3, include_text=False))
    assert only_md[0].metadata["session_id"] == "s2" and only_md[0].text is None
    # similarity threshold cuts the tail
    cut = storage.search(q, SearchConfig(top_k=10, similarity_threshold=res[4].similarity))
    assert [r.chunk_id for r in cut] == [r.chunk_id for r in res[:5]]
    # deletions: the row is orphaned (alive bit cleared, maps updated), later hits move up
    assert storage.delete_chunk("c0042") and not storage.delete_chunk("c0042")
    calls0 = storage.faiss_index._native.alive_calls
    removed = storage.remove_chunks_for_file("/f/3.jsonl")
    assert removed == 30
    assert storage.faiss_index._native.alive_calls == calls0 + 1       # ONE device call for the 30 orphaned rows
    res2 = storage.search(q, SearchConfig(top_k=10))
    gone = {"c0042"} | {f"c{i:04d}" for i in range(300) if i % 10 == 3}
    want2 = [f"c{i:04d}" for i in np.argsort(-(x @ (q / (np.linalg.norm(q) + 1e-8))), kind="stable") if f"c{i:04d}" not in gone][:10]
    assert [r.chunk_id for r in res2] == want2
    assert storage.get_chunk_by_id("c0042") is None and storage.get_chunk_by_id("c0001").text == "text 1"


def test_rows_deleted_behind_the_index_are_skipped(storage):
    """A chunk row that vanished from SQLite behind the storage object's back (its vector still alive) is
    skipped, as the reference skips orphans (src/storage.py:449-451); rows removed through the API clear the
    alive bit instead, so the device search never returns them and the list stays full (previous test)."""
    rng = np.random.default_rng(6)
    chunks, emb = _chunks(50, rng, True)
    storage.add_chunks(chunks)
    q = emb[9]
    first = storage.search(q, SearchConfig(top_k=5))
    storage.db.execute("DELETE FROM chunks WHERE id = ?", (first[1].chunk_id,))
    storage.db.commit()
    again = storage.search(q, SearchConfig(top_k=5))
    assert [r.chunk_id for r in again] == [r.chunk_id for i, r in enumerate(first) if i != 1]
