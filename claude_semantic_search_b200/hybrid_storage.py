"""HybridStorage: the reference's storage API (src/storage.py) on the B200 path.

Outer seam of the drop-in: same constructor, config
dataclasses, method names, argument meaning and error behaviour as the
reference's `HybridStorage`, so `SemanticSearchCLI`, the MCP server and the
watcher call it unchanged.  What differs is underneath:

  * vectors live in HBM behind libcss_b200.so (css_index_*), not in faiss-cpu;
  * `search` runs the exact top-k on the device with the metadata filter applied
    as a device-side bitmask prefilter (S4 + S2/S3) instead of a Python
    post-filter over the global top-100.  `StorageConfig.filter_mode="reference"`
    restores the reference's truncated result R; the default "prefilter" returns
    P = best top_k of the rows that pass, of which R is always a prefix;
  * `initialize()` is O(1) when the on-disk index is unchanged (the reference
    re-reads the whole file on every query, src/cli.py:237);
  * `save_index()` appends only the new rows to the faiss-format file.

SQLite bookkeeping (schema src/storage.py:153-218) is kept byte-compatible so an
existing data directory opens as-is.  No CPU fallback: `initialize()` raises if
no sm_100 device is present.
"""
from __future__ import annotations

import json
import logging
import os
import sqlite3
import struct
from dataclasses import dataclass
from datetime import datetime
from pathlib import Path
from typing import Any, Dict, List, Optional

import numpy as np

from . import _native, faiss_compat
from .chunk import Chunk
from .filters import COLUMN_INDEX, DEVICE_COLUMNS, TABLE_COLUMNS, ColumnStore, pack_bits


@dataclass
class StorageConfig:
    """Same knobs as the reference's StorageConfig (src/storage.py:43-58) plus
    `device` and `filter_mode`."""

    data_dir: str = "~/.claude-semantic-search/data"
    db_name: str = "metadata.db"
    index_name: str = "embeddings.faiss"
    embedding_dim: int = 768
    index_type: str = "flat"
    ivf_nlist: int = 100
    hnsw_m: int = 16
    normalize_embeddings: bool = True
    auto_save: bool = True
    backup_enabled: bool = True
    use_gpu: bool = False
    gpu_memory_fraction: float = 0.8
    device: int = 0
    devices: Optional[List[int]] = None  # several GPUs of this box: ONE index row-sharded over them (css_index_create_sharded)
    filter_mode: str = "prefilter"  # "prefilter" | "reference"


@dataclass
class SearchConfig:
    """src/storage.py:61-69."""

    top_k: int = 10
    similarity_threshold: float = 0.0
    include_metadata: bool = True
    include_text: bool = True
    max_results: int = 100


@dataclass
class SearchResult:
    """src/storage.py:72-80."""

    chunk_id: str
    similarity: float
    chunk: Optional[Chunk] = None
    metadata: Optional[Dict[str, Any]] = None
    text: Optional[str] = None


_SCHEMA = [
    """CREATE TABLE IF NOT EXISTS chunks (
        id TEXT PRIMARY KEY, text TEXT NOT NULL, metadata TEXT, faiss_id INTEGER,
        session_id TEXT, project_name TEXT, file_path TEXT, chunk_type TEXT,
        timestamp DATETIME, has_code BOOLEAN, has_tools BOOLEAN, message_count INTEGER,
        char_count INTEGER, word_count INTEGER,
        created_at DATETIME DEFAULT CURRENT_TIMESTAMP,
        updated_at DATETIME DEFAULT CURRENT_TIMESTAMP)""",
    """CREATE TABLE IF NOT EXISTS files (
        path TEXT PRIMARY KEY, last_modified DATETIME, last_indexed DATETIME,
        chunk_count INTEGER DEFAULT 0)""",
] + [
    f"CREATE INDEX IF NOT EXISTS idx_chunks_{name} ON chunks({col})"
    for name, col in [("session", "session_id"), ("project", "project_name"),
                      ("timestamp", "timestamp"), ("type", "chunk_type"),
                      ("has_code", "has_code"), ("has_tools", "has_tools"),
                      ("faiss_id", "faiss_id")]
]

_META_COLS = [name for name, _ in DEVICE_COLUMNS]
_HEADER_NTOTAL_OFF = 8    # fourcc(4) + d(4)
_HEADER_COUNT_OFF = 37    # + ntotal(8) + 2 dummies(16) + is_trained(1) + metric(4)
_HEADER_BYTES = 45


class HybridStorage:
    def __init__(self, config: Optional[StorageConfig] = None) -> None:
        self.config: StorageConfig = config or StorageConfig()
        self.logger = logging.getLogger(__name__)
        self.data_dir = Path(self.config.data_dir).expanduser()
        self.data_dir.mkdir(parents=True, exist_ok=True)
        self.db_path = self.data_dir / self.config.db_name
        self.index_path = self.data_dir / self.config.index_name

        self.db: Optional[sqlite3.Connection] = None
        self.faiss_index: Optional[faiss_compat.Index] = None
        self.chunk_id_to_faiss_id: Dict[str, int] = {}
        self.faiss_id_to_chunk_id: Dict[int, str] = {}

        # kept for attribute compatibility; the device probe is deferred to initialize()
        self._gpu_capability = None
        self._gpu_resources = None
        self._is_gpu_index = False

        self.total_chunks = 0
        self.embedding_dim = self.config.embedding_dim

        self._columns = ColumnStore()
        self._file_sig = None        # (mtime_ns, size) of the index file as we last read/wrote it
        self._rows_on_disk = 0       # rows of the current index known to be in the file
        self._disk_appendable = False

    # ------------------------------------------------------------------ init
    def initialize(self) -> None:
        """Open SQLite, create/load the device index.  Idempotent: a second call
        with an unchanged index file does not touch the vectors."""
        if self.db is None:
            self.db = sqlite3.connect(str(self.db_path), check_same_thread=False)
            self.db.row_factory = sqlite3.Row
            cur = self.db.cursor()
            for stmt in _SCHEMA:
                cur.execute(stmt)
            self.db.commit()
        if self.faiss_index is not None and self._disk_is_current():
            return
        self._init_faiss()
        self._load_existing_data()
        self.logger.info("Storage initialized with %d chunks", self.total_chunks)

    def _disk_is_current(self) -> bool:
        if not self.index_path.exists():
            return self._file_sig is None
        st = self.index_path.stat()
        return self._file_sig == (st.st_mtime_ns, st.st_size)

    def _create_index(self) -> faiss_compat.Index:
        kind = self.config.index_type
        if kind == "flat":
            cls = faiss_compat.IndexFlatIP if self.config.normalize_embeddings else faiss_compat.IndexFlatL2
            return cls(self.embedding_dim, self.config.device, devices=self.config.devices)
        if kind in ("ivf", "hnsw"):
            raise NotImplementedError(f"index_type={kind!r}: only the flat index is served by the B200 path")
        raise ValueError(f"Unknown index type: {kind}")

    # reference name kept: tests and callers reach for it
    _create_cpu_index = _create_index

    def _init_faiss(self) -> None:
        self.faiss_index = self._create_index()
        self._is_gpu_index = True
        self._columns.reset()
        self._rows_on_disk = 0
        self._disk_appendable = False
        self._file_sig = None

    def _load_existing_data(self) -> None:
        if not self.index_path.exists():
            return
        try:
            self.faiss_index._native.load(self.index_path)
            st = self.index_path.stat()
            self._file_sig = (st.st_mtime_ns, st.st_size)
            self._rows_on_disk = self.faiss_index.ntotal
            self._disk_appendable = True
            self._rebuild_id_mappings()
            self._rebuild_columns()
        except _native.NoDeviceError:
            raise
        except Exception as e:  # corrupt file: start empty, like the reference (:314-316)
            self.logger.warning("Could not load existing FAISS index: %s", e)
            self._init_faiss()

    def _rebuild_id_mappings(self) -> None:
        self.chunk_id_to_faiss_id.clear()
        self.faiss_id_to_chunk_id.clear()
        for cid, fid in self.db.execute("SELECT id, faiss_id FROM chunks WHERE faiss_id IS NOT NULL"):
            self.chunk_id_to_faiss_id[cid] = fid
            self.faiss_id_to_chunk_id[fid] = cid
        self.total_chunks = len(self.chunk_id_to_faiss_id)

    def _rebuild_columns(self) -> None:
        """Device columns + alive bits from the chunks table (after a load)."""
        n = self.faiss_index.ntotal
        self._columns.reset()
        if n == 0:
            return
        # one SELECT, transposed at C speed; no per-row Python work (the reload of a 10 M-row index would
        # otherwise spend minutes here)
        q = f"SELECT faiss_id, {', '.join(_META_COLS)} FROM chunks WHERE faiss_id IS NOT NULL ORDER BY faiss_id"
        rows = self.db.execute(q).fetchall()
        alive = np.zeros(n, dtype=np.uint8)
        columns: Dict[str, Any] = {}
        if rows:
            cols = list(zip(*rows))
            fids = np.asarray(cols[0], dtype=np.int64)
            ok = (fids >= 0) & (fids < n)
            dense = bool(ok.all()) and fids.shape[0] == n and bool((fids == np.arange(n)).all())
            alive[fids[ok]] = 1
            for i, name in enumerate(_META_COLS):
                if dense:
                    columns[name] = cols[i + 1]
                else:
                    vals = np.empty(n, dtype=object)      # None = SQL NULL for the orphaned rows
                    src = np.empty(len(rows), dtype=object)
                    src[:] = cols[i + 1]
                    vals[fids[ok]] = src[ok]
                    columns[name] = vals.tolist()
        self._columns.append_columns(n, columns)
        self._columns.sync(self.faiss_index._native)
        if not alive.all():
            self.faiss_index._native.set_alive(alive, 0)

    # ------------------------------------------------------------------- add
    def add_chunks(self, chunks: List[Chunk]) -> None:
        if not chunks:
            return
        todo = [c for c in chunks if c.embedding is not None]
        if not todo:
            self.logger.warning("No chunks with embeddings to add")
            return
        if not self.faiss_index:
            raise RuntimeError("FAISS index not initialized")
        if not self.db:
            raise RuntimeError("Database not initialized")
        embs = [c.embedding for c in todo]
        if all(isinstance(e, np.ndarray) and e.dtype == np.float32 and e.ndim == 1 for e in embs):
            x = np.stack(embs)   # ndarray rows (EmbeddingConfig.embedding_as_ndarray): one 3 KB memcpy per chunk
        else:
            x = np.asarray(embs, dtype=np.float32)
        # SQLite rows and metadata first (json.dumps may raise): nothing has touched the device yet
        native = self.faiss_index._native
        first = native.ntotal
        now = datetime.now().isoformat()
        rows, metas = [], []
        rebinds = []
        new_ids: Dict[str, int] = {}
        for i, c in enumerate(todo):
            fid = first + i
            md = c.metadata
            old = new_ids.get(c.id, self.chunk_id_to_faiss_id.get(c.id))
            if old is not None and old != fid:
                rebinds.append((old, md))
            new_ids[c.id] = fid
            meta_row = {
                "session_id": md.get("session_id"), "project_name": md.get("project_name"),
                "file_path": md.get("file_path"), "chunk_type": md.get("chunk_type"),
                "timestamp": md.get("timestamp"), "has_code": md.get("has_code", False),
                "has_tools": md.get("has_tools", False), "message_count": md.get("message_count", 0),
                "char_count": md.get("char_count", 0), "word_count": md.get("word_count", 0),
            }
            metas.append(meta_row)
            rows.append((c.id, c.text, json.dumps(md), fid, *[meta_row[k] for k in _META_COLS], now))
        # normalisation x / (||x|| + 1e-8) happens on the device (S1)
        got = native.add(x, normalize=self.config.normalize_embeddings)
        assert got == first, (got, first)
        try:
            self.db.executemany(
                "INSERT OR REPLACE INTO chunks (id, text, metadata, faiss_id, " + ", ".join(_META_COLS) +
                ", updated_at) VALUES (" + ", ".join("?" * (5 + len(_META_COLS))) + ")", rows)
            self.db.commit()
        except Exception:
            # the vectors are on the device but their rows are not in SQLite: keep the device columns
            # aligned with the vector rows (NULL metadata) and orphan the rows, like a deleted chunk
            self.db.rollback()
            self._columns.append_rows([{} for _ in todo])
            native.set_alive_ids(np.arange(first, first + len(todo), dtype=np.int64), False)
            raise
        for cid, fid in new_ids.items():
            self.chunk_id_to_faiss_id[cid] = fid
        for i, c in enumerate(todo):
            self.faiss_id_to_chunk_id[first + i] = c.id
        self._columns.append_rows(metas)
        # a re-added chunk id keeps its old faiss row bound to the (replaced) SQLite row in the
        # reference: mirror by giving the old row the new metadata
        for old, md in rebinds:
            self._columns.set_row(old, {k: md.get(k) for k in _META_COLS}, None)
            for codec in self._columns.codecs:
                codec.needs_full_upload = True
        self.total_chunks += len(todo)
        if self.config.auto_save:
            self.save_index()
        self.logger.info("Added %d chunks to storage", len(todo))

    # ---------------------------------------------------------------- search
    def _compile_filter(self, filters: Optional[Dict[str, Any]]) -> Optional[_native.Filter]:
        if not filters:
            return None
        device_part = {k: v for k, v in filters.items() if k in COLUMN_INDEX}
        sql_part = {k: v for k, v in filters.items() if k in TABLE_COLUMNS and k not in COLUMN_INDEX}
        flt = self._columns.compile(device_part) if device_part else None
        if sql_part:
            # columns that are not mirrored on the device (id, text, ...): evaluate the
            # reference predicate over SQLite rows once, ship as a row bitmask
            n = self.faiss_index.ntotal
            mask = np.zeros(n, dtype=bool)
            cols = ", ".join(sorted(sql_part))
            for row in self.db.execute(f"SELECT faiss_id, {cols} FROM chunks WHERE faiss_id IS NOT NULL"):
                fid = row["faiss_id"]
                if 0 <= fid < n:
                    mask[fid] = self._matches_filters({k: row[k] for k in sql_part}, sql_part)
            if flt is None:
                flt = _native.Filter()
            if flt.row_mask is not None:
                prev = np.unpackbits(flt.row_mask.view(np.uint8), bitorder="little")[:n].astype(bool)
                mask &= prev
            flt.set_row_mask(pack_bits(mask))
        return flt

    def search(self, query_embedding, config: Optional[SearchConfig] = None,
               filters: Optional[Dict[str, Any]] = None) -> List[SearchResult]:
        cfg = config or SearchConfig()
        if not self.faiss_index:
            return []
        n = self.faiss_index.ntotal
        if n == 0:
            return []
        q = np.asarray(query_embedding)
        if self.config.normalize_embeddings:
            q = q / (np.linalg.norm(q) + 1e-8)   # in the input dtype, as the reference (:425-426)
        q = q.reshape(1, -1).astype(np.float32)
        k_ref = min(cfg.max_results, n)
        if k_ref == 0 or cfg.top_k <= 0:
            return []
        native = self.faiss_index._native
        self._columns.sync(native)

        if self.config.filter_mode == "reference":
            pairs = self._search_reference_mode(native, q, k_ref, cfg, filters)
        else:
            flt = self._compile_filter(filters)
            k = min(cfg.top_k, n)   # top_k > CSS_MAX_K: repeated exact passes inside _native.Index.search
            D, I = native.search(q, k, flt)
            pairs = [(int(i), float(d)) for d, i in zip(D[0], I[0])
                     if i >= 0 and float(d) >= cfg.similarity_threshold]

        # one SELECT for all hits, only the columns the result needs (the reference issues SELECT * per hit,
        # src/storage.py:452-470: ten round trips and ten 14-column dicts per query cost as much host time as the
        # whole device search)
        hits = []
        for fid, sim in pairs:
            cid = self.faiss_id_to_chunk_id.get(fid)
            if cid:
                hits.append((cid, sim))
        rows = self._get_chunk_rows([cid for cid, _ in hits], cfg.include_text, cfg.include_metadata)
        results: List[SearchResult] = []
        for cid, sim in hits:
            data = rows.get(cid)
            if data is None:
                continue
            text, meta = data
            res = SearchResult(chunk_id=cid, similarity=sim)
            md = None
            if cfg.include_metadata:
                md = json.loads(meta) if meta else {}
                res.metadata = md
            if cfg.include_text:
                res.text = text
            if cfg.include_metadata and cfg.include_text:
                res.chunk = Chunk(id=cid, text=text, metadata=md, embedding=None)
            results.append(res)
            if len(results) >= cfg.top_k:
                break
        return results

    def _get_chunk_rows(self, chunk_ids: List[str], want_text: bool, want_metadata: bool) -> Dict[str, tuple]:
        """{chunk id: (text | None, metadata JSON | None)} of the ids that still exist, in one query."""
        if not self.db:
            raise RuntimeError("Database not initialized")
        out: Dict[str, tuple] = {}
        cols = "id, " + ("text" if want_text else "NULL") + ", " + ("metadata" if want_metadata else "NULL")
        for s0 in range(0, len(chunk_ids), 500):   # stay below SQLite's bound-variable limit
            part = chunk_ids[s0:s0 + 500]
            q = f"SELECT {cols} FROM chunks WHERE id IN ({','.join('?' * len(part))})"
            for r in self.db.execute(q, part):
                out[r[0]] = (r[1], r[2])
        return out

    def _search_reference_mode(self, native, q, k_ref, cfg, filters):
        """R = (global top-k' incl. orphans) walked best-first with threshold, orphan and
        filter checks: the reference's exact result (src/storage.py:432-490)."""
        raw = _native.Filter(ignore_alive=True)
        D, I = native.search(q, k_ref, raw)
        words = None
        if filters:
            flt = self._compile_filter(filters)
            if flt is not None:
                words, _ = native.filter_mask(flt)
        out = []
        for d, i in zip(D[0], I[0]):
            i = int(i)
            if i < 0 or float(d) < cfg.similarity_threshold:
                continue
            if i not in self.faiss_id_to_chunk_id:
                continue
            if words is not None and not ((int(words[i >> 5]) >> (i & 31)) & 1):
                continue
            out.append((i, float(d)))
            if len(out) >= cfg.top_k:
                break
        return out

    # --------------------------------------------------------- row utilities
    def _get_chunk_data(self, chunk_id: str) -> Optional[Dict[str, Any]]:
        if not self.db:
            raise RuntimeError("Database not initialized")
        row = self.db.execute("SELECT * FROM chunks WHERE id = ?", (chunk_id,)).fetchone()
        return {k: row[k] for k in row.keys()} if row else None

    def _matches_filters(self, chunk_data: Dict[str, Any], filters: Dict[str, Any]) -> bool:
        """Host restatement kept for callers/tests that use it directly; the search
        path evaluates the same predicate on the device (filters.py + S4)."""
        for key, want in filters.items():
            if key not in chunk_data:
                continue
            have = chunk_data[key]
            if isinstance(want, dict):
                for op, bad in (("gte", lambda a, b: a < b), ("lte", lambda a, b: a > b),
                                ("gt", lambda a, b: a <= b), ("lt", lambda a, b: a >= b)):
                    if op in want and bad(have, want[op]):
                        return False
            elif isinstance(want, list):
                if have not in want:
                    return False
            elif key == "project_name" and isinstance(want, str) and isinstance(have, str):
                if want.lower() not in have.lower():
                    return False
            elif have != want:
                return False
        return True

    def _chunk_from_row(self, row) -> Chunk:
        return Chunk(id=row["id"], text=row["text"],
                     metadata=json.loads(row["metadata"]) if row["metadata"] else {}, embedding=None)

    def get_chunk_by_id(self, chunk_id: str) -> Optional[Chunk]:
        if not self.db:
            raise RuntimeError("Database not initialized")
        data = self._get_chunk_data(chunk_id)
        return self._chunk_from_row(data) if data else None

    def get_chunks_by_session(self, session_id: str) -> List[Chunk]:
        if not self.db:
            raise RuntimeError("Database not initialized")
        cur = self.db.execute("SELECT * FROM chunks WHERE session_id = ? ORDER BY timestamp", (session_id,))
        return [self._chunk_from_row(r) for r in cur.fetchall()]

    def get_chunks_by_project(self, project_name: str) -> List[Chunk]:
        if not self.db:
            raise RuntimeError("Database not initialized")
        cur = self.db.execute("SELECT * FROM chunks WHERE project_name = ? ORDER BY timestamp", (project_name,))
        return [self._chunk_from_row(r) for r in cur.fetchall()]

    # ------------------------------------------------------------- deletions
    def _kill_rows(self, faiss_ids: List[int]) -> None:
        """Orphaned vectors stay in the index (as in faiss) but are masked out."""
        if not faiss_ids or not self.faiss_index:
            return
        # one device call for the whole batch (css_index_set_alive_ids)
        self.faiss_index._native.set_alive_ids(np.asarray(faiss_ids, dtype=np.int64), False)

    def delete_chunk(self, chunk_id: str) -> bool:
        fid = self.chunk_id_to_faiss_id.get(chunk_id)
        if fid is None:
            return False
        cur = self.db.execute("DELETE FROM chunks WHERE id = ?", (chunk_id,))
        if cur.rowcount == 0:
            return False
        del self.chunk_id_to_faiss_id[chunk_id]
        self.faiss_id_to_chunk_id.pop(fid, None)
        self._kill_rows([fid])
        self.db.commit()
        self.total_chunks -= 1
        return True

    def delete_chunks_by_session(self, session_id: str) -> int:
        if not self.db:
            raise RuntimeError("Database not initialized")
        ids = [r[0] for r in self.db.execute("SELECT id FROM chunks WHERE session_id = ?", (session_id,))]
        return sum(1 for cid in ids if self.delete_chunk(cid))

    def remove_chunks_for_file(self, file_path: str) -> int:
        rows = self.db.execute("SELECT id, faiss_id FROM chunks WHERE file_path = ?", (file_path,)).fetchall()
        if not rows:
            return 0
        self.db.execute("DELETE FROM chunks WHERE file_path = ?", (file_path,))
        self.db.commit()
        dead = []
        for r in rows:
            self.chunk_id_to_faiss_id.pop(r["id"], None)
            if r["faiss_id"] is not None:
                self.faiss_id_to_chunk_id.pop(r["faiss_id"], None)
                dead.append(r["faiss_id"])
        self._kill_rows(dead)
        return len(rows)

    def clear_all_data(self) -> None:
        self.faiss_index = self._create_index()
        self._columns.reset()
        self._rows_on_disk = 0
        self._disk_appendable = False
        self.db.execute("DELETE FROM chunks")
        self.db.execute("DELETE FROM files")
        self.db.commit()
        self.chunk_id_to_faiss_id.clear()
        self.faiss_id_to_chunk_id.clear()
        self.total_chunks = 0
        if self.config.auto_save:
            self.save_index()
        self.logger.info("Cleared all data from storage")

    # ----------------------------------------------------------------- stats
    def get_stats(self) -> Dict[str, Any]:
        if not self.db:
            raise RuntimeError("Database not initialized")
        one = lambda sql: self.db.execute(sql).fetchone()[0]
        try:
            projects = self.get_all_projects()
        except Exception as e:
            self.logger.warning("Failed to get projects list: %s", e)
            projects = []
        fsz = self.index_path.stat().st_size if self.index_path.exists() else 0
        dsz = self.db_path.stat().st_size if self.db_path.exists() else 0
        stats = {
            "total_chunks": one("SELECT COUNT(*) FROM chunks"),
            "total_sessions": one("SELECT COUNT(DISTINCT session_id) FROM chunks"),
            "total_projects": one("SELECT COUNT(DISTINCT project_name) FROM chunks"),
            "projects": projects,
            "chunk_types": dict(self.db.execute("SELECT chunk_type, COUNT(*) FROM chunks GROUP BY chunk_type").fetchall()),
            "faiss_index_size": fsz,
            "database_size": dsz,
            "total_storage_size": fsz + dsz,
            "embedding_dimension": self.embedding_dim,
            "index_type": self.config.index_type,
            "use_gpu": self.config.use_gpu,
            "is_gpu_index": self._is_gpu_index,
        }
        if self._is_gpu_index:
            try:
                info = _native.device_info(self.config.device)
                stats["gpu_info"] = {
                    "gpu_available": True, "gpu_count": _native.device_count(),
                    "gpu_names": ["NVIDIA B200 (sm_%d%d)" % info["cc"]],
                    "status_message": "libcss_b200 device index",
                    "gpu_memory_total_gb": info["hbm_total"] / 1024 ** 3,
                    "gpu_memory_free_gb": info["hbm_free"] / 1024 ** 3,
                }
            except _native.NativeError:
                pass
        return stats

    def get_all_projects(self) -> List[str]:
        if not self.db:
            raise RuntimeError("Database not initialized. Call initialize() first.")
        cur = self.db.execute("SELECT DISTINCT project_name FROM chunks WHERE project_name IS NOT NULL "
                              "AND project_name != '' ORDER BY project_name")
        return [r[0] for r in cur.fetchall()]

    # --------------------------------------------------------- file tracking
    def update_file_info(self, file_path: str, chunk_count: int) -> None:
        if not self.db:
            raise RuntimeError("Database not initialized")
        try:
            modified = datetime.fromtimestamp(os.path.getmtime(file_path))
        except OSError:
            modified = datetime.now()
        self.db.execute("INSERT OR REPLACE INTO files (path, last_modified, last_indexed, chunk_count) "
                        "VALUES (?, ?, ?, ?)", (file_path, modified.isoformat(sep=" "),
                                                 datetime.now().isoformat(sep=" "), chunk_count))
        self.db.commit()

    def is_file_modified(self, file_path: str) -> bool:
        try:
            current = datetime.fromtimestamp(os.path.getmtime(file_path))
        except OSError:
            return True
        row = self.db.execute("SELECT last_modified FROM files WHERE path = ?", (file_path,)).fetchone()
        if not row or not row["last_modified"]:
            return True
        return current > datetime.fromisoformat(row["last_modified"])

    # ----------------------------------------------------------- persistence
    def save_index(self) -> None:
        """Persist the vectors in faiss IndexFlat format.  When only appends happened
        since the file was last written/read by this object, write just the new rows
        and patch the header (the reference rewrites N*d*4 bytes per add_chunks)."""
        if not self.faiss_index:
            self.logger.warning("No FAISS index to save")
            return
        native = self.faiss_index._native
        n = native.ntotal
        if (self._disk_appendable and self.index_path.exists() and self._disk_is_current()
                and n >= self._rows_on_disk):
            if n > self._rows_on_disk:
                new_rows = native.get_rows(self._rows_on_disk, n - self._rows_on_disk)
                with open(self.index_path, "r+b") as fh:
                    fh.seek(_HEADER_BYTES + self._rows_on_disk * self.embedding_dim * 4)
                    fh.write(new_rows.tobytes())
                    fh.truncate()
                    fh.seek(_HEADER_NTOTAL_OFF)
                    fh.write(struct.pack("<q", n))
                    fh.seek(_HEADER_COUNT_OFF)
                    fh.write(struct.pack("<Q", n * self.embedding_dim))
        else:
            native.save(self.index_path)
        st = self.index_path.stat()
        self._file_sig = (st.st_mtime_ns, st.st_size)
        self._rows_on_disk = n
        self._disk_appendable = True

    def backup(self, backup_dir: str) -> None:
        dest = Path(backup_dir)
        dest.mkdir(parents=True, exist_ok=True)
        if self.faiss_index and self.faiss_index.ntotal > 0:
            self.faiss_index._native.save(dest / self.config.index_name)
        if self.db_path.exists() and self.db:
            out = sqlite3.connect(str(dest / self.config.db_name))
            self.db.backup(out)
            out.close()
        self.logger.info("Backup created in %s", dest)

    def restore(self, backup_dir: str) -> None:
        src = Path(backup_dir)
        idx = src / self.config.index_name
        if idx.exists():
            self.faiss_index = faiss_compat.read_index(str(idx), self.config.device, self.config.devices)
            self._disk_appendable = False
        dbf = src / self.config.db_name
        if dbf.exists():
            self.db.close()
            self.db = sqlite3.connect(str(self.db_path), check_same_thread=False)
            self.db.row_factory = sqlite3.Row
            inp = sqlite3.connect(str(dbf))
            inp.backup(self.db)
            inp.close()
        self._rebuild_id_mappings()
        self._rebuild_columns()
        self.logger.info("Restored from backup in %s", src)

    def optimize(self) -> None:
        """VACUUM, then compact orphaned vectors out of the device index (the
        reference's rebuild is a stub that drops every vector, src/storage.py:944-969)."""
        self.db.execute("VACUUM")
        if self.faiss_index is None:
            return
        nt = self.faiss_index.ntotal
        # the reference's trigger (:936) plus rows orphaned by remove_chunks_for_file, which leaves
        # total_chunks untouched there (:817-846) and here
        if self.total_chunks != nt or len(self.faiss_id_to_chunk_id) != nt:
            self._rebuild_faiss_index()

    def _rebuild_faiss_index(self) -> None:
        rows = self.db.execute("SELECT id, faiss_id FROM chunks WHERE faiss_id IS NOT NULL ORDER BY faiss_id").fetchall()
        old = self.faiss_index
        kept = [(r["id"], r["faiss_id"]) for r in rows if 0 <= r["faiss_id"] < old.ntotal]
        # device-side gather of the surviving rows (css_index_compact): no vector leaves HBM
        old._native.compact([fid for _, fid in kept])
        self.db.executemany("UPDATE chunks SET faiss_id = ? WHERE id = ?",
                            [(i, cid) for i, (cid, _) in enumerate(kept)])
        self.db.commit()
        self._disk_appendable = False
        self._rebuild_id_mappings()
        self._rebuild_columns()
        if self.config.auto_save:
            self.save_index()

    def close(self) -> None:
        if self.config.auto_save and self.faiss_index is not None:
            self.save_index()
        if self.db:
            self.db.close()
            self.db = None
        self.logger.info("Storage closed")

    def __enter__(self) -> "HybridStorage":
        self.initialize()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb) -> None:
        self.close()
