"""oracle/search_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

CPU restatement of half B of the hot path: the reference's HybridStorage vector
maths and the faiss-cpu flat index it calls.

Where the arithmetic lives: `faiss-cpu>=1.11.0` (reference pyproject.toml:9) is a
third-party dependency that is NOT vendored under /root/reference and is not
installable here (no wheel, no network).  Its published algorithm for
IndexFlatIP / IndexFlatL2 (exhaustive float32 scoring + k-selection, results
best-first, unfilled slots id -1) is restated below; parity is anchored on the
reference's own call sites and known-answer tests:

  add_chunks normalisation      src/storage.py:343-350
  query normalisation + k'      src/storage.py:424-436
  candidate walk / post-filter  src/storage.py:438-492
  _matches_filters              src/storage.py:508-543
  pins                          tests/test_storage.py:277-345,617-647,
                                tests/test_integration.py:312-353,
                                tests/test_environment_setup.py:199-220

Pinning: see tests/test_oracle_pins.py (golden vectors under tests/golden/ were
produced by oracle/make_golden.py; the score values quoted in SURVEY.md section 8c
are asserted there).
"""
from __future__ import annotations

import ctypes
import json
import os
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

METRIC_IP = 0
METRIC_L2 = 1
FLT_MAX = np.finfo(np.float32).max


# --------------------------------------------------------------------------
# vector maths
# --------------------------------------------------------------------------
def normalize_rows(x: np.ndarray) -> np.ndarray:
    """Row normalisation applied by add_chunks (src/storage.py:347-350)."""
    x = np.asarray(x, dtype=np.float32)
    norms = np.linalg.norm(x, axis=1, keepdims=True)
    return (x / (norms + 1e-8)).astype(np.float32)


def normalize_query(q: np.ndarray) -> np.ndarray:
    """Query normalisation of search (src/storage.py:425-429): the division runs
    in the *input* dtype, the cast to float32 comes after."""
    q = np.asarray(q)
    q = q / (np.linalg.norm(q) + 1e-8)
    return q.reshape(1, -1).astype(np.float32)


def _order_best_first(keys: np.ndarray, ids: np.ndarray) -> np.ndarray:
    # (key desc, id asc): lexsort sorts by last key first
    return np.lexsort((ids, -keys.astype(np.float64)))


def flat_search(x: np.ndarray, q: np.ndarray, k: int, metric: int = METRIC_IP,
                mask: Optional[np.ndarray] = None, block: int = 262144) -> Tuple[np.ndarray, np.ndarray]:
    """faiss IndexFlat{IP,L2}.search restated: exact float32 scores, best-first,
    ties by ascending id, unfilled slots (-FLT_MAX | FLT_MAX, -1).

    mask: optional bool array [n] (True = row may be returned).
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, x.shape[1] if x.ndim == 2 and x.shape[0] else q.shape[-1])
    n = x.shape[0]
    nq = q.shape[0]
    D = np.full((nq, k), -FLT_MAX if metric == METRIC_IP else FLT_MAX, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    if n == 0:
        return D, I
    best_keys = [np.empty(0, np.float32) for _ in range(nq)]
    best_ids = [np.empty(0, np.int64) for _ in range(nq)]
    for r0 in range(0, n, block):
        xb = x[r0:r0 + block]
        if metric == METRIC_IP:
            s = q @ xb.T
        else:
            # squared L2, computed the direct way (no ||x||^2 - 2qx expansion)
            s = -((q[:, None, :] - xb[None, :, :]) ** 2).sum(-1) if nq * xb.shape[0] * x.shape[1] < 5e7 else \
                -np.stack([((xb - q[i]) ** 2).sum(-1) for i in range(nq)])
        s = s.astype(np.float32)
        ids = np.arange(r0, r0 + xb.shape[0], dtype=np.int64)
        mb = None if mask is None else np.asarray(mask[r0:r0 + xb.shape[0]], dtype=bool)
        for i in range(nq):
            keys_i, ids_i = s[i], ids
            if mb is not None:
                keys_i, ids_i = keys_i[mb], ids_i[mb]
            if keys_i.shape[0] > 4 * k:
                # keep everything >= the k-th largest (ties included), then order exactly
                kth = np.partition(keys_i, keys_i.shape[0] - k)[keys_i.shape[0] - k]
                sel = keys_i >= kth
                keys_i, ids_i = keys_i[sel], ids_i[sel]
            ck = np.concatenate([best_keys[i], keys_i])
            ci = np.concatenate([best_ids[i], ids_i])
            o = _order_best_first(ck, ci)[:k]
            best_keys[i], best_ids[i] = ck[o], ci[o]
    for i in range(nq):
        m = best_keys[i].shape[0]
        D[i, :m] = best_keys[i] if metric == METRIC_IP else -best_keys[i]
        I[i, :m] = best_ids[i]
    return D, I


# --------------------------------------------------------------------------
# filter predicate (row-wise), src/storage.py:508-543
# --------------------------------------------------------------------------
def matches_filters(row: Dict[str, Any], filters: Dict[str, Any]) -> bool:
    """Row-wise truth value of the reference's _matches_filters."""
    for key, want in filters.items():
        if key not in row:           # unknown column: ignored (:513-514)
            continue
        have = row[key]
        if isinstance(want, dict):   # range (:518-527); raises TypeError on None like the reference
            if "gte" in want and have < want["gte"]:
                return False
            if "lte" in want and have > want["lte"]:
                return False
            if "gt" in want and have <= want["gt"]:
                return False
            if "lt" in want and have >= want["lt"]:
                return False
        elif isinstance(want, list):  # membership (:528-531)
            if have not in want:
                return False
        elif key == "project_name" and isinstance(want, str) and isinstance(have, str):
            if want.lower() not in have.lower():   # case-insensitive substring (:534-537)
                return False
        elif have != want:            # exact (:538-541)
            return False
    return True


def filter_mask(rows: Sequence[Optional[Dict[str, Any]]], filters: Optional[Dict[str, Any]]) -> np.ndarray:
    """Bool mask over faiss ids: rows[i] is the SQLite row of the chunk bound to
    faiss id i, or None when the id is an orphan (src/storage.py:449-456)."""
    out = np.zeros(len(rows), dtype=bool)
    for i, r in enumerate(rows):
        if r is None:
            continue
        out[i] = True if not filters else matches_filters(r, filters)
    return out


def pack_mask(mask: np.ndarray) -> np.ndarray:
    """bool[n] -> uint32 words, bit i%32 of word i//32 = row i."""
    n = mask.shape[0]
    padded = np.zeros((n + 31) // 32 * 32, dtype=np.uint8)
    padded[:n] = mask.astype(np.uint8)
    return np.packbits(padded.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1)


# --------------------------------------------------------------------------
# HybridStorage.search restated (post-filter semantics), src/storage.py:408-492
# --------------------------------------------------------------------------
def storage_search(x_norm: np.ndarray, rows: Sequence[Optional[Dict[str, Any]]], query: np.ndarray,
                   top_k: int = 10, similarity_threshold: float = 0.0, max_results: int = 100,
                   filters: Optional[Dict[str, Any]] = None, normalize: bool = True,
                   metric: int = METRIC_IP) -> List[Tuple[int, float]]:
    """Returns [(faiss_id, similarity)] exactly as the reference would emit them:
    global top-k' (k' = min(max_results, ntotal)), walked best-first, dropping
    below-threshold scores, orphans and rows failing the filter, stopping at top_k."""
    n = x_norm.shape[0]
    if n == 0:
        return []
    q = normalize_query(query) if normalize else np.asarray(query).reshape(1, -1).astype(np.float32)
    kk = min(max_results, n)
    if kk == 0:
        return []
    D, I = flat_search(x_norm, q, kk, metric)
    out: List[Tuple[int, float]] = []
    for score, fid in zip(D[0], I[0]):
        score = float(score)
        if score < similarity_threshold:
            continue
        if fid < 0 or rows[fid] is None:
            continue
        if filters and not matches_filters(rows[fid], filters):
            continue
        out.append((int(fid), score))
        if len(out) >= top_k:
            break
    return out


def prefilter_search(x_norm: np.ndarray, rows: Sequence[Optional[Dict[str, Any]]], query: np.ndarray,
                     top_k: int = 10, filters: Optional[Dict[str, Any]] = None,
                     normalize: bool = True, metric: int = METRIC_IP) -> List[Tuple[int, float]]:
    """The device-side prefilter semantics: top_k over the rows
    that pass the filter and are alive.  The reference's result is a prefix of this."""
    q = normalize_query(query) if normalize else np.asarray(query).reshape(1, -1).astype(np.float32)
    mask = filter_mask(rows, filters)
    D, I = flat_search(x_norm, q, top_k, metric, mask=mask)
    return [(int(i), float(d)) for d, i in zip(D[0], I[0]) if i >= 0]


# --------------------------------------------------------------------------
# tolerance-aware comparison (north_star: scores within 1e-4; ids identical
# except where adjacent score gaps fall below that tolerance)
# --------------------------------------------------------------------------
def compare_topk(D_ref: np.ndarray, I_ref: np.ndarray, D_got: np.ndarray, I_got: np.ndarray,
                 tol: float = 1e-4) -> Tuple[bool, str]:
    D_ref = np.asarray(D_ref, np.float32).reshape(-1, D_ref.shape[-1])
    I_ref = np.asarray(I_ref).reshape(D_ref.shape)
    D_got = np.asarray(D_got, np.float32).reshape(D_ref.shape)
    I_got = np.asarray(I_got).reshape(D_ref.shape)
    filled = I_ref >= 0
    if not np.array_equal(filled, I_got >= 0):
        return False, "different number of filled slots"
    if filled.any() and np.abs(D_ref[filled] - D_got[filled]).max() > tol:
        return False, f"score diff {np.abs(D_ref[filled] - D_got[filled]).max():.3e} > {tol}"
    bad = (I_ref != I_got) & filled
    for qi, j in zip(*np.nonzero(bad)):
        # a differing id is allowed only if it is a near-tie: the reference id
        # must appear in the candidate's list within tol of its score or the
        # k-th score must be within tol (boundary swap)
        ref_id, ref_s = I_ref[qi, j], D_ref[qi, j]
        where = np.nonzero(I_got[qi] == ref_id)[0]
        if where.size:
            if abs(float(D_got[qi, where[0]]) - float(ref_s)) > tol or abs(float(D_ref[qi, where[0]]) - float(ref_s)) > tol:
                return False, f"query {qi}: id {ref_id} moved across a gap > tol"
        else:
            kth = D_ref[qi][filled[qi]][-1]
            if abs(float(ref_s) - float(kth)) > tol:
                return False, f"query {qi}: id {ref_id} (score {ref_s}) missing, not a boundary tie"
    return True, "ok"


# --------------------------------------------------------------------------
# C restatement (oracle/flat_ip.c), used for the multi-threaded CPU baseline
# --------------------------------------------------------------------------
_C_LIB = None


def c_lib() -> ctypes.CDLL:
    global _C_LIB
    if _C_LIB is None:
        here = Path(__file__).resolve().parent
        so = here / "_build" / "liboracle_flat.so"
        if not so.exists():
            import subprocess
            so.parent.mkdir(exist_ok=True)
            subprocess.run(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-o", str(so),
                            str(here / "flat_ip.c"), "-lm"], check=True)
        lib = ctypes.CDLL(str(so))
        lib.oracle_flat_search.restype = ctypes.c_int
        lib.oracle_flat_search.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.oracle_normalize_rows.restype = None
        lib.oracle_normalize_rows.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
        lib.oracle_num_threads.restype = ctypes.c_int
        _C_LIB = lib
    return _C_LIB


def flat_search_c(x: np.ndarray, q: np.ndarray, k: int, metric: int = METRIC_IP,
                  mask_words: Optional[np.ndarray] = None, nthreads: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    lib = c_lib()
    x = np.ascontiguousarray(x, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, x.shape[1])
    nq = q.shape[0]
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    mw = None
    if mask_words is not None:
        mw = np.ascontiguousarray(mask_words, dtype=np.uint32)
    rc = lib.oracle_flat_search(x.ctypes.data, x.shape[0], x.shape[1], q.ctypes.data, nq, k, metric,
                                mw.ctypes.data if mw is not None else None, D.ctypes.data, I.ctypes.data, nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle_flat_search failed: {rc}")
    return D, I
