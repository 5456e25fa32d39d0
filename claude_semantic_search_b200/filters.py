"""Host side of the device filter (S4): metadata columns -> int32 codes, and the
reference's filter dict -> css_filter clauses.

Semantics mirrored: HybridStorage._matches_filters (reference src/storage.py:508-543)
  * keys that are not columns of the `chunks` table are ignored        (:513-514)
  * dict value  -> range with gte / lte / gt / lt on the raw column      (:518-527)
                   (timestamps are TEXT: lexicographic string comparison)
  * list value  -> membership                                          (:528-531)
  * project_name + str -> case-insensitive substring                   (:534-537)
  * anything else -> exact equality (has_code: SQLite 0/1 == True)     (:538-541)

Device representation (bit-exact by construction):
  * "dict" columns (session_id, project_name, file_path, chunk_type): values are
    dictionary-encoded in order of first appearance; every predicate is evaluated
    on the host over the DISTINCT values only and shipped as an allowed-id bitset
    (CSS_CLAUSE_SET).
  * "ordered" columns (timestamp): order-preserving gapped int32 codes
    (code order == Python string order), so range predicates become a
    CSS_CLAUSE_RANGE on codes found by bisecting the query bounds.
  * "int" columns (has_code, has_tools, message_count, char_count, word_count):
    raw integers; ranges/equality become CSS_CLAUSE_RANGE.
  * SQL NULL is CSS_NULL_VALUE and never matches (the reference raises TypeError
    on a NULL under a range filter and rejects it otherwise).
Predicates that cannot be expressed this way are evaluated on the host over the
column codes and shipped as an explicit row bitmask.
"""
from __future__ import annotations

import bisect
import math
from typing import Any, Dict, Iterable, List, Optional, Sequence

import numpy as np

from . import _native

# column order on the device (index == css column id)
DEVICE_COLUMNS = [
    ("session_id", "dict"),
    ("project_name", "dict"),
    ("file_path", "dict"),
    ("chunk_type", "dict"),
    ("timestamp", "ordered"),
    ("has_code", "int"),
    ("has_tools", "int"),
    ("message_count", "int"),
    ("char_count", "int"),
    ("word_count", "int"),
]
COLUMN_INDEX = {name: i for i, (name, _) in enumerate(DEVICE_COLUMNS)}
# every column of the reference's chunks table (src/storage.py:160-181)
TABLE_COLUMNS = ["id", "text", "metadata", "faiss_id", "session_id", "project_name", "file_path",
                 "chunk_type", "timestamp", "has_code", "has_tools", "message_count", "char_count",
                 "word_count", "created_at", "updated_at"]

NULL = _native.NULL_VALUE
_GAP = 1 << 10
_INT_MIN = -(2 ** 31) + 1
_INT_MAX = 2 ** 31 - 1


def _range_ok(value, spec: Dict[str, Any]) -> bool:
    if "gte" in spec and value < spec["gte"]:
        return False
    if "lte" in spec and value > spec["lte"]:
        return False
    if "gt" in spec and value <= spec["gt"]:
        return False
    if "lt" in spec and value >= spec["lt"]:
        return False
    return True


class ColumnCodec:
    """Maps the Python values of one metadata column to int32 device codes."""

    def __init__(self, name: str, kind: str):
        self.name = name
        self.kind = kind
        self.codes = np.empty(0, dtype=np.int32)   # host copy, one code per row
        self.n = 0
        # dict
        self.value_to_id: Dict[Any, int] = {}
        self.values: List[Any] = []
        # ordered
        self.sorted_values: List[str] = []
        self.sorted_codes: List[int] = []
        self.needs_full_upload = False
        # rows whose value could not be represented (wrong type): host-evaluated
        self.exotic: Dict[int, Any] = {}

    # -- encoding -----------------------------------------------------------
    def _ensure(self, n: int) -> None:
        if n > self.codes.shape[0]:
            new = np.full(max(n, 2 * self.codes.shape[0], 1024), NULL, dtype=np.int32)
            new[: self.n] = self.codes[: self.n]
            self.codes = new

    def _encode_dict(self, v) -> int:
        i = self.value_to_id.get(v)
        if i is None:
            i = len(self.values)
            self.value_to_id[v] = i
            self.values.append(v)
        return i

    def _encode_ordered(self, v: str) -> int:
        i = bisect.bisect_left(self.sorted_values, v)
        if i < len(self.sorted_values) and self.sorted_values[i] == v:
            return self.sorted_codes[i]
        lo = self.sorted_codes[i - 1] if i > 0 else 0
        if i < len(self.sorted_codes):
            code = (lo + self.sorted_codes[i]) // 2
            fits = code > lo
        else:
            code = lo + _GAP
            fits = code < _INT_MAX
        self.sorted_values.insert(i, v)
        if fits:
            self.sorted_codes.insert(i, code)
            return code
        self.sorted_codes.insert(i, lo)  # placeholder (duplicates its left neighbour)
        self._rebalance_by_position(i)
        return self.sorted_codes[i]

    def _bulk_ordered(self, values: Sequence[Any], start: int) -> None:
        """Vectorised append for large batches: merge the distinct strings, re-space
        every code evenly, remap existing rows, encode the new ones by searchsorted."""
        strs = [v for v in values if isinstance(v, str)]
        merged = sorted(set(self.sorted_values).union(strs))
        m = len(merged)
        step = max(1, min(_GAP, (_INT_MAX - 1) // max(m, 1)))
        new_codes = (np.arange(1, m + 1, dtype=np.int64) * step)
        merged_arr = np.asarray(merged, dtype=str) if m else np.empty(0, dtype=str)
        live = self.codes[:start]
        if live.size and self.sorted_values:
            old_pos = np.searchsorted(merged_arr, np.asarray(self.sorted_values, dtype=str))
            lut_keys = np.asarray(self.sorted_codes, dtype=np.int64)
            lut_vals = new_codes[old_pos]
            nn = live != NULL
            pos = np.searchsorted(lut_keys, live[nn].astype(np.int64))
            live[nn] = lut_vals[pos].astype(np.int32)
        self.sorted_values = merged
        self.sorted_codes = [int(c) for c in new_codes]
        out = np.full(len(values), NULL, dtype=np.int32)
        is_str = np.fromiter((isinstance(v, str) for v in values), dtype=bool, count=len(values))
        if is_str.any():
            sv = np.asarray([v for v in values if isinstance(v, str)], dtype=str)
            out[is_str] = new_codes[np.searchsorted(merged_arr, sv)].astype(np.int32)
        for j, v in enumerate(values):
            if v is not None and not isinstance(v, str):
                self.exotic[start + j] = v
        self.codes[start:start + len(values)] = out
        self.n = start + len(values)
        self.needs_full_upload = True

    def _rebalance_by_position(self, inserted_at: int) -> None:
        # codes before insertion were strictly increasing except the placeholder at
        # `inserted_at` (== its left neighbour).  Build old->new from the old list
        # without the placeholder, then assign the placeholder its own new code.
        m = len(self.sorted_values)
        step = max(1, min(_GAP, (_INT_MAX - 1) // max(m, 1)))
        new_codes = [(_i + 1) * step for _i in range(m)]
        old_wo = self.sorted_codes[:inserted_at] + self.sorted_codes[inserted_at + 1:]
        new_wo = new_codes[:inserted_at] + new_codes[inserted_at + 1:]
        live = self.codes[: self.n]
        if live.size and old_wo:
            lut_keys = np.asarray(old_wo, dtype=np.int64)
            lut_vals = np.asarray(new_wo, dtype=np.int64)
            nn = live != NULL
            pos = np.searchsorted(lut_keys, live[nn].astype(np.int64))
            live[nn] = lut_vals[pos].astype(np.int32)
        self.sorted_codes = new_codes
        self.needs_full_upload = True

    def encode(self, v, row: int) -> int:
        if v is None:
            return NULL
        if self.kind == "dict":
            try:
                return self._encode_dict(v)
            except TypeError:  # unhashable
                self.exotic[row] = v
                return NULL
        if self.kind == "ordered":
            if isinstance(v, str):
                return self._encode_ordered(v)
            self.exotic[row] = v
            return NULL
        # int
        if isinstance(v, (bool, np.bool_)):
            return int(v)
        if isinstance(v, (int, np.integer)) and _INT_MIN <= int(v) <= _INT_MAX:
            return int(v)
        if isinstance(v, float) and v.is_integer() and _INT_MIN <= v <= _INT_MAX:
            return int(v)
        self.exotic[row] = v
        return NULL

    def append(self, values: Sequence[Any]) -> None:
        start = self.n
        self._ensure(start + len(values))
        if self.kind == "ordered" and len(values) >= 2048:
            self._bulk_ordered(list(values), start)
            return
        if len(values) >= 64 and self._bulk_plain(values, start):
            return
        for j, v in enumerate(values):
            # encode may rebalance (rewrites self.codes[:self.n]); keep n current
            c = self.encode(v, start + j)
            self.codes[start + j] = c
            self.n = start + j + 1

    def _bulk_plain(self, values: Sequence[Any], start: int) -> bool:
        """Batch append for the common shapes of an indexing batch (add_chunks hands over thousands of rows):
        `int` columns whose values are all Python / numpy bools or in-range integers, `dict` columns whose
        values are all hashable.  Same codes as the per-value path; False = not applicable, nothing changed."""
        n = len(values)
        if self.kind == "int":
            try:
                arr = np.asarray(values)
            except (ValueError, TypeError):
                return False
            if arr.ndim != 1 or arr.dtype.kind not in "biu":
                return False
            if arr.dtype.kind != "b" and arr.size and (arr.min() < _INT_MIN or arr.max() > _INT_MAX):
                return False
            self.codes[start:start + n] = arr.astype(np.int32)
            self.n = start + n
            return True
        if self.kind == "dict":
            v2i, vals = self.value_to_id, self.values
            out = np.empty(n, dtype=np.int32)
            try:
                for j, v in enumerate(values):
                    if v is None:
                        out[j] = NULL
                        continue
                    i = v2i.get(v)
                    if i is None:
                        i = len(vals)
                        v2i[v] = i
                        vals.append(v)
                    out[j] = i
            except TypeError:   # an unhashable value: the per-value path records it as exotic (ids assigned so far stand)
                return False
            self.codes[start:start + n] = out
            self.n = start + n
            return True
        return False

    def set_row(self, row: int, value) -> None:
        self.exotic.pop(row, None)
        self.codes[row] = self.encode(value, row)

    def reset(self) -> None:
        self.__init__(self.name, self.kind)

    # -- predicate compilation ------------------------------------------------
    def _distinct_allowed(self, pred) -> List[int]:
        out = []
        for i, v in enumerate(self.values):
            try:
                if pred(v):
                    out.append(i)
            except TypeError:
                raise
        return out

    def compile(self, want, flt: _native.Filter, host_masks: list) -> None:
        """Append the clause(s) for `column <op> want` to flt, or a host row mask."""
        col = COLUMN_INDEX[self.name]
        wants_null = want is None or (isinstance(want, list) and any(v is None for v in want))
        if self.exotic or wants_null:
            host_masks.append(self._host_mask(want))
            return
        if self.kind == "dict":
            if isinstance(want, dict):
                allowed = self._distinct_allowed(lambda v: _range_ok(v, want))
            elif isinstance(want, list):
                allowed = self._distinct_allowed(lambda v: v in want)
            elif self.name == "project_name" and isinstance(want, str):
                w = want.lower()
                allowed = self._distinct_allowed(lambda v: (w in v.lower()) if isinstance(v, str) else (v == want))
            else:
                allowed = self._distinct_allowed(lambda v: v == want)
            flt.add_set(col, allowed, len(self.values))
            return
        if self.kind == "ordered":
            if isinstance(want, dict):
                lo_i, hi_i = 0, len(self.sorted_values)  # slice [lo_i, hi_i) of sorted distinct values
                for op, bound in want.items():
                    if op not in ("gte", "lte", "gt", "lt"):
                        continue
                    if not isinstance(bound, str):
                        if self.n:
                            raise TypeError(f"'{op}' not supported between instances of 'str' and "
                                            f"'{type(bound).__name__}'")
                        continue
                    if op == "gte":
                        lo_i = max(lo_i, bisect.bisect_left(self.sorted_values, bound))
                    elif op == "gt":
                        lo_i = max(lo_i, bisect.bisect_right(self.sorted_values, bound))
                    elif op == "lte":
                        hi_i = min(hi_i, bisect.bisect_right(self.sorted_values, bound))
                    elif op == "lt":
                        hi_i = min(hi_i, bisect.bisect_left(self.sorted_values, bound))
                if lo_i >= hi_i:
                    flt.add_range(col, 1, 0)  # empty
                else:
                    flt.add_range(col, self.sorted_codes[lo_i], self.sorted_codes[hi_i - 1])
                return
            if isinstance(want, list):
                host_masks.append(self._host_mask(want))
                return
            i = bisect.bisect_left(self.sorted_values, want) if isinstance(want, str) else len(self.sorted_values)
            if i < len(self.sorted_values) and self.sorted_values[i] == want:
                flt.add_range(col, self.sorted_codes[i], self.sorted_codes[i])
            else:
                flt.add_range(col, 1, 0)
            return
        # int columns
        if isinstance(want, dict):
            lo, hi = _INT_MIN, _INT_MAX
            for op, bound in want.items():
                if op not in ("gte", "lte", "gt", "lt"):
                    continue
                if isinstance(bound, (bool, np.bool_)):
                    bound = int(bound)
                if not isinstance(bound, (int, float, np.integer, np.floating)):
                    if self.n:
                        raise TypeError(f"'{op}' not supported between instances of 'int' and "
                                        f"'{type(bound).__name__}'")
                    continue
                if isinstance(bound, float) and math.isnan(bound):
                    lo, hi = 1, 0
                    continue
                if op == "gte":
                    lo = max(lo, math.ceil(bound))
                elif op == "gt":
                    lo = max(lo, math.floor(bound) + 1)
                elif op == "lte":
                    hi = min(hi, math.floor(bound))
                elif op == "lt":
                    hi = min(hi, math.ceil(bound) - 1)
            lo = max(lo, _INT_MIN)
            hi = min(hi, _INT_MAX)
            if lo > hi:
                flt.add_range(col, 1, 0)
            else:
                flt.add_range(col, int(lo), int(hi))
            return
        if isinstance(want, list):
            host_masks.append(self._host_mask(want))
            return
        if isinstance(want, (bool, np.bool_)):
            want = int(want)
        if isinstance(want, (int, np.integer)) and _INT_MIN <= int(want) <= _INT_MAX:
            flt.add_range(col, int(want), int(want))
        elif isinstance(want, float) and want.is_integer() and _INT_MIN <= want <= _INT_MAX:
            flt.add_range(col, int(want), int(want))
        else:
            flt.add_range(col, 1, 0)  # can never equal an int column value

    # -- host evaluation (rare predicates) --------------------------------------
    def decode(self, row: int):
        if row in self.exotic:
            return self.exotic[row]
        c = int(self.codes[row])
        if c == NULL:
            return None
        if self.kind == "dict":
            return self.values[c]
        if self.kind == "ordered":
            return self.sorted_values[bisect.bisect_left(self.sorted_codes, c)]
        return c

    def _host_mask(self, want) -> np.ndarray:
        """Row mask evaluated on the host over the DISTINCT codes (numpy), for the
        predicates a clause cannot express (None inside a list, exotic values...)."""
        codes = self.codes[: self.n]
        out = np.zeros(self.n, dtype=bool)
        uniq = np.unique(codes)
        allowed = []
        for c in uniq.tolist():
            if c == NULL:
                continue
            if self.kind == "dict":
                v = self.values[c]
            elif self.kind == "ordered":
                v = self.sorted_values[bisect.bisect_left(self.sorted_codes, c)]
            else:
                v = c
            if _row_matches(self.name, v, want):
                allowed.append(c)
        if allowed:
            out = np.isin(codes, np.asarray(allowed, dtype=np.int32))
        null_rows = np.nonzero(codes == NULL)[0]
        for r in null_rows.tolist():
            out[r] = _row_matches(self.name, self.exotic.get(r), want)
        return out


def _row_matches(key: str, have, want) -> bool:
    if isinstance(want, dict):
        return _range_ok(have, want)
    if isinstance(want, list):
        return have in want
    if key == "project_name" and isinstance(want, str) and isinstance(have, str):
        return want.lower() in have.lower()
    return have == want


def pack_bits(mask: np.ndarray) -> np.ndarray:
    n = mask.shape[0]
    padded = np.zeros((n + 31) // 32 * 32, dtype=np.uint8)
    padded[:n] = mask.astype(np.uint8)
    return np.packbits(padded.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1)


class ColumnStore:
    """All device columns of one index + the alive set, with incremental upload."""

    def __init__(self):
        self.codecs = [ColumnCodec(n, k) for n, k in DEVICE_COLUMNS]
        self.n = 0
        self.uploaded = 0  # rows already on the device

    def reset(self) -> None:
        self.__init__()

    def append_rows(self, metas: Iterable[Dict[str, Any]]) -> None:
        metas = list(metas)
        for codec in self.codecs:
            codec.append([m.get(codec.name) for m in metas])
        self.n += len(metas)

    def append_columns(self, n: int, columns: Dict[str, Sequence[Any]]) -> None:
        """Column-wise append of n rows (index reload: one SELECT, no per-row dict); a column that is
        missing from `columns` is NULL for these rows."""
        for codec in self.codecs:
            vals = columns.get(codec.name)
            codec.append(list(vals) if vals is not None else [None] * n)
        self.n += n

    def set_row(self, row: int, meta: Dict[str, Any], index: Optional[_native.Index]) -> None:
        for ci, codec in enumerate(self.codecs):
            codec.set_row(row, meta.get(codec.name))
            if index is not None and row < self.uploaded and not codec.needs_full_upload:
                index.set_column(ci, codec.codes[row:row + 1], start=row)

    def sync(self, index: _native.Index) -> None:
        """Upload rows the device has not seen (and re-encoded columns)."""
        n = min(self.n, index.ntotal)
        for ci, codec in enumerate(self.codecs):
            if codec.needs_full_upload and n:
                index.set_column(ci, codec.codes[:n], start=0)
                codec.needs_full_upload = False
            elif n > self.uploaded:
                index.set_column(ci, codec.codes[self.uploaded:n], start=self.uploaded)
        self.uploaded = n

    def compile(self, filters: Optional[Dict[str, Any]]) -> Optional[_native.Filter]:
        """Reference filter dict -> native Filter (None when nothing restricts)."""
        if not filters:
            return None
        flt = _native.Filter()
        host_masks: List[np.ndarray] = []
        for key, want in filters.items():
            if key in COLUMN_INDEX:
                self.codecs[COLUMN_INDEX[key]].compile(want, flt, host_masks)
            elif key in TABLE_COLUMNS:
                # id / text / metadata / faiss_id / created_at / updated_at: not mirrored on the
                # device; the storage layer resolves these through SQLite into a row mask.
                raise KeyError(key)
            # other keys are ignored, as in the reference (:513-514)
        if host_masks:
            m = host_masks[0]
            for other in host_masks[1:]:
                m = m & other
            flt.set_row_mask(pack_bits(m))
        if flt.n_clauses == 0 and flt.row_mask is None:
            return None
        return flt
