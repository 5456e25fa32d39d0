"""ctypes binding of libcss_b200.so (the C ABI declared in include/css_b200.h).

There is no CPU fallback: if the library is missing, or no sm_100 device is
present when a compute entry point is called, a NativeError is raised.
"""
from __future__ import annotations

import ctypes
import threading
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint8, c_uint32, c_void_p
from pathlib import Path
from typing import Optional

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libcss_b200.so"

CSS_OK = 0
CSS_ERR_INVALID = -1
CSS_ERR_NO_DEVICE = -2
CSS_ERR_CUDA = -3
CSS_ERR_OOM = -4
CSS_ERR_IO = -5
CSS_ERR_UNSUPPORTED = -6
CSS_ERR_OVERFLOW = -7

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1
MAX_K = 128
MAX_COLUMNS = 12
MAX_CLAUSES = 16
MAX_RANKS = 8
IPC_HANDLE_BYTES = 64
EXCHANGE_MAX_NQ = 64
NULL_VALUE = -(2 ** 31)
CLAUSE_RANGE = 0
CLAUSE_SET = 1


class NativeError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libcss_b200 error {status}: {message}")
        self.status = status


class NoDeviceError(NativeError):
    pass


class css_clause(ctypes.Structure):
    _fields_ = [
        ("column", c_int32),
        ("kind", c_int32),
        ("lo", c_int32),
        ("hi", c_int32),
        ("set_bits", POINTER(c_uint32)),
        ("set_nbits", c_int32),
        ("reserved", c_int32),
    ]


class css_filter(ctypes.Structure):
    _fields_ = [
        ("n_clauses", c_int32),
        ("ignore_alive", c_int32),
        ("clauses", POINTER(css_clause)),
        ("row_mask", POINTER(c_uint32)),
    ]


class css_mpnet_config(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in ("vocab_size", "hidden_size", "num_layers", "num_heads", "intermediate_size",
                                       "max_position", "rel_buckets", "rel_max_distance", "pad_token_id")] + \
               [("layer_norm_eps", c_float)]


_LAYER_FIELDS = ("q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b", "ln1_w", "ln1_b",
                 "ffn1_w", "ffn1_b", "ffn2_w", "ffn2_b", "ln2_w", "ln2_b")


class css_mpnet_layer(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in _LAYER_FIELDS]


class css_mpnet_weights(ctypes.Structure):
    _fields_ = [("word_emb", c_void_p), ("pos_emb", c_void_p), ("emb_ln_w", c_void_p), ("emb_ln_b", c_void_p),
                ("rel_bias", c_void_p), ("layers", POINTER(css_mpnet_layer))]


# name -> (restype, argtypes); every symbol include/css_b200.h declares
SIGNATURES = {
    "css_abi_version": (c_int, []),
    "css_last_error": (c_char_p, []),
    "css_device_count": (c_int, [POINTER(c_int)]),
    "css_device_info": (c_int, [c_int, POINTER(c_int64)]),
    "css_index_create": (c_int, [c_int, c_int, c_int, POINTER(c_void_p)]),
    "css_index_create_sharded": (c_int, [c_int, c_int, POINTER(c_int), c_int, POINTER(c_void_p)]),
    "css_index_n_devices": (c_int, [c_void_p]),
    "css_index_set_alive_ids": (c_int, [c_void_p, c_void_p, c_int64, c_int]),
    "css_index_scan_stats": (c_int, [c_void_p, POINTER(c_int64)]),
    "css_exchange_create": (c_int, [c_int, c_int, c_int, POINTER(c_void_p), c_void_p]),
    "css_exchange_connect": (c_int, [c_void_p, c_void_p]),
    "css_exchange_destroy": (c_int, [c_void_p]),
    "css_index_search_exchange_device": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int64,
                                                 c_void_p, c_void_p, c_void_p]),
    "css_index_destroy": (c_int, [c_void_p]),
    "css_index_dim": (c_int, [c_void_p]),
    "css_index_metric": (c_int, [c_void_p]),
    "css_index_ntotal": (c_int64, [c_void_p]),
    "css_index_capacity": (c_int64, [c_void_p]),
    "css_index_reserve": (c_int, [c_void_p, c_int64]),
    "css_index_reset": (c_int, [c_void_p]),
    "css_index_add": (c_int, [c_void_p, c_void_p, c_int64, c_int, POINTER(c_int64)]),
    "css_index_add_device": (c_int, [c_void_p, c_void_p, c_int64, c_int, POINTER(c_int64), c_void_p]),
    "css_index_compact": (c_int, [c_void_p, c_void_p, c_int64]),
    "css_index_get_rows": (c_int, [c_void_p, c_int64, c_int64, c_void_p]),
    "css_index_set_column": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int64]),
    "css_index_set_alive": (c_int, [c_void_p, c_void_p, c_int64, c_int64]),
    "css_index_filter_mask": (c_int, [c_void_p, POINTER(css_filter), c_void_p, POINTER(c_int64)]),
    "css_index_filter_mask_device": (c_int, [c_void_p, POINTER(css_filter), POINTER(c_void_p), POINTER(c_int64), c_void_p]),
    "css_index_search": (c_int, [c_void_p, c_void_p, c_int, c_int, POINTER(css_filter), c_void_p, c_void_p]),
    "css_index_search_device": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "css_topk_merge_device": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "css_topk_merge_strided_device": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_int,
                                              c_void_p, c_void_p, c_void_p]),
    "css_index_save": (c_int, [c_void_p, c_char_p]),
    "css_index_load": (c_int, [c_void_p, c_char_p]),
    "css_kernel_launch_count": (c_int64, []),
    "css_set_option": (c_int, [c_char_p, c_int]),
    "css_mpnet_relative_bucket": (c_int, [c_int, c_int, c_int]),
    "css_encoder_create": (c_int, [POINTER(css_mpnet_config), POINTER(css_mpnet_weights), c_int, c_int64,
                                   POINTER(c_void_p)]),
    "css_encoder_destroy": (c_int, [c_void_p]),
    "css_encoder_dim": (c_int, [c_void_p]),
    "css_encoder_max_tokens": (c_int64, [c_void_p]),
    "css_encoder_max_seq_len": (c_int, [c_void_p]),
    "css_encoder_encode": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int, c_void_p]),
    "css_index_search_exchange": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p]),
    "css_debug_scan_bf16": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "css_debug_scan_int8": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "css_debug_scan_trace": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "css_tokenizer_create": (c_int, [ctypes.c_char_p, c_int, POINTER(c_void_p)]),
    "css_tokenizer_destroy": (c_int, [c_void_p]),
    "css_tokenizer_vocab_size": (c_int, [c_void_p]),
    "css_tokenizer_add_special": (c_int, [c_void_p, ctypes.c_char_p, c_int32]),
    "css_tokenizer_encode_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p,
                                           c_void_p, c_int32]),
    "css_debug_gemm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "css_debug_gemm_resid_ln": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                        ctypes.c_float, c_int, c_int, c_void_p]),
    "css_debug_attention": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "css_encoder_encode_device": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int, c_void_p,
                                          c_void_p]),
}

_lib: Optional[ctypes.CDLL] = None


def register_signatures(extra: dict) -> None:
    """Other modules (encoder) add their entry points here before load()."""
    SIGNATURES.update(extra)


def load() -> ctypes.CDLL:
    """Load libcss_b200.so; raises NativeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("CSS_B200_LIB", LIB_PATH))
    if not path.exists():
        raise NativeError(CSS_ERR_UNSUPPORTED,
                          f"{path} not found: build it with `python -m claude_semantic_search_b200.build` "
                          "(there is no CPU fallback)")
    lib = ctypes.CDLL(str(path))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.css_abi_version() != 1:
        raise NativeError(CSS_ERR_UNSUPPORTED, f"ABI version {lib.css_abi_version()} != 1")
    _lib = lib
    return lib


def check(status: int) -> None:
    if status == CSS_OK:
        return
    msg = load().css_last_error().decode("utf-8", "replace")
    if status == CSS_ERR_NO_DEVICE:
        raise NoDeviceError(status, msg)
    raise NativeError(status, msg)


def device_count() -> int:
    n = c_int(0)
    check(load().css_device_count(ctypes.byref(n)))
    return n.value


def has_device() -> bool:
    try:
        return device_count() > 0
    except NativeError:
        return False


def device_info(device: int = 0) -> dict:
    info = (c_int64 * 5)()
    check(load().css_device_info(device, info))
    return {"sm_count": info[0], "hbm_total": info[1], "hbm_free": info[2], "cc": (info[3], info[4])}


def set_option(name: str, value: int) -> None:
    """Process-wide switch of libcss_b200 (css_set_option): scan_bf16, scan_int8, scan_interleave, scan_list, scan_adaptive."""
    check(load().css_set_option(name.encode(), int(value)))


def kernel_launch_count() -> int:
    return int(load().css_kernel_launch_count())


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


class Filter:
    """Host-side builder of a css_filter; keeps the numpy buffers alive."""

    def __init__(self, ignore_alive: bool = False):
        self._clauses = []
        self._keep = []
        self.ignore_alive = ignore_alive
        self.row_mask: Optional[np.ndarray] = None

    def add_range(self, column: int, lo: int, hi: int) -> "Filter":
        self._clauses.append((column, CLAUSE_RANGE, int(lo), int(hi), None, 0))
        return self

    def add_set(self, column: int, allowed_ids, universe: int) -> "Filter":
        """Rows whose column value is one of allowed_ids (0 <= id < universe)."""
        bits = np.zeros((max(universe, 1) + 31) // 32, dtype=np.uint32)
        for v in allowed_ids:
            v = int(v)
            if 0 <= v < universe:
                bits[v >> 5] |= np.uint32(1 << (v & 31))
        self._clauses.append((column, CLAUSE_SET, 0, 0, bits, int(universe)))
        return self

    def set_row_mask(self, mask_words: np.ndarray) -> "Filter":
        self.row_mask = np.ascontiguousarray(mask_words, dtype=np.uint32)
        return self

    @property
    def n_clauses(self) -> int:
        return len(self._clauses)

    def build(self) -> css_filter:
        n = len(self._clauses)
        if n > MAX_CLAUSES:
            raise NativeError(CSS_ERR_INVALID, f"{n} clauses > {MAX_CLAUSES}")
        arr = (css_clause * max(n, 1))()
        for i, (col, kind, lo, hi, bits, nbits) in enumerate(self._clauses):
            arr[i].column, arr[i].kind, arr[i].lo, arr[i].hi = col, kind, lo, hi
            arr[i].set_nbits = nbits
            if bits is not None:
                arr[i].set_bits = bits.ctypes.data_as(POINTER(c_uint32))
        f = css_filter()
        f.n_clauses = n
        f.ignore_alive = 1 if self.ignore_alive else 0
        f.clauses = ctypes.cast(arr, POINTER(css_clause))
        if self.row_mask is not None:
            f.row_mask = self.row_mask.ctypes.data_as(POINTER(c_uint32))
        self._keep = [arr]
        return f


class Exchange:
    """css_exchange: the in-kernel result exchange between the row shards of one search
    (one per rank / GPU).  `handle` is the 64-byte CUDA IPC handle the ranks all-gather."""

    def __init__(self, device: int, n_ranks: int, rank: int):
        self._lib = load()
        self._h = c_void_p()
        buf = (ctypes.c_ubyte * IPC_HANDLE_BYTES)()
        check(self._lib.css_exchange_create(device, n_ranks, rank, ctypes.byref(self._h), buf))
        self.handle = bytes(buf)
        self.n_ranks = n_ranks
        self.rank = rank
        self.device = device

    def connect(self, handles) -> None:
        """handles: the n_ranks IPC handles in rank order (bytes each)."""
        blob = b"".join(handles)
        if len(blob) != self.n_ranks * IPC_HANDLE_BYTES:
            raise ValueError("expected n_ranks handles of 64 bytes")
        buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
        check(self._lib.css_exchange_connect(self._h, buf))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.css_exchange_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Index:
    """Thin OO wrapper over the css_index_* entry points.  `devices` (a list of device ordinals)
    builds ONE index row-sharded over several GPUs of this process (css_index_create_sharded)."""

    def __init__(self, dim: int, metric: int = METRIC_INNER_PRODUCT, device: int = 0, devices=None):
        self._lib = load()
        self._h = c_void_p()
        if devices is not None and len(devices) > 0:
            devs = (c_int * len(devices))(*[int(d) for d in devices])
            check(self._lib.css_index_create_sharded(dim, metric, devs, len(devices), ctypes.byref(self._h)))
            device = int(devices[0])
        else:
            check(self._lib.css_index_create(dim, metric, device, ctypes.byref(self._h)))
        self.dim = dim
        self.metric = metric
        self.device = device
        self.devices = [int(d) for d in devices] if devices else [device]
        self._tl = threading.local()   # per-thread result buffers of single-query searches

    # -- lifecycle --------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.css_index_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> c_void_p:
        return self._h

    @property
    def ntotal(self) -> int:
        return int(self._lib.css_index_ntotal(self._h))

    @property
    def capacity(self) -> int:
        return int(self._lib.css_index_capacity(self._h))

    def reserve(self, capacity: int) -> None:
        check(self._lib.css_index_reserve(self._h, capacity))

    def reset(self) -> None:
        check(self._lib.css_index_reset(self._h))

    # -- data ---------------------------------------------------------------
    def add(self, x, normalize: bool = False) -> int:
        x = _f32(x)
        if x.ndim != 2 or x.shape[1] != self.dim:
            raise ValueError(f"expected [n, {self.dim}] float32, got {x.shape}")
        first = c_int64(0)
        check(self._lib.css_index_add(self._h, x.ctypes.data, x.shape[0], 1 if normalize else 0,
                                      ctypes.byref(first)))
        return first.value

    def add_device(self, x_dev_ptr: int, n: int, normalize: bool = False, stream: int = 0) -> int:
        first = c_int64(0)
        check(self._lib.css_index_add_device(self._h, c_void_p(x_dev_ptr), n, 1 if normalize else 0,
                                             ctypes.byref(first), c_void_p(stream)))
        return first.value

    def compact(self, keep_ids) -> None:
        """Keep exactly the rows `keep_ids` (strictly ascending), renumbered 0..len-1, on the device."""
        k = np.ascontiguousarray(keep_ids, dtype=np.int64)
        check(self._lib.css_index_compact(self._h, k.ctypes.data, k.shape[0]))

    def get_rows(self, start: int, n: int) -> np.ndarray:
        out = np.empty((n, self.dim), np.float32)
        check(self._lib.css_index_get_rows(self._h, start, n, out.ctypes.data))
        return out

    def set_column(self, column: int, values, start: int = 0) -> None:
        v = np.ascontiguousarray(values, dtype=np.int32)
        check(self._lib.css_index_set_column(self._h, column, v.ctypes.data, start, v.shape[0]))

    def set_alive(self, alive, start: int = 0) -> None:
        a = np.ascontiguousarray(alive, dtype=np.uint8)
        check(self._lib.css_index_set_alive(self._h, a.ctypes.data, start, a.shape[0]))

    def set_alive_ids(self, ids, alive: bool) -> None:
        """Mark the listed rows alive / dead in one device call."""
        a = np.ascontiguousarray(ids, dtype=np.int64)
        check(self._lib.css_index_set_alive_ids(self._h, a.ctypes.data, a.shape[0], 1 if alive else 0))

    def scan_stats(self) -> dict:
        out = (c_int64 * 6)()
        check(self._lib.css_index_scan_stats(self._h, out))
        return {"two_phase_queries": int(out[0]), "unproven_queries": int(out[1]), "bypassed": bool(out[2]),
                "max_bf16_error_norm": out[3] * 1e-9,
                "max_int8_error_norm": float("inf") if out[4] == 2 ** 63 - 1 else out[4] * 1e-9,
                "last_tier": int(out[5])}

    # -- filter / search ----------------------------------------------------
    def filter_mask(self, flt: Optional[Filter]) -> tuple:
        n = self.ntotal
        words = np.zeros((n + 31) // 32, dtype=np.uint32)
        n_pass = c_int64(0)
        cf = flt.build() if flt is not None else None
        check(self._lib.css_index_filter_mask(self._h, ctypes.byref(cf) if cf is not None else None,
                                              words.ctypes.data, ctypes.byref(n_pass)))
        return words, n_pass.value

    def search(self, q, k: int, flt: Optional[Filter] = None) -> tuple:
        q = _f32(q)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        if q.shape[1] != self.dim:
            raise ValueError(f"expected [nq, {self.dim}] float32, got {q.shape}")
        nq = q.shape[0]
        if k > MAX_K:
            return self._search_large_k(q, k, flt)
        if nq == 1 and flt is None:
            # the single-query call is ~150 us on the device side: keep the result buffers and their addresses per
            # thread instead of paying numpy.empty + ndarray.ctypes (1 us each) on every call
            tl = self._tl
            ent = tl.__dict__.get(k)
            if ent is None:
                D0, I0 = np.empty((1, k), np.float32), np.empty((1, k), np.int64)
                ent = tl.__dict__[k] = (D0, I0, D0.ctypes.data, I0.ctypes.data)
            check(self._lib.css_index_search(self._h, q.ctypes.data, 1, k, None, ent[2], ent[3]))
            return ent[0].copy(), ent[1].copy()
        D = np.empty((nq, k), np.float32)
        I = np.empty((nq, k), np.int64)
        cf = flt.build() if flt is not None else None
        check(self._lib.css_index_search(self._h, q.ctypes.data, nq, k,
                                         ctypes.byref(cf) if cf is not None else None,
                                         D.ctypes.data, I.ctypes.data))
        return D, I

    def _search_large_k(self, q: np.ndarray, k: int, flt: Optional[Filter]) -> tuple:
        """k > CSS_MAX_K (the reference's max_results is a free field, src/storage.py:69,432): the exact
        top-k is the top-128, then the top-128 of the remaining rows, ... -- each pass is one exact device
        search with the rows found so far cleared from the row mask."""
        nq = q.shape[0]
        fill = -np.finfo(np.float32).max if self.metric == METRIC_INNER_PRODUCT else np.finfo(np.float32).max
        D = np.full((nq, k), fill, np.float32)
        I = np.full((nq, k), -1, np.int64)
        base, _ = self.filter_mask(flt)          # honours the alive bits unless flt.ignore_alive
        for qi in range(nq):
            words = base.copy()
            got = 0
            while got < k:
                kk = min(MAX_K, k - got)
                f = Filter(ignore_alive=True).set_row_mask(words)
                d, i = self.search(q[qi:qi + 1], kk, f)
                ok = i[0] >= 0
                n = int(ok.sum())
                D[qi, got:got + n] = d[0][ok]
                I[qi, got:got + n] = i[0][ok]
                got += n
                if n < kk:
                    break
                ids = i[0][ok]
                np.bitwise_and.at(words, ids >> 5, ~(np.uint32(1) << (ids & 31).astype(np.uint32)))
        return D, I

    def filter_mask_device(self, flt: Optional[Filter], stream: int = 0, want_count: bool = False):
        ptr = c_void_p()
        n_pass = c_int64(0)
        cf = flt.build() if flt is not None else None
        check(self._lib.css_index_filter_mask_device(
            self._h, ctypes.byref(cf) if cf is not None else None, ctypes.byref(ptr),
            ctypes.byref(n_pass) if want_count else None, c_void_p(stream)))
        return (ptr.value or 0), (n_pass.value if want_count else None)

    def search_device(self, q_ptr: int, nq: int, k: int, D_ptr: int, I_ptr: int, mask_ptr: int = 0,
                      id_offset: int = 0, stream: int = 0) -> None:
        check(self._lib.css_index_search_device(self._h, c_void_p(q_ptr), nq, k,
                                                c_void_p(mask_ptr) if mask_ptr else None, id_offset,
                                                c_void_p(D_ptr), c_void_p(I_ptr), c_void_p(stream)))

    def search_exchange_device(self, ex: Exchange, q_ptr: int, nq: int, k: int, D_ptr: int, I_ptr: int,
                               mask_ptr: int = 0, id_offset: int = 0, stream: int = 0) -> None:
        """Collective over the ranks of `ex`: local scan + in-kernel NVLink exchange + merge; D/I hold
        the merged global top-k on every rank when the stream reaches this point."""
        check(self._lib.css_index_search_exchange_device(self._h, ex._h, c_void_p(q_ptr), nq, k,
                                                         c_void_p(mask_ptr) if mask_ptr else None, id_offset,
                                                         c_void_p(D_ptr), c_void_p(I_ptr), c_void_p(stream)))

    def search_exchange_host(self, ex: "Exchange", q: np.ndarray, k: int, id_offset: int = 0, mask_ptr: int = 0) -> tuple:
        """One query in host memory, searched collectively over the ranks of `ex` (css_index_search_exchange)."""
        q = _f32(q).reshape(-1)
        D = np.empty((1, k), np.float32)
        I = np.empty((1, k), np.int64)
        check(self._lib.css_index_search_exchange(self._h, ex._h, q.ctypes.data, k, c_void_p(mask_ptr), id_offset,
                                                  D.ctypes.data, I.ctypes.data))
        return D, I

    def debug_scan_bf16(self, q_ptr: int, nq: int, stream: int = 0) -> None:
        """Phase 1 alone of the two-phase scan (benchmark hook, css_debug_scan_bf16)."""
        check(self._lib.css_debug_scan_bf16(self._h, c_void_p(q_ptr), nq, c_void_p(stream)))

    def debug_scan_int8(self, q_ptr: int, nq: int, stream: int = 0) -> None:
        """The int8 shadow sweep alone (first tier of the two-phase scan; benchmark hook, css_debug_scan_int8)."""
        check(self._lib.css_debug_scan_int8(self._h, c_void_p(q_ptr), nq, c_void_p(stream)))

    # -- persistence --------------------------------------------------------
    def debug_scan_trace(self, q_ptr: int, k: int = 10, stream: int = 0, blocks: int = 160) -> np.ndarray:
        """Timeline (ns) of one batch-1 scan, [blocks + 1, 8] (css_debug_scan_trace); all-zero rows are unused."""
        out = np.zeros((blocks + 1) * 8, dtype=np.int64)
        check(self._lib.css_debug_scan_trace(self._h, c_void_p(q_ptr), k, out.ctypes.data, out.shape[0], c_void_p(stream)))
        return out.reshape(-1, 8)

    def save(self, path) -> None:
        check(self._lib.css_index_save(self._h, str(path).encode()))

    def load(self, path) -> None:
        check(self._lib.css_index_load(self._h, str(path).encode()))
        self.metric = int(self._lib.css_index_metric(self._h))


def topk_merge_strided_device(D_in_ptr: int, d_stride: int, I_in_ptr: int, i_stride: int, n_lists: int, nq: int, k: int,
                              metric: int, D_out_ptr: int, I_out_ptr: int, stream: int = 0) -> None:
    check(load().css_topk_merge_strided_device(c_void_p(D_in_ptr), d_stride, c_void_p(I_in_ptr), i_stride, n_lists, nq, k,
                                               metric, c_void_p(D_out_ptr), c_void_p(I_out_ptr), c_void_p(stream)))


def topk_merge_device(D_in_ptr: int, I_in_ptr: int, n_lists: int, nq: int, k: int, metric: int,
                      D_out_ptr: int, I_out_ptr: int, stream: int = 0) -> None:
    check(load().css_topk_merge_device(c_void_p(D_in_ptr), c_void_p(I_in_ptr), n_lists, nq, k, metric,
                                       c_void_p(D_out_ptr), c_void_p(I_out_ptr), c_void_p(stream)))
