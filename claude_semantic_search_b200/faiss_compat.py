"""Inner seam: the slice of the `faiss` Python API the reference calls
(src/storage.py:252-299,306,358-359,436,879-895,913; src/gpu_utils.py:108-139),
served by libcss_b200.so.  Installing this module as `sys.modules["faiss"]` lets
the reference's own src/storage.py run unmodified on the B200 path
(see INTEGRATION.md).

Only flat indexes exist here: the reference's CLI can only ever build
index_type="flat" (src/cli.py:49-64); IVF/HNSW raise NotImplementedError.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np

from . import _native

__version__ = "1.11.0+css_b200"
METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


class Index:
    """faiss.Index look-alike backed by a css_index handle."""

    metric_type = METRIC_INNER_PRODUCT

    def __init__(self, d: int, metric: int, device: int = 0, devices=None):
        self.d = int(d)
        self.metric_type = metric
        self.is_trained = True
        self._native = _native.Index(self.d, metric, device, devices=devices)

    @property
    def ntotal(self) -> int:
        return self._native.ntotal

    def add(self, x) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise AssertionError(f"add expects [n, {self.d}] float32")
        self._native.add(x, normalize=False)

    def search(self, x, k: int) -> Tuple[np.ndarray, np.ndarray]:
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise AssertionError(f"search expects [nq, {self.d}] float32")
        return self._native.search(x, int(k))   # k > CSS_MAX_K: repeated exact passes (see _native.Index)

    def reset(self) -> None:
        self._native.reset()

    def reconstruct_n(self, i0: int, n: int) -> np.ndarray:
        return self._native.get_rows(i0, n)

    def reconstruct(self, i: int) -> np.ndarray:
        return self._native.get_rows(i, 1)[0]


class IndexFlat(Index):
    pass


class IndexFlatIP(IndexFlat):
    def __init__(self, d: int, device: int = 0, devices=None):
        super().__init__(d, METRIC_INNER_PRODUCT, device, devices)


class IndexFlatL2(IndexFlat):
    def __init__(self, d: int, device: int = 0, devices=None):
        super().__init__(d, METRIC_L2, device, devices)


class IndexIVFFlat(Index):
    def __init__(self, *a, **kw):
        raise NotImplementedError("only flat indexes are served by the B200 path (the reference's "
                                  "CLI never builds anything else)")


class IndexHNSWFlat(IndexIVFFlat):
    pass


def write_index(index: Index, path: str) -> None:
    index._native.save(path)


def read_index(path: str, device: int = 0, devices=None) -> Index:
    # peek at fourcc + d to build the right class (faiss IndexFlat header)
    with open(path, "rb") as fh:
        head = fh.read(8)
    if len(head) < 8 or head[:4] not in (b"IxFI", b"IxF2"):
        raise RuntimeError(f"{path}: not a flat faiss index")
    d = int(np.frombuffer(head[4:8], dtype=np.int32)[0])
    idx = IndexFlatIP(d, device, devices) if head[:4] == b"IxFI" else IndexFlatL2(d, device, devices)
    idx._native.load(path)
    return idx


# --- the GPU-resource shims the reference probes (src/storage.py:269-299) -------
class StandardGpuResources:
    def __init__(self):
        _native.device_count()  # raises when there is no sm_100 device


def get_num_gpus() -> int:
    try:
        return _native.device_count()
    except _native.NativeError:
        return 0


def index_cpu_to_gpu(res, device: int, index: Index) -> Index:
    return index  # already device-resident


def index_gpu_to_cpu(index: Index) -> Index:
    return index
