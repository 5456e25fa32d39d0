"""oracle/warp_select.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Lane-level restatement of WarpBufTop32::compact and of the block tournament merge
(claude_semantic_search_b200/csrc/index_kernels.cuh): 32 lanes, one (key, id) entry each, compare-exchange steps
with the lane's xor partner exactly as the shuffles do them.  Contract: after compact() the list holds the 32 best
of (old list + pending buffer) in descending (key, id-ascending) order; merge_lists() of two sorted lists yields the
sorted 32 best of both.
"""
from __future__ import annotations

import math
from typing import List, Tuple

Entry = Tuple[float, int]
EMPTY: Entry = (-math.inf, 2 ** 31 - 1)


def better(a: Entry, b: Entry) -> bool:
    return a[0] > b[0] or (a[0] == b[0] and a[1] < b[1])


def _exchange(v: List[Entry], o: int, keep_better) -> List[Entry]:
    """Every lane looks at lane ^ o and takes the partner's entry iff (partner better) == keep_better(lane)."""
    out = list(v)
    for lane in range(32):
        other = v[lane ^ o]
        if better(other, v[lane]) == keep_better(lane):
            out[lane] = other
    return out


def sort_desc(v: List[Entry]) -> List[Entry]:
    """The bitonic network of compact(): best entry ends in lane 0."""
    size = 2
    while size <= 32:
        o = size >> 1
        while o > 0:
            v = _exchange(v, o, lambda lane, size=size, o=o: ((lane & o) == 0) == ((lane & size) == 0 or size == 32))
            o >>= 1
        size <<= 1
    return v


def merge_lists(a: List[Entry], b: List[Entry]) -> List[Entry]:
    """a, b sorted descending: lane i takes better(a[i], b[31 - i]) (a bitonic sequence holding the 32 best of both),
    five exchange steps sort it."""
    v = [b[31 - i] if better(b[31 - i], a[i]) else a[i] for i in range(32)]
    o = 16
    while o > 0:
        v = _exchange(v, o, lambda lane, o=o: (lane & o) == 0)
        o >>= 1
    return v


def compact(lst: List[Entry], buf: List[Entry], nb: int) -> List[Entry]:
    buf = [buf[i] if i < nb else EMPTY for i in range(32)]
    return merge_lists(lst, sort_desc(buf))
