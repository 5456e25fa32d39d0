"""oracle/two_phase_proof.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Plain-Python restatement of the decision logic of two_phase_finish
(claude_semantic_search_b200/csrc/index_kernels.cuh): given, per scan block, the kp best rows of that block by
SHADOW score (sorted, ties by id) together with their EXACT scores, and eps >= |shadow - exact| for every row of the
corpus, either return the exact top-k or report "unproven".  No reference counterpart: faiss scans in fp32
(src/storage.py:436); the contract is that a PROVEN answer equals the brute-force top-k by exact score
(score descending, id ascending).  tests/test_two_phase_proof_cpu.py hammers that on small random and adversarial
instances where lists are short enough for every branch of the proof to matter.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

Entry = Tuple[float, float, int]   # (shadow score, exact score, row id)


def block_lists(shadow: Sequence[float], exact: Sequence[float], blocks: int, kp: int, owner: Sequence[int]) -> List[List[Entry]]:
    """Phase 1: block b keeps the kp best of its rows by (shadow desc, id asc).  Rows whose shadow score ties with the
    kp-th may be dropped (WarpBufTop32 drops ties): modelled by keeping exactly kp."""
    lists: List[List[Entry]] = [[] for _ in range(blocks)]
    for i, b in enumerate(owner):
        lists[b].append((shadow[i], exact[i], i))
    for b in range(blocks):
        lists[b].sort(key=lambda e: (-e[0], e[2]))
        lists[b] = lists[b][:kp]
    return lists


def kth_best(entries: Sequence[Tuple[float, int]], k: int) -> float:
    """k-th best key by (key desc, id asc); -inf with fewer than k entries."""
    s = sorted(entries, key=lambda e: (-e[0], e[1]))
    return s[k - 1][0] if len(s) >= k else -math.inf


def finish(lists: List[List[Entry]], kp: int, k: int, eps: float, cap: int = 2048) -> Optional[List[Tuple[float, int]]]:
    """Phase 2.  Returns [(exact score, id)] best first, or None when the query must be re-run by the fp32 scan."""
    if not (eps < math.inf):
        return None
    blocks = len(lists)
    per = 1 if blocks >= k else k
    heads = [(e[1], e[2]) for lst in lists for e in lst[:per]]           # (A) exact scores of the list heads
    t0 = kth_best(heads, k)
    thr = t0 - eps if t0 > -math.inf else -math.inf
    cand = [(e[1], e[2]) for lst in lists for e in lst if e[0] >= thr]    # (C) shadow score >= thr
    full_last = [lst[-1][0] for lst in lists if len(lst) == kp]
    if len(cand) > cap:
        return None
    cand.sort(key=lambda e: (-e[0], e[1]))                                # (E)
    top = cand[:k]
    t1 = top[k - 1][0] if len(top) >= k else -math.inf                    # (F)
    if full_last and not (max(full_last) < t1 - eps):
        return None
    return top
