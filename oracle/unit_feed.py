"""oracle/unit_feed.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Plain-Python restatement of UnitFeed (claude_semantic_search_b200/csrc/index_kernels.cuh): how the dense int8 sweep
deals the corpus' 8-row units to the warps of the grid -- static block-cyclic rounds first, then tickets drawn from
a shared cursor (four units per ticket, one unit for the tickets of the last two rounds).  No reference counterpart
(faiss scans rows in order, src/storage.py:436); the contract is that every unit of [0, units) is dealt exactly
once whatever the order in which the warps draw their tickets.
"""
from __future__ import annotations

from typing import Dict, List


class Feed:
    def __init__(self, units: int, nwarps: int, gw: int, cursor: List[int]):
        self.units, self.nwarps, self.gw, self.cursor = units, nwarps, gw, cursor
        per_warp = units // nwarps
        dyn_rounds = per_warp // 8 + 2
        self.stat_rounds = per_warp - dyn_rounds if per_warp > dyn_rounds else 0
        self.stat_units = self.stat_rounds * nwarps
        dyn_units = units - self.stat_units
        b_units = min(dyn_units, 2 * nwarps)
        self.a_tickets = (dyn_units - b_units) // 4
        self.a_units = self.a_tickets * 4
        self.r = 0
        self.cur = 0
        self.left = 0
        self.done = False
        self.raw = self._draw()          # fetched one ticket ahead

    def _draw(self) -> int:
        t = self.cursor[0]
        self.cursor[0] += 1
        return t

    def next(self) -> int:
        if self.r < self.stat_rounds:
            u = self.r * self.nwarps + self.gw
            self.r += 1
            return u
        if self.left > 0:
            self.left -= 1
            u = self.cur
            self.cur += 1
            return u
        if self.done:
            return -1
        t = self.raw
        base = self.stat_units + 4 * t if t < self.a_tickets else self.stat_units + self.a_units + (t - self.a_tickets)
        if base >= self.units:
            self.done = True
            return -1
        self.raw = self._draw()
        cnt = 4 if t < self.a_tickets else 1
        self.left = min(cnt, self.units - base) - 1
        self.cur = base + 1
        return base


def deal(units: int, nwarps: int, pick) -> Dict[int, int]:
    """Run all warps to exhaustion; pick(live, step) -> position in the list of live warps decides who steps next
    (the hardware's interleaving is arbitrary).  Returns {unit: times dealt}."""
    cursor = [0]
    feeds = [Feed(units, nwarps, gw, cursor) for gw in range(nwarps)]
    live = list(range(nwarps))
    dealt: Dict[int, int] = {}
    step = 0
    while live:
        pos = pick(live, step) % len(live)
        step += 1
        u = feeds[live[pos]].next()
        if u < 0:
            live[pos] = live[-1]
            live.pop()
        else:
            dealt[u] = dealt.get(u, 0) + 1
    return dealt
