"""The shuffle networks of the shadow sweeps' selection (oracle/warp_select.py restates them lane by lane): the
buffer sort, the list/buffer merge and the block tournament keep exactly the 32 best entries, in order."""
import math

import numpy as np

from oracle import warp_select as ws


def _rand_entries(rng, n, ties):
    keys = rng.standard_normal(n).astype(np.float32)
    if ties:
        keys = np.round(keys * 2) / 2
    ids = rng.permutation(10_000)[:n]
    return [(float(k), int(i)) for k, i in zip(keys, ids)]


def _best32(entries):
    s = sorted(entries, key=lambda e: (-e[0], e[1]))[:32]
    return s + [ws.EMPTY] * (32 - len(s))


def test_sort_merge_compact():
    rng = np.random.default_rng(0)
    for trial in range(400):
        ties = trial % 2 == 1
        buf = _rand_entries(rng, 32, ties)
        assert ws.sort_desc(buf) == _best32(buf)
        n_list = int(rng.integers(0, 33))
        lst = _best32(_rand_entries(rng, n_list, ties))
        nb = int(rng.integers(0, 33))
        stale = _rand_entries(rng, 32, ties)                      # lanes >= nb hold stale entries: must be ignored
        got = ws.compact(lst, stale, nb)
        want = _best32([e for e in lst if e != ws.EMPTY] + stale[:nb])
        assert got == want, (trial, n_list, nb)


def test_tournament_of_sixteen_lists():
    rng = np.random.default_rng(1)
    for trial in range(60):
        lists = [_best32(_rand_entries(rng, int(rng.integers(0, 33)), trial % 2 == 1)) for _ in range(16)]
        cur = list(lists)
        stride = 1
        while stride < 16:
            for w in range(0, 16, 2 * stride):
                cur[w] = ws.merge_lists(cur[w], cur[w + stride])
            stride <<= 1
        want = _best32([e for lst in lists for e in lst if e != ws.EMPTY])
        assert cur[0] == want
        assert all(math.isinf(e[0]) or e[1] != ws.EMPTY[1] for e in cur[0])
