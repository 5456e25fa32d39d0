"""UnitFeed (the unit dealer of the dense int8 sweep, restated in oracle/unit_feed.py): every unit is dealt exactly
once, for any corpus size, grid size and ticket order."""
import numpy as np
import pytest

from oracle import unit_feed as uf


def _pickers(rng):
    yield lambda live, step: step                                   # round robin over the live warps
    yield lambda live, step: -1 - step                              # the other way round
    yield lambda live, step: int(rng.integers(0, 1 << 30))          # random interleaving
    yield lambda live, step: 0                                      # one warp runs dry before the next starts


@pytest.mark.parametrize("nwarps", [1, 2, 16, 37, 2368])
def test_every_unit_dealt_exactly_once(nwarps):
    rng = np.random.default_rng(nwarps)
    sizes = {0, 1, 2, 3, nwarps - 1, nwarps, nwarps + 1, 2 * nwarps, 3 * nwarps + 5, 10 * nwarps, 10 * nwarps + 3,
             53 * nwarps - 1, 125_000 if nwarps == 2368 else 17 * nwarps + 9}
    sizes |= {int(x) for x in rng.integers(0, 60 * nwarps, size=6)}
    for units in sorted(u for u in sizes if 0 <= u <= 130_000):
        for pick in _pickers(rng):
            dealt = uf.deal(units, nwarps, pick)
            assert sorted(dealt) == list(range(units)), (units, nwarps)
            assert set(dealt.values()) <= {1}, (units, nwarps)
            if nwarps == 2368 and units > 20_000:
                break                                                       # one order is enough at full size


def test_static_share_and_ticket_sizes_at_the_headline_shape():
    f = uf.Feed(125_000, 2368, 0, [0])
    assert f.stat_rounds == 52 - (52 // 8 + 2)                              # 44 of 52 whole rounds are static
    assert f.a_units % 4 == 0 and f.stat_units + f.a_units <= 125_000
    assert 125_000 - f.stat_units - f.a_units >= 2 * 2368                   # the last two rounds go out one unit at a time
