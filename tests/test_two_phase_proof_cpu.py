"""The proof logic of the two-phase scan's finish (oracle/two_phase_proof.py restates two_phase_finish): whenever it
answers, the answer is the brute-force top-k by exact score -- on instances small enough (few blocks, short lists)
that every branch matters, with shadow errors drawn at random and adversarially (+-eps)."""
import math

import numpy as np
import pytest

from oracle import two_phase_proof as tp


def brute(exact, k):
    order = sorted(range(len(exact)), key=lambda i: (-exact[i], i))[:k]
    return [(exact[i], i) for i in order]


def run(rng, n, blocks, kp, k, eps, mode, spread):
    exact = (rng.standard_normal(n) * spread).astype(np.float32).astype(float)
    if mode == "ties":
        exact = np.round(exact / (eps if eps > 0 else 1.0)).astype(float) * (eps if eps > 0 else 1.0)   # many exact ties
    if mode == "adversarial":
        rank = np.argsort(np.argsort(-exact))
        err = np.where(rank < k, -eps, eps)              # the true top-k look worse, everything else looks better
    elif mode == "adversarial2":
        err = rng.choice([-eps, eps], size=n)
    else:
        err = rng.uniform(-eps, eps, size=n)
    shadow = exact + err
    owner = rng.integers(0, blocks, size=n) if mode != "clustered" else (np.arange(n) * blocks // n)
    if mode == "clustered":
        exact = np.sort(exact)[::-1].copy()              # the best rows all sit in the first block
        shadow = exact + err
    lists = tp.block_lists(list(shadow), list(exact), blocks, kp, list(owner))
    got = tp.finish(lists, kp, k, eps)
    if got is None:
        return False
    want = brute(list(exact), k)
    assert got == want, (mode, n, blocks, kp, k, eps, got, want)
    return True


@pytest.mark.parametrize("mode", ["random", "adversarial", "adversarial2", "ties", "clustered"])
def test_proven_answers_are_exact(mode):
    rng = np.random.default_rng(hash(mode) % 2 ** 32)
    proven = total = 0
    for trial in range(4000):
        n = int(rng.integers(1, 120))
        blocks = int(rng.integers(1, 9))
        kp = int(rng.integers(1, 9))
        k = int(rng.integers(1, kp + 1))
        eps = float(rng.choice([0.0, 1e-3, 0.05, 0.3, 1.0]))
        spread = float(rng.choice([0.2, 1.0, 5.0]))
        total += 1
        proven += run(rng, n, blocks, kp, k, eps, mode, spread)
    assert proven > total * 0.05, (mode, proven, total)      # the proof is not vacuous on these instances


def test_unbounded_eps_proves_nothing_and_empty_corpus_is_fine():
    lists = tp.block_lists([0.5, 0.4], [0.5, 0.4], 2, 4, [0, 1])
    assert tp.finish(lists, 4, 1, math.inf) is None
    assert tp.finish([[], []], 4, 3, 0.1) == []
    assert tp.finish(lists, 4, 3, 0.0) == [(0.5, 0), (0.4, 1)]       # fewer than k rows: all of them, nothing dropped
