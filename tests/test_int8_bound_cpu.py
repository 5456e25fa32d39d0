"""The error bound of the int8 tier of the two-phase scan, checked in float64 on the CPU (oracle/int8_bound.py restates
the kernels' quantisers and eps).  |x.q - x^.q^| <= eps must hold for EVERY row -- it is what makes the tier exact."""
import numpy as np
import pytest

from oracle import int8_bound as ib


def _check(x, q):
    x = np.ascontiguousarray(x, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    c, scale, err = ib.quantize_rows(x)
    norms = np.sqrt((x.astype(np.float64) ** 2).sum(axis=1))
    max_norm = np.float32(norms.max() * 1.0001)
    max_err8 = np.float32(err.max())
    q1, q2, a1, a2, qn, dq = ib.quantize_query(q)
    eps = ib.eps_bound(qn, dq, max_err8, max_norm)
    exact = x.astype(np.float64) @ q.astype(np.float64)
    approx = ib.approx_scores(c, scale, q1, q2, a1, a2)
    worst = float(np.abs(exact - approx).max())
    assert worst <= eps, (worst, eps)
    return worst, eps


@pytest.mark.parametrize("seed", range(6))
def test_bound_on_random_and_heavy_tailed_rows(seed):
    rng = np.random.default_rng(seed)
    d = 768
    x = rng.standard_normal((4000, d)).astype(np.float32)
    if seed % 3 == 1:
        x = (rng.standard_t(2.0, size=(4000, d)) * 0.1).astype(np.float32)          # outlier dimensions
    if seed % 3 == 2:
        x *= rng.uniform(1e-3, 30.0, size=(4000, 1)).astype(np.float32)              # un-normalised rows
    else:
        x /= np.linalg.norm(x, axis=1, keepdims=True) + 1e-8
    q = rng.standard_normal(d).astype(np.float32) * np.float32(10.0 ** rng.integers(-3, 3))
    worst, eps = _check(x, q)
    assert eps < 0.2 * np.linalg.norm(q) * np.linalg.norm(x, axis=1).max()          # a useful bound, not a vacuous one


def test_bound_is_tight_for_aligned_rounding_errors():
    """Rows whose every component sits 0.49 steps off the grid, queried along the all-ones direction: equality in
    Cauchy-Schwarz up to the slack factors (the case tests/test_search_int8_gpu.py plants on the device)."""
    d = 768
    s = np.float32(2.0 ** -9)
    x = np.full((8, d), np.float32(20.49) * s, np.float32)
    x[:, 0] = 127 * s
    x[4:, 1:] = np.float32(20.51) * s
    q = np.ones(d, np.float32) / np.float32(np.sqrt(d))
    worst, eps = _check(x, q)
    assert worst > 0.9 * eps / 1.01, (worst, eps)


def test_bound_degenerate_inputs():
    d = 768
    rng = np.random.default_rng(3)
    x = rng.standard_normal((64, d)).astype(np.float32)
    x[0] = 0                                     # zero row: scale 0, codes 0
    x[1, :] = 0
    x[1, 5] = 3.0                                # one-hot
    x[2] = 1e-30                                 # tiny
    x[3] = np.float32(1e18)                      # huge
    for q in (np.zeros(d, np.float32), np.eye(d, dtype=np.float32)[7] * 5, rng.standard_normal(d).astype(np.float32) * 1e-20,
              np.full(d, 1e15, np.float32)):
        _check(x, q)


def test_non_finite_rows_make_the_bound_infinite():
    rng = np.random.default_rng(4)
    x = rng.standard_normal((8, 768)).astype(np.float32)
    x[3, 9] = np.inf
    x[5, 2] = np.nan
    c, scale, err = ib.quantize_rows(x)
    assert np.isinf(err[3]) and np.isinf(err[5]) and np.isfinite(np.delete(err, [3, 5])).all()
    assert (c[3] == 0).all() and (c[5] == 0).all() and scale[3] == 0 and scale[5] == 0
