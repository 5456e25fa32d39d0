"""oracle/int8_bound.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

numpy restatement of the arithmetic and of the error bound of the int8 tier of the two-phase exact scan
(claude_semantic_search_b200/csrc/index_kernels.cuh: append_rows_kernel -- row codes, scale, max_err8;
quantize_query_i8 -- the query's two code vectors; dot_i8 -- the approximate score; scan_one_query -- eps).
The reference has no counterpart (faiss IndexFlatIP scores in fp32, src/storage.py:436): the tier is an
internal accelerator whose only contract is that, with eps as computed here,

    | x . q  -  x^ . q^ |  <=  eps        for every stored row x,

so that the candidate set {rows with approximate score >= t - 2 eps} provably contains the exact top-k.
tests/test_int8_bound_cpu.py checks that inequality in float64 on random, heavy-tailed, adversarial and
degenerate inputs; the GPU tests (tests/test_search_int8_gpu.py) check the kernels' end result.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def quantize_rows(x: np.ndarray):
    """append_rows_kernel: codes rint(v / scale) with scale = max|v| / 127 per row, and the quantisation-error norm
    the kernel folds into max_err8 (x 1.001, computed from the stored codes).  x: float32 [n, d]."""
    x = np.ascontiguousarray(x, F)
    amax = np.abs(x).max(axis=1)
    finite = np.isfinite((x.astype(np.float64) ** 2).sum(axis=1))
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = np.where(finite, amax / F(127), F(0)).astype(F)
        inv = np.where(finite & (amax > 0), F(127) / amax, F(0)).astype(F)
        c = np.clip(np.rint(x * inv[:, None]), -127, 127)
    c = np.where(finite[:, None], np.nan_to_num(c, nan=-127.0), 0.0).astype(np.int32)
    e = x.astype(np.float64) - scale.astype(np.float64)[:, None] * c
    err = np.sqrt((e * e).sum(axis=1))
    err = np.where(finite, err * 1.001, np.inf)
    return c, scale, err


def quantize_query(q: np.ndarray):
    """quantize_query_i8: q ~= a1 q1 + a2 q2, a2 = a1 / 254; returns (q1, q2, a1, a2, qn, dq) with qn >= ||q|| and
    dq >= ||q - (a1 q1 + a2 q2)|| as the kernel computes them."""
    q = np.ascontiguousarray(q, F).reshape(-1)
    amax = F(np.abs(q).max())
    a1 = F(amax / F(127))
    a2 = F(a1 * F(1.0 / 254.0))
    inv1 = F(127) / amax if amax > 0 else F(0)
    inv2 = F(1) / a2 if a2 > 0 else F(0)
    with np.errstate(over="ignore", invalid="ignore"):
        c1 = np.clip(np.rint(q * inv1), -127, 127).astype(F)
        r = (q.astype(np.float64) - np.float64(a1) * c1).astype(F)            # fmaf(-a1, c1, v): one rounding
        c2 = np.clip(np.nan_to_num(np.rint(r * inv2), nan=-127.0), -127, 127).astype(F)
        r2 = (r.astype(np.float64) - np.float64(a2) * c2).astype(F)           # fmaf(-a2, c2, r)
    qn = F(np.sqrt(np.float64((q.astype(np.float64) ** 2).sum()))) * F(1.00001)
    dq = F(np.sqrt(np.float64((r2.astype(np.float64) ** 2).sum()))) * F(1.001)
    return c1.astype(np.int32), c2.astype(np.int32), a1, a2, qn, dq


def approx_scores(c: np.ndarray, scale: np.ndarray, q1, q2, a1, a2) -> np.ndarray:
    """dot_i8, in exact arithmetic: x^ . q^ = scale * (a1 I1 + a2 I2) with integer dot products."""
    i1 = c.astype(np.int64) @ q1.astype(np.int64)
    i2 = c.astype(np.int64) @ q2.astype(np.int64)
    return scale.astype(np.float64) * (np.float64(a1) * i1 + np.float64(a2) * i2)


def eps_bound(qn, dq, max_err8, max_norm) -> float:
    """scan_one_query (int8 tier): eps = 1.001 (qn e8 + dq (mn + e8)) + 4e-6 qn mn, in float32 like the kernel."""
    qn, dq, e8, mn = F(qn), F(dq), F(max_err8), F(max_norm)
    return float(F(1.001) * (qn * e8 + dq * (mn + e8)) + F(4e-6) * qn * mn)
