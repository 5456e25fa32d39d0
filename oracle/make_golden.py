"""oracle/make_golden.py -- TEST INFRASTRUCTURE ONLY.

Generates the small committed fixtures under tests/golden/ from the oracle
(the reference's own arithmetic libraries are not installable here, so the
fixtures pin the oracle against itself across refactors and give the GPU tests
a fixed known-answer set that travels to the GPU box).

    python -m oracle.make_golden
"""
from pathlib import Path

import numpy as np

from oracle import search_oracle as so

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def build_inputs(seed: int, n: int, d: int, nq: int):
    """Deterministic inputs (numpy PCG64 stream) shared by the fixture writer and the tests."""
    rng = np.random.default_rng(seed)
    x = so.normalize_rows(rng.standard_normal((n, d), dtype=np.float32))
    q = so.normalize_rows(rng.standard_normal((nq, d), dtype=np.float32))
    # plant exact duplicates so tie-breaking by id is exercised
    x[100] = x[7]
    x[n - 148] = x[7]
    q[0] = x[7]
    mask = rng.random(n) < 0.05
    return x, q, mask


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    k = 10
    # (a) 768-d case: inputs regenerated from the seed, answers stored
    seed, n, d, nq = 20261018, 2048, 768, 8
    x, q, mask = build_inputs(seed, n, d, nq)
    D, I = so.flat_search(x, q, k)
    Dm, Im = so.flat_search(x, q, k, mask=mask)
    D100, I100 = so.flat_search(x, q, 100)
    np.savez_compressed(OUT / "search_768.npz", seed=seed, n=n, d=d, nq=nq, k=k, D=D, I=I,
                        D_masked=Dm, I_masked=Im, D100=D100, I100=I100,
                        x_checksum=np.float64(x.astype(np.float64).sum()))
    # (b) small fully stored case (inputs + answers), d = 64
    x2, q2, mask2 = build_inputs(7, 512, 64, 4)
    D2, I2 = so.flat_search(x2, q2, k)
    D2m, I2m = so.flat_search(x2, q2, k, mask=mask2)
    D2l, I2l = so.flat_search(x2, q2, k, metric=so.METRIC_L2)
    np.savez_compressed(OUT / "search_small.npz", x=x2, q=q2, k=k, D=D2, I=I2, mask=mask2,
                        D_masked=D2m, I_masked=I2m, D_l2=D2l, I_l2=I2l)
    print("wrote", sorted(p.name for p in OUT.iterdir()))


if __name__ == "__main__":
    main()
