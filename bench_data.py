"""Synthetic corpora for bench.py and the GPU tests (data generation only -- no search logic here).

`clustered_*`: the i.i.d. random unit vectors of BASELINE configs[1] have no near neighbours (score sigma 0.036,
top-10 at ~0.18), which flatters any scan that prunes by score.  Real chunk embeddings are clustered: the chunks of
one session sit next to each other in the index AND in embedding space.  This generator draws C cluster centres on
the sphere and rows as normalise(centre + sigma_c * g / sqrt(d)), cos(row, centre) = 1 / sqrt(1 + sigma_c^2) drawn
per cluster from U[cos_lo, cos_hi] (von-Mises-Fisher-like caps); order="session" keeps the members of a cluster
in adjacent rows (the realistic layout, and the adversarial one for per-block candidate lists), "shuffled"
deals them at random.  Queries are perturbed corpus rows, so every query has a dense neighbourhood.
"""
from __future__ import annotations

import numpy as np


def _cluster_sigma(rng, n_clusters, cos_lo, cos_hi):
    cos = rng.uniform(cos_lo, cos_hi, size=n_clusters)
    return np.sqrt(1.0 / (cos * cos) - 1.0).astype(np.float32)


def clustered_numpy(n: int, d: int = 768, n_clusters: int = 2000, cos_lo: float = 0.6, cos_hi: float = 0.95,
                    order: str = "session", seed: int = 7, n_queries: int = 64):
    """(x [n,d] float32 unit rows, q [n_queries,d] float32 unit rows, cluster id per row)."""
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((n_clusters, d)).astype(np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    sigma = _cluster_sigma(rng, n_clusters, cos_lo, cos_hi)
    if order == "session":
        cid = (np.arange(n, dtype=np.int64) * n_clusters // max(n, 1)).astype(np.int64)
    else:
        cid = rng.integers(0, n_clusters, size=n)
    x = np.empty((n, d), np.float32)
    for r0 in range(0, n, 65536):
        c = cid[r0:r0 + 65536]
        g = rng.standard_normal((c.shape[0], d)).astype(np.float32)
        v = centres[c] + (sigma[c] / np.sqrt(d))[:, None] * g
        x[r0:r0 + 65536] = v / np.linalg.norm(v, axis=1, keepdims=True)
    pick = rng.choice(n, size=n_queries, replace=n_queries > n)
    q = x[pick] + (0.3 / np.sqrt(d)) * rng.standard_normal((n_queries, d)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return x, q.astype(np.float32), cid


def clustered_torch_tiles(torch, dev, n: int, d: int = 768, n_clusters: int = 2000, cos_lo: float = 0.6,
                          cos_hi: float = 0.95, order: str = "session", seed: int = 7, tile: int = 250_000):
    """Generator of device tiles [rows, d] float32 (unit rows) of the same distribution, for corpora that are
    built on the device; also yields nothing else -- queries come from `clustered_queries_torch`."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    centres = torch.randn((n_clusters, d), generator=g, device=dev, dtype=torch.float32)
    centres /= centres.norm(dim=1, keepdim=True)
    cos = torch.rand(n_clusters, generator=g, device=dev) * (cos_hi - cos_lo) + cos_lo
    sigma = torch.sqrt(1.0 / (cos * cos) - 1.0)
    for r0 in range(0, n, tile):
        nr = min(tile, n - r0)
        if order == "session":
            cid = (torch.arange(r0, r0 + nr, device=dev, dtype=torch.int64) * n_clusters) // max(n, 1)
        else:
            cid = torch.randint(0, n_clusters, (nr,), generator=g, device=dev)
        v = centres[cid] + (sigma[cid] / d ** 0.5)[:, None] * torch.randn((nr, d), generator=g, device=dev)
        yield v / v.norm(dim=1, keepdim=True)


def perturbed_queries_torch(torch, rows, n_queries: int, seed: int = 11, noise: float = 0.3):
    """Queries = perturbed rows of a device tile `rows` [m, d]."""
    g = torch.Generator(device=rows.device)
    g.manual_seed(seed)
    pick = torch.randint(0, rows.shape[0], (n_queries,), generator=g, device=rows.device)
    q = rows[pick] + (noise / rows.shape[1] ** 0.5) * torch.randn((n_queries, rows.shape[1]), generator=g, device=rows.device)
    return q / q.norm(dim=1, keepdim=True)
