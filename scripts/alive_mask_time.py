"""Batch-1 search time of a 1 M-row index before and after a few rows were deleted (after a deletion every search
is a masked scan: the alive bits).  python scripts/alive_mask_time.py"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import D, K, build_shard, time_region  # noqa: E402
from claude_semantic_search_b200 import _native as native  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
torch.cuda.set_stream(torch.cuda.Stream(dev))
rows = 1_000_000
idx = build_shard(torch, native, dev, rows, seed=42)
g = torch.Generator(device=dev).manual_seed(43)
q = torch.randn((256, D), generator=g, device=dev)
q = (q / q.norm(dim=1, keepdim=True)).contiguous()
qh = q.cpu().numpy()
sp = torch.cuda.current_stream(dev).cuda_stream
Dd = torch.empty((1, K), device=dev, dtype=torch.float32)
Id = torch.empty((1, K), device=dev, dtype=torch.int64)


def measure(tag):
    f = lambda i: idx.search_device(q[i % 256].data_ptr(), 1, K, Dd.data_ptr(), Id.data_ptr(), 0, 0, sp)
    for i in range(10):
        f(i)
    ms, _ = time_region(torch, dev, f, 300)
    for i in range(10):
        idx.search(qh[i:i + 1], K)
    t0 = time.perf_counter()
    for i in range(300):
        idx.search(qh[i % 256:i % 256 + 1], K)
    e2e = (time.perf_counter() - t0) / 300 * 1e3
    print(f"{tag:28s} device {ms / 300:.4f} ms per query, host API {e2e:.4f} ms per query")


measure("all rows alive (dense sweep)")
ref = [idx.search(qh[i:i + 1], K) for i in range(8)]
dead = np.random.default_rng(1).choice(rows, size=100, replace=False).astype(np.int64)
idx.set_alive_ids(dead, False)
measure("100 rows deleted (masked)")
for i in range(8):
    Dn, In = idx.search(qh[i:i + 1], K)
    keep = ~np.isin(ref[i][1][0], dead)
    assert (In[0][:keep.sum()] == ref[i][1][0][keep]).all(), "masked result differs from the dense result minus the deleted rows"
print("masked results == dense results minus the deleted rows")
idx.close()
