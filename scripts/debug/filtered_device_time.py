"""Debug aid: device time of the masked batch-1 search at 5 % selectivity (2M rows) through css_index_search_device
with an external mask pointer vs css_index_search with the same filter; prints the two-phase counters."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from claude_semantic_search_b200 import _native as native  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_stream(torch.cuda.Stream(dev))
rows, D, K = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000, 768, 10
idx = native.Index(D)
g = torch.Generator(device=dev).manual_seed(1)
for r0 in range(0, rows, 500_000):
    x = torch.randn((500_000, D), device=dev, generator=g)
    x /= x.norm(dim=1, keepdim=True)
    idx.add_device(x.data_ptr(), 500_000, False)
    torch.cuda.synchronize()
mask = (np.random.default_rng(0).random(rows) < 0.05)
words = np.packbits(np.pad(mask, (0, (-rows) % 32)).reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1)
flt = native.Filter().set_row_mask(words)
if len(sys.argv) > 2:   # a clause filter like bench.py's configs[4] leg instead of a row mask
    rng = np.random.default_rng(99)
    ts = rng.integers(0, 731, size=rows).astype(np.int32)
    hc = (rng.random(rows) < 0.4).astype(np.int32)
    idx.set_column(4, ts)
    idx.set_column(5, hc)
    flt = native.Filter().add_range(4, 100, 100 + 91 - 1).add_range(5, 1, 1)   # 12.4 % x 40 % = 5 %
q = np.random.default_rng(43).standard_normal((64, D)).astype(np.float32)
q /= np.linalg.norm(q, axis=1, keepdims=True)
for i in range(8):
    idx.search(q[i:i + 1], K, flt)
print("after host searches:", idx.scan_stats())
t0 = time.perf_counter()
for i in range(100):
    idx.search(q[i % 64:i % 64 + 1], K, flt)
print("host API, same filter: %.3f ms per query" % ((time.perf_counter() - t0) * 10))
sp = torch.cuda.current_stream(dev).cuda_stream
mptr, _ = idx.filter_mask_device(flt, sp)
Dd = torch.empty((1, K), device=dev)
Id = torch.empty((1, K), device=dev, dtype=torch.int64)
qd = torch.from_numpy(q).to(dev)
for name in ("search_device + mask pointer",):
    for i in range(5):
        idx.search_device(qd[i].data_ptr(), 1, K, Dd.data_ptr(), Id.data_ptr(), mptr, 0, sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100):
        idx.search_device(qd[i % 64].data_ptr(), 1, K, Dd.data_ptr(), Id.data_ptr(), mptr, 0, sp)
    e1.record()
    torch.cuda.synchronize()
    print(name, "%.3f ms per query (device)" % (e0.elapsed_time(e1) / 100), idx.scan_stats())
