"""SURVEY 8(f) row 1 / VERDICT r1 item 8: how fast does css_index_load bring an index file back into HBM, next to the
rate at which the same file can be read at all (a plain sequential read into a pinned buffer), and css_index_save next
to a plain write.  Usage: python scripts/load_bandwidth.py [rows] [dir]"""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from claude_semantic_search_b200 import _native as native  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
where = Path(sys.argv[2] if len(sys.argv) > 2 else "/tmp")
path = where / "css_load_bandwidth.faiss"
dev = torch.device("cuda", 0)
idx = native.Index(768)
g = torch.Generator(device=dev).manual_seed(1)
for r0 in range(0, rows, 500_000):
    n = min(500_000, rows - r0)
    x = torch.randn((n, 768), device=dev, generator=g)
    x /= x.norm(dim=1, keepdim=True)
    torch.cuda.synchronize()
    idx.add_device(x.data_ptr(), n, False)
    idx.search(np.zeros((1, 768), np.float32), 1)   # drains the handle stream before x is overwritten
q = np.random.default_rng(0).standard_normal((4, 768)).astype(np.float32)
D0, I0 = idx.search(q, 10)
t0 = time.perf_counter()
idx.save(str(path))
t_save = time.perf_counter() - t0
size = path.stat().st_size
idx.close()
os.sync()
# plain sequential read of the same file (page cache dropped when permitted, else reported as cached)
dropped = False
try:
    with open("/proc/sys/vm/drop_caches", "w") as fh:
        fh.write("3\n")
    dropped = True
except OSError:
    pass
buf = torch.empty(64 << 20, dtype=torch.uint8).pin_memory().numpy()
t0 = time.perf_counter()
with open(path, "rb", buffering=0) as fh:
    while fh.readinto(buf):
        pass
t_read = time.perf_counter() - t0
if dropped:
    with open("/proc/sys/vm/drop_caches", "w") as fh:
        fh.write("3\n")
idx2 = native.Index(768)
t0 = time.perf_counter()
idx2.load(str(path))
t_load = time.perf_counter() - t0
D1, I1 = idx2.search(q, 10)
assert idx2.ntotal == rows and np.array_equal(I0, I1) and np.array_equal(D0, D1)
idx2.close()
path.unlink()
gb = size / 1e9
print(f"{rows} rows, {gb:.2f} GB file in {where} (page cache {'dropped' if dropped else 'NOT dropped: reads may be cached'}): "
      f"plain read {gb / t_read:.2f} GB/s, css_index_load {gb / t_load:.2f} GB/s = {t_read / t_load * 100:.0f} % of it "
      f"(rows, bf16 shadow and maxima rebuilt on the device), css_index_save {gb / t_save:.2f} GB/s; results identical after reload")
