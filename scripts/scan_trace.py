"""Timeline of batch-1 scans (css_debug_scan_trace): when the blocks finish their sweep, merge, re-score and draw
their ticket, and what the finishing block adds.  python scripts/scan_trace.py [rows] [queries]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import D, build_shard  # noqa: E402
from claude_semantic_search_b200 import _native as native  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
torch.cuda.set_stream(torch.cuda.Stream(dev))
idx = build_shard(torch, native, dev, rows, seed=42)
g = torch.Generator(device=dev).manual_seed(43)
q = torch.randn((nq + 4, D), generator=g, device=dev)
q = (q / q.norm(dim=1, keepdim=True)).contiguous()
sp = torch.cuda.current_stream(dev).cuda_stream
nb = 148
acc = {k: [] for k in ("sweep first", "sweep median", "sweep last", "merged last", "re-scored last", "ticket last",
                       "finish starts", "emitted", "finish")}
for rep in range(nq + 4):
    t = idx.debug_scan_trace(q[rep].data_ptr(), 10, sp, blocks=nb)
    if rep < 4:
        continue
    blk, fin = t[:nb], t[nb]
    t0 = blk[:, 0].min()
    us = lambda a: (a - t0) / 1e3
    acc["sweep first"].append(us(blk[:, 1]).min())
    acc["sweep median"].append(np.median(us(blk[:, 1])))
    acc["sweep last"].append(us(blk[:, 1]).max())
    acc["merged last"].append(us(blk[:, 2]).max())
    acc["re-scored last"].append(us(blk[:, 3]).max())
    acc["ticket last"].append(us(blk[:, 4]).max())
    acc["finish starts"].append(us(fin[0]))
    acc["emitted"].append(us(fin[4]))
    acc["finish"].append((fin[4] - fin[0]) / 1e3)
print(f"{rows} rows, {nq} queries; us after the first block's start: mean / median / min / max")
for k, v in acc.items():
    v = np.asarray(v)
    print(f"  {k:16s} {v.mean():8.2f} {np.median(v):8.2f} {v.min():8.2f} {v.max():8.2f}")
idx.close()
