"""Timeline of one batch-1 scan (css_debug_scan_trace): when the blocks finish their sweep, merge, re-score and draw
their ticket, and what the finishing block spends on each step.  python scripts/scan_trace.py [rows]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import D, build_shard  # noqa: E402
from claude_semantic_search_b200 import _native as native  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
torch.cuda.set_stream(torch.cuda.Stream(dev))
idx = build_shard(torch, native, dev, rows, seed=42)
g = torch.Generator(device=dev).manual_seed(43)
q = torch.randn((64, D), generator=g, device=dev)
q = (q / q.norm(dim=1, keepdim=True)).contiguous()
sp = torch.cuda.current_stream(dev).cuda_stream
info = native.device_info(0)
nb = 148
for rep in range(6):
    t = idx.debug_scan_trace(q[rep].data_ptr(), 10, sp, blocks=nb)
    blk, fin = t[:nb], t[nb]
    t0 = blk[:, 0].min()
    us = lambda a: (a - t0) / 1e3
    if rep < 2:
        continue
    print(f"--- query {rep}: all times in us after the first block's start")
    for name, col in (("start", 0), ("sweep done", 1), ("list merged", 2), ("re-scored", 3), ("ticket", 4)):
        v = us(blk[:, col])
        print(f"  blocks {name:12s} min {v.min():8.2f}  median {np.median(v):8.2f}  max {v.max():8.2f}")
    names = {0: "lists requested", 5: "t0 from heads", 1: "threshold", 2: "candidates", 6: "rank-sorted", 3: "ordered", 4: "emitted"}
    for i in (0, 5, 1, 2, 6, 3, 4):
        if fin[i]:
            print(f"  last block {names[i]:16s} {us(fin[i]):8.2f}")
idx.close()
