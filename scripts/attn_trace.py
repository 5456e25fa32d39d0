"""Profiling aid: run the tcgen05 attention kernel on 296 x 384-token sequences with CSS_ATTN_TRACE
and print CTA 0's per-item timeline (clock cycles relative to the item's first stamp)."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
out = os.environ.setdefault("CSS_ATTN_TRACE", "gpurun_out/attn_trace.bin")
from claude_semantic_search_b200 import _native  # noqa: E402

n_seq, L = 296, 384
rng = np.random.default_rng(0)
qkv = rng.standard_normal((n_seq * L, 2304)).astype(np.float32)
cu = (np.arange(n_seq + 1) * L).astype(np.int32)
half = 511
rel = rng.standard_normal((12, 2 * half + 1)).astype(np.float32)
ctx = np.empty((n_seq * L, 768), np.float32)
_native.check(_native.load().css_debug_attention(qkv.ctypes.data, cu.ctypes.data, n_seq, rel.ctypes.data, half, 0,
                                                 ctx.ctypes.data))
t = np.fromfile(out, dtype=np.int64).reshape(5, 128, 8)
names = ["issuer0", "issuer1", "group0", "group1", "combine"]
base = t[2, 0, 0]
for item in range(20, 30):
    print(f"--- item {item}")
    for r in range(5):
        row = t[r, item]
        print(f"{names[r]:8s}", " ".join(f"{(v - base) if v else -1:8d}" for v in row))
d = np.diff(t[2, 10:70, 0])
print("cycles per item (group0 start to start): mean", d.mean(), "min", d.min(), "max", d.max())
