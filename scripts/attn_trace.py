"""Profiling aid: run the tcgen05 attention kernel on 296 x 384-token sequences with CSS_ATTN_TRACE and print the
per-item phase breakdown (clock cycles, mean over items 10..70 of the last of 30 launches) of the CTAs given on the
command line (default 0), plus the raw timeline of a few items of the first one."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
out = os.environ.setdefault("CSS_ATTN_TRACE", "gpurun_out/attn_trace.bin")
os.environ.setdefault("CSS_ATTN_TIME", "30")
from claude_semantic_search_b200 import _native  # noqa: E402

ctas = [int(a) for a in sys.argv[1:]] or [0]
n_seq, L = 296, 384
rng = np.random.default_rng(0)
qkv = rng.standard_normal((n_seq * L, 2304)).astype(np.float32)
cu = (np.arange(n_seq + 1) * L).astype(np.int32)
half = 511
rel = rng.standard_normal((12, 2 * half + 1)).astype(np.float32)
ctx = np.empty((n_seq * L, 768), np.float32)
names = ["qk", "pv_go", "group0", "pv_done", "combine"]
for k, cta in enumerate(ctas):
    os.environ["CSS_ATTN_TRACE_CTA"] = str(cta)
    _native.check(_native.load().css_debug_attention(qkv.ctypes.data, cu.ctypes.data, n_seq, rel.ctypes.data, half, 0,
                                                     ctx.ctypes.data))
    t = np.fromfile(out, dtype=np.int64)[:6 * 128 * 8].reshape(6, 128, 8)
    g = t[2, 10:71].astype(np.float64)          # group 0: start, S ready, pass 1 done, barrier issued, block 0 / 1 / 2 done
    per = np.diff(g[:, 0])
    x = g[:-1]
    print(f"CTA {cta}: {per.mean():.0f} cycles per item (min {per.min():.0f}, max {per.max():.0f}): wait S {np.mean(x[:, 1] - x[:, 0]):.0f}"
          f" | pass 1 {np.mean(x[:, 2] - x[:, 1]):.0f} | barrier + block 0 {np.mean(x[:, 4] - x[:, 2]):.0f}"
          f" | block 1 {np.mean(x[:, 5] - x[:, 4]):.0f} | block 2 {np.mean(x[:, 6] - x[:, 5]):.0f} | tail {np.mean(g[1:, 0] - x[:, 6]):.0f}",
          flush=True)
    if k == 0:
        base = t[2, 0, 0]
        for item in range(21, 25):
            print(f"--- item {item}")
            for r in range(5):
                print(f"{names[r]:8s}", " ".join(f"{(v - base) if v else -1:8d}" for v in t[r, item]))
