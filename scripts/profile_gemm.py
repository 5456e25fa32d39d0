"""Profiling aid: run the encoder's QKV-shaped GEMM (98304 x 2304 x 768) through css_debug_gemm
with the single-CTA (mode 2) and 2-CTA (mode 4) kernels.  Usage: python scripts/profile_gemm.py [M N K gelu]"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from claude_semantic_search_b200 import _native  # noqa: E402

M, N, K, gelu = (int(a) for a in (sys.argv[1:5] + ["98304", "2304", "768", "0"][len(sys.argv) - 1:]))
rng = np.random.default_rng(0)
A = rng.standard_normal((M, K), dtype=np.float32)
B = (rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K)).astype(np.float32)
bias = rng.standard_normal(N).astype(np.float32)
out = np.empty((M, N), np.float32)
lib = _native.load()
for mode in (2, 4, 2, 4):
    t0 = time.perf_counter()
    _native.check(lib.css_debug_gemm(A.ctypes.data, B.ctypes.data, bias.ctypes.data, M, N, K, gelu | mode, 0,
                                     out.ctypes.data))
    print("mode", mode, "wall", time.perf_counter() - t0, "checksum", float(out[::997].sum()))
