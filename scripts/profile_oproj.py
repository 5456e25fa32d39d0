"""Profiling aid: the K = 768, N = 768 projection at M = 113664 with three epilogues -- bias only
(css_debug_gemm), + residual + row statistics + apply kernel (mode 1), + fused LayerNorm (mode 3)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from claude_semantic_search_b200 import _native  # noqa: E402

M, N, K = 113664, 768, 768
rng = np.random.default_rng(0)
A = rng.standard_normal((M, K), dtype=np.float32)
B = (rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K)).astype(np.float32)
bias = rng.standard_normal(N).astype(np.float32)
resid = rng.standard_normal((M, N), dtype=np.float32)
gamma = np.ones(N, np.float32)
beta = np.zeros(N, np.float32)
out = np.empty((M, N), np.float32)
lib = _native.load()
for rep in range(2):
    _native.check(lib.css_debug_gemm(A.ctypes.data, B.ctypes.data, bias.ctypes.data, M, N, K, 4, 0, out.ctypes.data))
    for mode in (1, 3):
        _native.check(lib.css_debug_gemm_resid_ln(A.ctypes.data, B.ctypes.data, bias.ctypes.data, resid.ctypes.data,
                                                  gamma.ctypes.data, beta.ctypes.data, M, K, 1e-5, mode, 0, out.ctypes.data))
print("done")
