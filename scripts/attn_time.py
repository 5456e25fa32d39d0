"""Profiling aid: CUDA-event time of the attention kernel alone (296 x 384 tokens, CSS_ATTN_TIME launches) of the
library in use (CSS_B200_LIB selects another build, e.g. one of an earlier commit, for A/B runs on the same box)."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("CSS_ATTN_TIME", "50")
from claude_semantic_search_b200 import _native  # noqa: E402

n_seq, L, half = 296, 384, 511
rng = np.random.default_rng(0)
qkv = rng.standard_normal((n_seq * L, 2304)).astype(np.float32)
cu = (np.arange(n_seq + 1) * L).astype(np.int32)
rel = rng.standard_normal((12, 2 * half + 1)).astype(np.float32)
ctx = np.empty((n_seq * L, 768), np.float32)
_native.check(_native.load().css_debug_attention(qkv.ctypes.data, cu.ctypes.data, n_seq, rel.ctypes.data, half, 0,
                                                 ctx.ctypes.data))
