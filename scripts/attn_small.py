"""Debug aid: one small ragged launch of the tcgen05 attention kernel (for compute-sanitizer)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from claude_semantic_search_b200 import _native  # noqa: E402

lens = [int(a) for a in sys.argv[1:]] or [128, 5, 200]
half = 511
rng = np.random.default_rng(0)
T = sum(lens)
qkv = rng.standard_normal((T, 2304)).astype(np.float32)
cu = np.zeros(len(lens) + 1, np.int32)
cu[1:] = np.cumsum(lens)
rel = rng.standard_normal((12, 2 * half + 1)).astype(np.float32)
ctx = np.empty((T, 768), np.float32)
_native.check(_native.load().css_debug_attention(qkv.ctypes.data, cu.ctypes.data, len(lens), rel.ctypes.data, half, 0,
                                                 ctx.ctypes.data))
print("ok", float(np.abs(ctx).mean()))
