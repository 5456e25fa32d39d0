"""Kernel list of the interactive query path (one 16-token query through css_encoder_encode), for an
ncu launch-list pass:  CSS_QUERY_GRAPH=0 ncu --metrics gpu__time_duration.sum ... python scripts/query_profile.py"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from claude_semantic_search_b200.encoder import MPNetEncoder, random_state_dict  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
enc = MPNetEncoder(random_state_dict(0), device=0, max_tokens=4096)
rng = np.random.default_rng(3)
for _ in range(4):
    enc.encode_ids([[0] + rng.integers(4, 30000, size=L - 2).tolist() + [2]])
