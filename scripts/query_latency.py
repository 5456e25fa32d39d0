"""latency of the interactive query path -- encode one short query
(css_encoder_encode, host ids in, host embedding out) + exact top-10 over a 1M x 768 corpus
(css_index_search, host in/out)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from claude_semantic_search_b200 import _native  # noqa: E402
from claude_semantic_search_b200.encoder import MPNetEncoder, random_state_dict  # noqa: E402


def pct(a, p):
    return float(np.percentile(np.asarray(a) * 1e3, p))


def main():
    import torch
    enc = MPNetEncoder(random_state_dict(0), device=0, max_tokens=4096)
    rng = np.random.default_rng(3)
    out = {}
    for L in (8, 16, 32, 48, 64, 128, 384):
        qs = [[0] + rng.integers(4, 30000, size=L - 2).tolist() + [2] for _ in range(200)]
        for q in qs[:20]:
            enc.encode_ids([q])
        lat = []
        for q in qs:
            t0 = time.perf_counter()
            enc.encode_ids([q])
            lat.append(time.perf_counter() - t0)
        out[f"encode_L{L}"] = {"p50_ms": pct(lat, 50), "p99_ms": pct(lat, 99)}
    if "--encode-only" in sys.argv:
        print(json.dumps(out))
        return
    dev = torch.device("cuda", 0)
    idx = _native.Index(768, device=0)
    idx.reserve(1_000_000)
    g = torch.Generator(device=dev).manual_seed(42)
    for _ in range(4):
        blk = torch.randn((250_000, 768), generator=g, device=dev)
        idx.add_device(blk.data_ptr(), 250_000, normalize=True, stream=torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.synchronize()
    qs = [[0] + rng.integers(4, 30000, size=14).tolist() + [2] for _ in range(300)]
    lat = []
    for i, q in enumerate(qs):
        t0 = time.perf_counter()
        e = enc.encode_ids([q])
        idx.search(e, 10)
        if i >= 50:
            lat.append(time.perf_counter() - t0)
    out["encode_L16_plus_top10_over_1M"] = {"p50_ms": pct(lat, 50), "p99_ms": pct(lat, 99)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
