"""Development aid for the tcgen05 attention kernel: correctness against an fp64 softmax(QK^T/8 + bias)V on a few
ragged shapes (including scores with a wide dynamic range), then the CUDA-event time of the kernel alone at the
benchmarked shape (296 sequences x 384 tokens), printed by css_debug_attention (CSS_ATTN_TIME)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from claude_semantic_search_b200 import _native  # noqa: E402

lib = _native.load()
half = 511


def bf16(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(torch.bfloat16).to(torch.float64)


def run(lens, scale, seed, check=True):
    rng = np.random.default_rng(seed)
    T = sum(lens)
    cu = np.zeros(len(lens) + 1, np.int32)
    cu[1:] = np.cumsum(lens)
    qkv = (rng.standard_normal((T, 2304)) * scale).astype(np.float32)
    rel = rng.standard_normal((12, 2 * half + 1)).astype(np.float32)
    ctx = np.empty((T, 768), np.float32)
    _native.check(lib.css_debug_attention(qkv.ctypes.data, cu.ctypes.data, len(lens), rel.ctypes.data, half, 0,
                                          ctx.ctypes.data))
    if not check:
        return 0.0, 0.0
    qb = bf16(qkv)
    worst, mean = 0.0, 0.0
    for s, L in enumerate(lens):
        blk = qb[cu[s]:cu[s + 1]]
        q = blk[:, :768].view(L, 12, 64).transpose(0, 1)
        k = blk[:, 768:1536].view(L, 12, 64).transpose(0, 1)
        v = blk[:, 1536:].view(L, 12, 64).transpose(0, 1)
        i = torch.arange(L)
        bias = torch.from_numpy(rel).double()[:, (i[None, :] - i[:, None]) + half]
        p = torch.softmax(q @ k.transpose(1, 2) / 8.0 + bias, dim=-1)
        ref = (p @ v).transpose(0, 1).reshape(L, 768).numpy()
        err = np.abs(ctx[cu[s]:cu[s + 1]] - ref) / scale
        worst = max(worst, float(err.max()))
        mean = max(mean, float(err.mean()))
    return worst, mean


ok = True
cases = [([1], 1.5), ([2, 3], 1.5), ([65, 63], 1.5), ([128, 5, 200], 1.5), ([384], 1.5), ([129, 300, 384], 1.5),
         ([384, 383, 321, 320, 257, 193, 192, 129, 100, 7] * 4, 1.5),
         ([384, 200, 77], 0.2), ([384, 200, 77], 4.0), ([384, 257, 64], 8.0)]
for lens, scale in cases:
    w, m = run(lens, scale, seed=sum(lens))
    good = w < 4e-2 / 1.5 and m < 3e-3 / 1.5 and np.isfinite(w)
    ok &= good
    print(f"lens {str(lens)[:40]:40s} scale {scale:4.1f}  max err/scale {w:.3e}  mean {m:.3e}  {'ok' if good else 'FAIL'}",
          flush=True)
os.environ["CSS_ATTN_TIME"] = os.environ.get("CSS_ATTN_TIME", "200")
os.environ.setdefault("CSS_ATTN_TRACE", "gpurun_out/attn_trace_bench.bin")
import subprocess
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader",
                        "-lms", "10"], stdout=subprocess.PIPE, text=True)
run([384] * 296, 1.5, seed=0, check=False)
smi.terminate()
rows = [r.split(",") for r in smi.stdout.read().strip().splitlines()]
clk = sorted(int(r[0].split()[0]) for r in rows if len(r) >= 3)
if clk:
    print(f"SM clock during the run: median {clk[len(clk) // 2]} MHz, max {clk[-1]} MHz ({len(clk)} samples), last reasons {rows[-1][2].strip()}")
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
