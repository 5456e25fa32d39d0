#!/usr/bin/env python
"""bench.py -- hot-path benchmark of claude_semantic_search_b200 (contract in the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Headline workload (BASELINE.json configs[1]): exact top-10 over a 1M x 768 fp32 corpus per GPU, batch-1
queries.  A "step" is one query call.  With N > 1 ranks (torchrun) every rank holds its own 1M-row shard
(weak scaling: the corpus grows with N), queries are replicated, each rank scans its shard and the local
top-k lists are exchanged and merged INSIDE the scan kernel over NVLink (no NCCL call on the query path).
`value` counts 1M-row shard scans per second over all ranks (= QPS x N; at N = 1 it IS the QPS), so the
unit is the same at every N and in the reference arm.

In-run correctness (before anything is timed): the merged top-k of 8 queries -- one of them filtered -- is
compared on every rank with a brute-force recomputation (regenerated tiles, torch matmul + topk, all-gather).

Extra legs (own timers, reported under "extra"):
  batch1024        tensor-core batch path on the same corpus
  fp32_sweep       the single fp32 sweep (SURVEY 8d's 3072 B/row definition) timed beside the two-phase scan
  clustered_1M     the same search on a clustered corpus (bench_data.py), fallback rate of the proof
  filtered_10M     BASELINE configs[4] (N = 1): date range AND project AND has_code, ~5 % selectivity
  config4          BASELINE configs[3] (N = 8): 100M x 768 = 12.5M rows per GPU, planted needles asserted
  sharded_handle   the single-process multi-device index (css_index_create_sharded) behind the host API
  encode           BASELINE configs[2]: MPNet encode at seq len 384 (bench_encoder.py)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

D = 768
K = 10
ROWS = 1_000_000
BYTES_PER_ROW = D * 4  # 3072 B per corpus row per batch-1 query
WORKLOAD = "exact top-10, 1M x 768 fp32 corpus per GPU, batch-1 queries (BASELINE configs[1])"
METRIC = "exact top-10 search throughput @ 1Mx768 fp32 per GPU, batch-1 (1M-row shard scans/s = QPS x n_gpus)"
UNIT = "1M-row shard scans/s"   # rows_per_gpu-row shards; renamed by unit_for() when --rows is overridden


def unit_for(rows: int) -> str:
    return UNIT if rows == ROWS else f"{rows}-row shard scans/s"
# DRAM traffic per launch from the committed ncu --set full captures, keyed by rows per GPU
NCU_SCAN_TRAFFIC = {1_000_000: 3_072_071_000 + 4_015_872}       # fp32 sweep (profiles/r1_ncu_kernels_summary.txt)
NCU_SCAN_TRAFFIC_BF16 = {1_000_000: 1_536_171_000 + 7_699_456}  # bf16 sweep incl. its last-CTA finish (profiles/r2_scan_two_phase_ncu_summary.txt)
NCU_SCAN_TRAFFIC_INT8 = {1_000_000: 786_737_152 + 6_053_632}     # int8 sweep incl. block re-scores + last-CTA finish (profiles/r2_scan_int8_ncu_summary.txt)
INT8_ROW_BYTES = D + 4                                           # 768 codes + the row's fp32 scale


def shared_config(rows: int, world: int) -> dict:
    """The `config` object: identical in the b200 and the reference arm."""
    return {"workload": WORKLOAD, "rows_per_gpu": rows, "corpus_rows": rows * world, "dim": D, "k": K,
            "query_batch": 1, "l2": "corpus (>= 1.5 GB per sweep) >> 126 MB L2, no flush needed"}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j["bf16_tflops"],
                "bf16_tflops_sustained": j.get("bf16_tflops_sustained", j["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# --------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------
# CPU legs: the oracle's C restatement of faiss IndexFlatIP.search (oracle/flat_ip.c, OpenMP over rows)
def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def host_corpus(rows: int):
    """1M x 768 fp32 unit rows on the host (the distribution of the GPU arm's shards), generated in parallel
    slabs so that the reference arm spends its time scanning, not drawing random numbers."""
    from concurrent.futures import ThreadPoolExecutor
    x = np.empty((rows, D), np.float32)
    slab = 62_500

    def fill(i):
        r0 = i * slab
        rng = np.random.default_rng(1000 + i)
        blk = rng.standard_normal((min(slab, rows - r0), D), dtype=np.float32)
        blk /= (np.linalg.norm(blk, axis=1, keepdims=True) + 1e-8)
        x[r0:r0 + blk.shape[0]] = blk
    with ThreadPoolExecutor(max_workers=min(16, host_threads())) as ex:
        list(ex.map(fill, range((rows + slab - 1) // slab)))
    return x


def cpu_search_baseline(rows: int = ROWS, seconds: float = 12.0):
    """Bounded sample: batch-1 queries over the FULL rows x 768 corpus on all host threads for ~`seconds`."""
    from oracle import search_oracle as so
    x = host_corpus(rows)
    q = so.normalize_rows(np.random.default_rng(43).standard_normal((64, D), dtype=np.float32))
    T = host_threads()
    so.flat_search_c(x, q[:1], K, nthreads=T)  # warm-up (page faults, thread team)
    n, t0 = 0, time.perf_counter()
    while True:
        so.flat_search_c(x, q[n % 64:n % 64 + 1], K, nthreads=T)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= seconds or n >= 400:
            break
    qps = n / dt
    return {"value": qps, "unit": unit_for(rows), "cores": int(T), "kind": "port",
            "sample": f"{n} batch-1 queries over all {rows} rows x {D} fp32 on the host "
                      f"(oracle/flat_ip.c, the port of faiss IndexFlatIP.search, {T} OpenMP threads), {dt:.1f} s"}


def run_reference(args):
    """The reference's CPU path for the headline workload, same config / metric / unit as the b200 arm: every
    step is one batch-1 query over the whole corpus of that arm (N x 1M rows: the host scans the 1M-row shard
    N times -- one resident copy, same bytes streamed), all host threads.  Rank 0 alone runs under torchrun."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, args.gpus)
    steps, warm = args.steps, max(args.warmup, 1)
    from oracle import search_oracle as so
    rows = args.rows
    x = host_corpus(rows)
    q = so.normalize_rows(np.random.default_rng(43).standard_normal((64, D), dtype=np.float32))
    T = host_threads()

    def step(i):
        for _ in range(world):
            so.flat_search_c(x, q[i % 64:i % 64 + 1], K, nthreads=T)
    for i in range(min(warm, 3)):
        step(i)
    lat = []
    t0 = time.perf_counter()
    for i in range(steps):
        t1 = time.perf_counter()
        step(i)
        lat.append((time.perf_counter() - t1) * 1e3)
    dt = time.perf_counter() - t0
    ms = dt / steps * 1e3
    value = world * 1e3 / ms          # 1M-row shard scans per second
    sample = (f"each step = 1 batch-1 query over {world} x {rows} rows x {D} fp32 on the host (oracle/flat_ip.c, the port of "
              f"faiss IndexFlatIP.search; {T} OpenMP threads" + (f"; the {rows}-row shard is scanned {world} times" if world > 1 else "") + ")")
    UNIT = unit_for(rows)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": shared_config(rows, world),
            "latency_ms": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99))},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(T), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------
TILE = 250_000


def gen_tiles(torch, dev, rows: int, seed: int):
    """The corpus of one shard as device tiles (N(0,1); row-normalised by the add kernel with the reference's
    x / (||x|| + 1e-8)).  Deterministic in (seed, rows): the brute-force checker regenerates the same tiles."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    for r0 in range(0, rows, TILE):
        nr = min(TILE, rows - r0)
        yield r0, torch.randn((nr, D), generator=g, device=dev, dtype=torch.float32)


def build_shard(torch, native, dev, rows: int, seed: int, tiles=None, plant=None):
    idx = native.Index(D, native.METRIC_INNER_PRODUCT, dev.index)
    idx.reserve(rows)
    stream = torch.cuda.current_stream(dev).cuda_stream
    for r0, blk in (tiles if tiles is not None else gen_tiles(torch, dev, rows, seed)):
        if plant is not None:
            plant(r0, blk)
        idx.add_device(blk.data_ptr(), blk.shape[0], normalize=True, stream=stream)
        torch.cuda.current_stream(dev).synchronize()
        del blk
    return idx


def brute_force_topk(torch, dev, tiles, q, k, id_offset=0, mask=None, plant=None):
    """Slow, trivially correct: fp32 matmul + topk per regenerated tile, merged.  Returns (scores, ids) [nq, k]."""
    best_s = torch.full((q.shape[0], 0), 0.0, device=dev)
    best_i = torch.zeros((q.shape[0], 0), dtype=torch.int64, device=dev)
    for r0, blk in tiles:
        if plant is not None:
            plant(r0, blk)
        blk = blk / (blk.norm(dim=1, keepdim=True) + 1e-8)
        s = q @ blk.T
        if mask is not None:
            s = s.masked_fill(~mask[r0:r0 + blk.shape[0]].unsqueeze(0), float("-inf"))
        kk = min(k, s.shape[1])
        ts, ti = torch.topk(s, kk, dim=1)
        best_s = torch.cat([best_s, ts], dim=1)
        best_i = torch.cat([best_i, ti + r0 + id_offset], dim=1)
        kk = min(k, best_s.shape[1])
        ts, sel = torch.topk(best_s, kk, dim=1)
        best_s, best_i = ts, torch.gather(best_i, 1, sel)
    return best_s, best_i


def check_topk(D_got, I_got, ref_s, ref_i, k, tol=1e-4):
    """Tolerance-aware: scores within tol of the brute-force list position by position; ids identical except
    where the brute-force scores of the two candidates differ by less than tol (north_star's parity definition)."""
    D_got, I_got = np.asarray(D_got), np.asarray(I_got)
    ref_s, ref_i = np.asarray(ref_s), np.asarray(ref_i)
    for qi in range(D_got.shape[0]):
        valid = np.isfinite(ref_s[qi][:k])
        nv = int(valid.sum())
        if not np.allclose(D_got[qi][:nv], ref_s[qi][:nv], atol=tol):
            return False, f"query {qi}: scores {D_got[qi][:nv]} vs {ref_s[qi][:nv]}"
        if (I_got[qi][nv:] != -1).any():
            return False, f"query {qi}: expected {k - nv} unfilled slots"
        lookup = {int(i): float(s) for s, i in zip(ref_s[qi], ref_i[qi])}
        for j in range(nv):
            gid = int(I_got[qi][j])
            if gid == int(ref_i[qi][j]):
                continue
            if gid not in lookup or abs(lookup[gid] - float(ref_s[qi][j])) > tol:
                return False, f"query {qi} slot {j}: id {gid} vs {int(ref_i[qi][j])} (not a near-tie)"
    return True, ""


def time_region(torch, dev, fn, steps: int, dist=None):
    """Barrier + sync, K steps bracketed by CUDA events on the launching stream, sync + barrier."""
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    if dist is not None:
        dist.barrier()
    return e0.elapsed_time(e1), wall * 1e3


def max_over_ranks(torch, dev, dist, v: float) -> float:
    if dist is None:
        return v
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def per_query_latency(torch, dev, fn, n: int, dist=None):
    """p50 / p99 of per-query device times (one CUDA event pair per query) and of wall times."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    for i in range(n):
        ev[i][0].record()
        fn(i)
        ev[i][1].record()
    torch.cuda.synchronize(dev)
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    return {"p50_ms": float(np.percentile(ms, 50)), "p99_ms": float(np.percentile(ms, 99)), "mean_ms": float(ms.mean()),
            "n": n, "clock": "CUDA events around every query, launching stream"}


def wall_latency(fn, n: int):
    lat = np.empty(n)
    for i in range(n):
        t0 = time.perf_counter()
        fn(i)
        lat[i] = (time.perf_counter() - t0) * 1e3
    return {"p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)), "mean_ms": float(lat.mean()),
            "n": n, "clock": "host wall clock around every call (host buffers in, host result out)"}


# --------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS, help="corpus rows per GPU")
    ap.add_argument("--no-extra", action="store_true", help="headline only")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--only", default="", choices=["", "encode", "batched", "config4", "filtered", "clustered", "headline"],
                    help="profiling aid: run one leg only")
    ap.add_argument("--skip", default="", help="comma-separated extra legs to skip (encode, filtered, clustered, config4, "
                                               "sharded_handle, batched, fp32_sweep)")
    ap.add_argument("--config4-rows", type=int, default=12_500_000, help="rows per GPU of the config4 leg")
    ap.add_argument("--config4", action="store_true", help="run the config4 leg at any N (default: N = 8 only)")
    ap.add_argument("--encode-seqs", type=int, default=296,
                    help="chunks per encoder pass (296 x 384 tokens = 148 SMs x 768: every kernel's tile count is a multiple of the SM count)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)
    skip = set(s for s in args.skip.split(",") if s)

    import torch
    from claude_semantic_search_b200 import _native as native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    # a non-default stream: libcss_b200 treats stream 0 as "use the handle's own stream",
    # and the CUDA events below must sit on the stream the kernels are launched on
    torch.cuda.set_stream(torch.cuda.Stream(dev))
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=dev)
        dist = dist_mod
    pk = peaks()

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu and args.only in ("", "headline"):
        cpu_base = cpu_search_baseline()

    if args.only == "encode":
        from bench_encoder import bench_encoder
        print(json.dumps(bench_encoder(torch, dev, pk, world, rank, dist, args)))
        return
    if args.only == "config4":
        out = bench_config4(torch, native, dev, pk, world, rank, dist, args.config4_rows)
        if rank == 0:
            print(json.dumps(out))
        return
    if args.only == "filtered":
        print(json.dumps(bench_filtered(torch, native, dev, pk)))
        return
    if args.only == "clustered":
        print(json.dumps(bench_clustered(torch, native, dev, pk, args.rows)))
        return

    rows = args.rows
    idx = build_shard(torch, native, dev, rows, seed=42 + rank)
    gq = torch.Generator(device=dev)
    gq.manual_seed(43)
    nq_pool = 1024
    qs = torch.randn((nq_pool, D), generator=gq, device=dev, dtype=torch.float32)
    qs = qs / (qs.norm(dim=1, keepdim=True) + 1e-8)
    qs_host = qs.cpu().pin_memory()
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream
    if args.only == "batched":
        print(json.dumps(bench_batched(torch, native, dev, idx, qs, pk, world, rank, rows, dist)))
        idx.close()
        return

    from claude_semantic_search_b200.sharded import ShardedSearch
    id_offset = rank * rows
    sharded = ShardedSearch(idx, id_offset)
    two_phase = os.environ.get("CSS_SCAN_BF16", "1") != "0"
    int8_tier = two_phase and os.environ.get("CSS_SCAN_INT8", "1") != "0"

    # ---- in-run correctness, before any timing: merged top-k vs brute force, one filtered ----
    check = verify_headline(torch, native, dev, dist, world, rank, rows, idx, sharded, qs)

    D_loc = torch.empty((1, K), device=dev, dtype=torch.float32)
    I_loc = torch.empty((1, K), device=dev, dtype=torch.int64)

    def step(i):
        # local scan; for N > 1 the lists are exchanged and merged inside the scan kernel (NVLink peer stores)
        sharded.search_device(qs[i % nq_pool:i % nq_pool + 1], K)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = native.kernel_launch_count()
    ms_total, _ = time_region(torch, dev, step, args.steps, dist)
    n_launch = native.kernel_launch_count() - l0
    ms_total = max_over_ranks(torch, dev, dist, ms_total)
    ms_step = ms_total / args.steps
    qps = 1e3 / ms_step

    # per-query latency distribution (BASELINE's metric names p50): >= 500 queries whatever --steps is
    n_lat = max(500, args.steps)
    lat_dev = per_query_latency(torch, dev, step, n_lat, dist)

    # kernel-only duration of the local scan (no exchange), for the roofline
    def scan_only(i):
        idx.search_device(qs[i % nq_pool].data_ptr(), 1, K, D_loc.data_ptr(), I_loc.data_ptr(), 0, id_offset, sp)
    n_k = max(args.steps, 100)
    ms_scan, _ = time_region(torch, dev, scan_only, n_k, dist)
    ms_scan /= n_k
    # The default path is the two-phase exact scan: the dominant kernel sweeps the int8 shadow rows (768 + 4 B per row,
    # a quarter of SURVEY 8(d)'s 3072 B fp32 row; CSS_SCAN_INT8=0: the bf16 shadow rows, 1536 B), the fp32 rows are
    # touched only for the few re-scored candidates.  Its roofline is quoted on the bytes it has to read, timed alone
    # through css_debug_scan_int8 / css_debug_scan_bf16.
    if two_phase:
        def phase1_only(i):
            (idx.debug_scan_int8 if int8_tier else idx.debug_scan_bf16)(qs[i % nq_pool].data_ptr(), 1, sp)
        for i in range(5):
            phase1_only(i)
        ms_kernel, _ = time_region(torch, dev, phase1_only, n_k, dist)
        ms_kernel /= n_k
        kernel_bytes = rows * (INT8_ROW_BYTES if int8_tier else D * 2)
        kernel_name = (f"scan_topk_kernel<{'int8' if int8_tier else 'bf16'} shadow> (sweep of the two-phase exact scan, timed alone "
                       "without its last-CTA proof + fp32 re-score; the step adds that and one idle fp32-fallback launch)")
    else:
        ms_kernel, kernel_bytes, kernel_name = ms_scan, rows * BYTES_PER_ROW, "scan_topk_kernel (fp32 sweep)"
    achieved = kernel_bytes / (ms_kernel * 1e-3) / 1e9

    # ---- e2e: the plugin-facing call with HOST buffers (H2D of the query + D2H of D/I inside) ----
    qh = qs_host.numpy()

    def e2e_step(i):
        sharded.search_host(qh[i % nq_pool:i % nq_pool + 1], K)
    for i in range(args.warmup):
        e2e_step(i)
    _, wall_ms = time_region(torch, dev, e2e_step, args.steps, dist)
    wall_ms = max_over_ranks(torch, dev, dist, wall_ms)
    e2e_qps = args.steps / (wall_ms * 1e-3)
    if dist is not None:
        dist.barrier()
    lat_wall = wall_latency(e2e_step, n_lat)
    clocks = sampler.stop() if rank == 0 else None
    stats = idx.scan_stats()

    extra = {"in_run_check": check}
    if not args.no_extra and args.only == "":
        if "fp32_sweep" not in skip and two_phase:
            extra["fp32_sweep"] = bench_fp32_sweep(torch, native, dev, idx, qs, pk, rows, id_offset, sp, dist, D_loc, I_loc)
        if "bf16_tier" not in skip and int8_tier:
            extra["bf16_two_phase"] = bench_bf16_tier(torch, native, dev, idx, qs, pk, rows, id_offset, sp, dist, D_loc, I_loc)
        if "batched" not in skip:
            try:
                extra.update(bench_batched(torch, native, dev, idx, qs, pk, world, rank, rows, dist))
            except Exception as e:  # report, never hide
                extra["batch1024_error"] = repr(e)
    sharded.close()
    idx.close()
    torch.cuda.empty_cache()
    if not args.no_extra and args.only == "":
        if "clustered" not in skip and world == 1:
            try:
                extra.update(bench_clustered(torch, native, dev, pk, rows))
            except Exception as e:
                extra["clustered_error"] = repr(e)
        if "filtered" not in skip and world == 1:
            try:
                extra.update(bench_filtered(torch, native, dev, pk))
            except Exception as e:
                extra["filtered_error"] = repr(e)
        if "config4" not in skip and (world == 8 or args.config4):
            try:
                extra["config4"] = bench_config4(torch, native, dev, pk, world, rank, dist, args.config4_rows)
            except Exception as e:
                extra["config4_error"] = repr(e)
        if "sharded_handle" not in skip and world > 1:
            try:
                extra["sharded_handle"] = bench_sharded_handle(torch, native, dev, dist, world, rank, rows)
            except Exception as e:
                extra["sharded_handle_error"] = repr(e)
        if "encode" not in skip:
            try:
                from bench_encoder import bench_encoder
                extra.update(bench_encoder(torch, dev, pk, world, rank, dist, args))
            except ImportError:
                pass
            except Exception as e:
                extra["encode_error"] = repr(e)

    if rank == 0:
        UNIT = unit_for(rows)
        launches_per_step = n_launch / max(args.steps, 1)
        line = {
            "metric": METRIC, "value": qps * world, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(rows, world),
            "qps": qps,
            "latency_ms": {"device": lat_dev, "e2e": lat_wall},
            "path": {"scan": (f"two-phase exact scan ({'int8' if int8_tier else 'bf16'} shadow sweep; proof + fp32 re-score in the "
                              "sweep's last CTA; unproven queries re-run by the fp32 sweep)") if two_phase else "single fp32 sweep",
                     "exchange": "none" if world == 1 else "in-kernel: k x 16 B stored into every peer over NVLink (CUDA IPC), "
                                                          "flags awaited and lists merged by the last CTA of the scan; no NCCL, no merge launch",
                     "two_phase_queries": stats["two_phase_queries"], "unproven_queries": stats["unproven_queries"],
                     "max_bf16_error_norm": stats["max_bf16_error_norm"],
                     "max_int8_error_norm": stats["max_int8_error_norm"], "last_tier": stats["last_tier"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / pk["hbm_gbs"],
                         "traffic": (NCU_SCAN_TRAFFIC_INT8 if int8_tier else NCU_SCAN_TRAFFIC_BF16 if two_phase
                                     else NCU_SCAN_TRAFFIC).get(rows),
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture under profiles/",
                         "peak_source": pk["source"],
                         "kernel": kernel_name, "kernel_ms": ms_kernel, "step_ms_device": ms_scan,
                         "step_frac_of_peak": kernel_bytes / (ms_scan * 1e-3) / 1e9 / pk["hbm_gbs"],
                         "algorithmic_bytes_per_launch": kernel_bytes,
                         "fp32_row_definition": {"bytes_per_row": BYTES_PER_ROW,
                                                 "step_gbs": rows * BYTES_PER_ROW / (ms_scan * 1e-3) / 1e9,
                                                 "note": "SURVEY 8(d) counts 3072 B per row; the two-phase scan answers the same exact "
                                                         "query from 772 B per row (int8 shadow; extra.bf16_two_phase: 1536 B per row), "
                                                         "see extra.fp32_sweep for the kernel that streams the fp32 rows"}},
            "cpu_baseline": cpu_base,
            "e2e": {"value": e2e_qps * world, "unit": UNIT, "h2d_bytes_per_step": D * 4, "d2h_bytes_per_step": K * 12,
                    "qps": e2e_qps,
                    "api": "css_index_search (host q -> host D,I)" if world == 1 else
                           "ShardedSearch.search_host -> css_index_search_exchange (host q -> host D,I on every rank; result through mapped memory, in-kernel exchange)"},
            "gpu_launches": int(n_launch), "launches_per_step": launches_per_step,
            "clocks": clocks, "extra": extra,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


# --------------------------------------------------------------------------
def verify_headline(torch, native, dev, dist, world, rank, rows, idx, sharded, qs):
    """Merged top-k of 8 queries (device path and host path) and of a filtered query against a brute-force
    recomputation over regenerated tiles, all-gathered over the ranks.  Raises on a mismatch."""
    nqc = 8
    q = qs[:nqc].contiguous()
    sp = torch.cuda.current_stream(dev).cuda_stream
    ref_s, ref_i = brute_force_topk(torch, dev, gen_tiles(torch, dev, rows, 42 + rank), q, K + 8, id_offset=rank * rows)
    # filter: rows whose (local row % 10) in [3, 5] -- evaluated by the device filter kernel on every shard
    col = (torch.arange(rows, device=dev, dtype=torch.int64) % 10).to(torch.int32)
    idx.set_column(2, col.cpu().numpy())
    fmask = (col >= 3) & (col <= 5)
    reff_s, reff_i = brute_force_topk(torch, dev, gen_tiles(torch, dev, rows, 42 + rank), q[:2], K + 8,
                                      id_offset=rank * rows, mask=fmask)
    if world > 1:
        def gather(t):
            out = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(out, t.contiguous())
            return torch.cat(out, dim=1)
        ref_s, ref_i, reff_s, reff_i = gather(ref_s), gather(ref_i), gather(reff_s), gather(reff_i)

        def resort(s, i):
            o = torch.argsort(s, dim=1, descending=True, stable=True)
            return torch.gather(s, 1, o), torch.gather(i, 1, o)
        ref_s, ref_i = resort(ref_s, ref_i)
        reff_s, reff_i = resort(reff_s, reff_i)
    got = []
    for i in range(nqc):
        Dd, Id = sharded.search_device(q[i:i + 1], K)
        torch.cuda.synchronize(dev)
        got.append((Dd.cpu().numpy().copy(), Id.cpu().numpy().copy()))
    Dg = np.concatenate([g[0] for g in got])
    Ig = np.concatenate([g[1] for g in got])
    ok, why = check_topk(Dg, Ig, ref_s.cpu().numpy(), ref_i.cpu().numpy(), K)
    assert ok, f"rank {rank}: merged top-k differs from brute force: {why}"
    Dh, Ih = sharded.search_host(q[:4].cpu().numpy(), K)           # nq = 4 in one call, host buffers
    ok, why = check_topk(Dh, Ih, ref_s[:4].cpu().numpy(), ref_i[:4].cpu().numpy(), K)
    assert ok, f"rank {rank}: host path differs from brute force: {why}"
    mptr, _ = idx.filter_mask_device(native.Filter().add_range(2, 3, 5), sp)
    Df, If = sharded.search_device(q[:2], K, mask_ptr=mptr)
    torch.cuda.synchronize(dev)
    ok, why = check_topk(Df.cpu().numpy(), If.cpu().numpy(), reff_s.cpu().numpy(), reff_i.cpu().numpy(), K)
    assert ok, f"rank {rank}: filtered merged top-k differs from brute force: {why}"
    return {"queries": nqc, "host_queries": 4, "filtered_queries": 2, "n_gpus": world, "result": "merged top-10 == brute force "
            "(regenerated tiles, fp32 matmul + topk, all-gathered) within 1e-4 on every rank, incl. a filtered search"}


def bench_fp32_sweep(torch, native, dev, idx, qs, pk, rows, id_offset, sp, dist, D_loc, I_loc):
    """SURVEY 8(d)'s definition of the batch-1 work: 3072 B per fp32 row.  The kernel that streams exactly those
    bytes (also the path of k > 32, L2, d != 768 and unproven queries), timed in the same run."""
    native.set_option("scan_bf16", 0)
    try:
        def f(i):
            idx.search_device(qs[i % 1024].data_ptr(), 1, K, D_loc.data_ptr(), I_loc.data_ptr(), 0, id_offset, sp)
        for i in range(5):
            f(i)
        ms, _ = time_region(torch, dev, f, 200, dist)
        ms /= 200
    finally:
        native.set_option("scan_bf16", 1)
    gbs = rows * BYTES_PER_ROW / (ms * 1e-3) / 1e9
    return {"roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                         "traffic": NCU_SCAN_TRAFFIC.get(rows), "kernel": "scan_topk_kernel<fp32> (one launch per query)",
                         "kernel_ms": ms, "algorithmic_bytes_per_launch": rows * BYTES_PER_ROW},
            "qps": 1e3 / ms}


def bench_bf16_tier(torch, native, dev, idx, qs, pk, rows, id_offset, sp, dist, D_loc, I_loc):
    """The two-phase scan with the bf16 shadow rows as its sweep (the second tier: CSS_SCAN_INT8=0, and what the
    adaptive switch falls back to when the int8 tier cannot prove most queries), timed in the same run."""
    native.set_option("scan_int8", 0)
    try:
        def f(i):
            idx.search_device(qs[i % 1024].data_ptr(), 1, K, D_loc.data_ptr(), I_loc.data_ptr(), 0, id_offset, sp)
        for i in range(5):
            f(i)
        ms, _ = time_region(torch, dev, f, 200, dist)
        ms /= 200

        def g(i):
            idx.debug_scan_bf16(qs[i % 1024].data_ptr(), 1, sp)
        for i in range(5):
            g(i)
        msk, _ = time_region(torch, dev, g, 200, dist)
        msk /= 200
    finally:
        native.set_option("scan_int8", 1)
    gbs = rows * D * 2 / (msk * 1e-3) / 1e9
    return {"roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                         "traffic": NCU_SCAN_TRAFFIC_BF16.get(rows), "kernel": "scan_topk_kernel<bf16 shadow> (sweep timed alone)",
                         "kernel_ms": msk, "algorithmic_bytes_per_launch": rows * D * 2,
                         "step_ms_device": ms, "step_frac_of_peak": rows * D * 2 / (ms * 1e-3) / 1e9 / pk["hbm_gbs"]},
            "qps": 1e3 / ms}


def bench_clustered(torch, native, dev, pk, rows):
    """The headline search on a clustered corpus (2000 caps, intra-cluster cosine 0.6-0.95, members adjacent in
    row order): QPS, share of queries the two-phase proof could not close (they take the fp32 sweep), and an
    in-run brute-force check.  VERDICT r1 item 3."""
    from bench_data import clustered_torch_tiles, perturbed_queries_torch
    out = {}
    for order in ("session", "shuffled"):
        tiles = lambda: ((i * TILE, t) for i, t in enumerate(clustered_torch_tiles(torch, dev, rows, order=order, seed=7, tile=TILE)))
        first = next(iter(tiles()))[1]
        q = perturbed_queries_torch(torch, first, 512, seed=11).contiguous()
        # queries drawn from the first tile only would all hit the first clusters: spread them over the corpus
        idx = build_shard(torch, native, dev, rows, seed=0, tiles=tiles())
        mid = torch.empty((TILE, D), device=dev)
        idx_rows = native.Index  # noqa: F841 (keep flake quiet)
        # second query set from the middle of the corpus
        for i, (r0, t) in enumerate(tiles()):
            if i == (rows // TILE) // 2:
                mid = t
                break
        q = torch.cat([q[:256], perturbed_queries_torch(torch, mid, 256, seed=12)]).contiguous()
        sp = torch.cuda.current_stream(dev).cuda_stream
        Dd = torch.empty((1, K), device=dev, dtype=torch.float32)
        Id = torch.empty((1, K), device=dev, dtype=torch.int64)
        # correctness on 8 queries (4 from each end)
        sel = torch.cat([q[:4], q[256:260]]).contiguous()
        ref_s, ref_i = brute_force_topk(torch, dev, tiles(), sel, K + 8)
        got_D, got_I = [], []
        for i in range(8):
            idx.search_device(sel[i].data_ptr(), 1, K, Dd.data_ptr(), Id.data_ptr(), 0, 0, sp)
            torch.cuda.synchronize(dev)
            got_D.append(Dd.cpu().numpy().copy())
            got_I.append(Id.cpu().numpy().copy())
        ok, why = check_topk(np.concatenate(got_D), np.concatenate(got_I), ref_s.cpu().numpy(), ref_i.cpu().numpy(), K)
        assert ok, f"clustered ({order}): {why}"
        s0 = idx.scan_stats()

        def f(i):
            idx.search_device(q[i % 512].data_ptr(), 1, K, Dd.data_ptr(), Id.data_ptr(), 0, 0, sp)
        for i in range(10):
            f(i)
        ms, _ = time_region(torch, dev, f, 512)
        ms /= 512
        s1 = idx.scan_stats()
        asked = s1["two_phase_queries"] - s0["two_phase_queries"]
        unproven = s1["unproven_queries"] - s0["unproven_queries"]
        qh = q.cpu().numpy()
        for i in range(5):
            idx.search(qh[i:i + 1], K)
        t0 = time.perf_counter()
        for i in range(256):
            idx.search(qh[i:i + 1], K)
        e2e_ms = (time.perf_counter() - t0) / 256 * 1e3
        # batch-1024 on the same corpus
        qb = torch.cat([q, q.flip(0)]).contiguous()
        Db = torch.empty((1024, K), device=dev, dtype=torch.float32)
        Ib = torch.empty((1024, K), device=dev, dtype=torch.int64)
        for _ in range(2):
            idx.search_device(qb.data_ptr(), 1024, K, Db.data_ptr(), Ib.data_ptr(), 0, 0, sp)
        msb, _ = time_region(torch, dev, lambda i: idx.search_device(qb.data_ptr(), 1024, K, Db.data_ptr(), Ib.data_ptr(), 0, 0, sp), 5)
        msb /= 5
        torch.cuda.synchronize(dev)
        # the batched result of the 8 checked queries must equal the batch-1 result bit for bit
        same = bool((Ib[:4].cpu().numpy() == np.concatenate(got_I)[:4]).all() and (Db[:4].cpu().numpy() == np.concatenate(got_D)[:4]).all())
        assert same, f"clustered ({order}): batched and batch-1 results differ"
        out[order] = {"qps": 1e3 / ms, "ms_per_query": ms, "e2e_qps": 1e3 / e2e_ms,
                      "two_phase_queries": asked, "unproven_queries": unproven,
                      "fallback_rate": (unproven / asked) if asked else None,
                      "last_tier": s1["last_tier"],
                      "step_frac_of_hbm_peak_on_int8_bytes": rows * INT8_ROW_BYTES / (ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                      "step_frac_of_hbm_peak_on_bf16_bytes": rows * D * 2 / (ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                      "batch1024_ms": msb, "batch1024_tflops": 2.0 * 1024 * rows * D / (msb * 1e-3) / 1e12,
                      "check": "8 queries == brute force within 1e-4; batched == batch-1 bit for bit"}
        idx.close()
        torch.cuda.empty_cache()
    return {"clustered_1M": dict(out, corpus=f"{rows} x {D}, 2000 clusters, cos(row, centre) ~ U[0.6, 0.95], queries = perturbed rows; "
                                            "'session' = cluster members adjacent in row order, 'shuffled' = dealt at random")}


def bench_filtered(torch, native, dev, pk, rows=10_000_000, n_queries=300):
    """BASELINE configs[4]: date range AND project AND has_code (~5 % selectivity) over 10M x 768, batch-1, p50
    latency of css_index_search (host query in, host result out; the filter is compiled to clauses, evaluated on
    the device into a row bitmask, and the scan skips masked rows).  The device mask is compared bit for bit
    with a numpy evaluation of the same predicate over all 10M rows, and 4 results with brute force."""
    idx = build_shard(torch, native, dev, rows, seed=42)
    rng = np.random.default_rng(99)
    ts = rng.integers(0, 731, size=rows).astype(np.int32)                  # day rank 2023-01-01 .. 2024-12-31
    zipf = 1.0 / np.arange(1, 201) ** 1.1
    proj = rng.choice(200, size=rows, p=zipf / zipf.sum()).astype(np.int32)
    has_code = (rng.random(rows) < 0.4).astype(np.int32)
    idx.set_column(4, ts)
    idx.set_column(1, proj)
    idx.set_column(5, has_code)
    # project "substring" -> allowed id set covering ~50 % of rows; date window ~25 %
    order = rng.permutation(200)
    mass = np.bincount(proj, minlength=200) / rows
    allowed, acc = [], 0.0
    for p_ in order:                      # ~50 % of the rows, never overshooting by a heavy project
        if acc + mass[p_] <= 0.505:
            allowed.append(int(p_))
            acc += mass[p_]
    # date window sized so that window x project x has_code = 5.0 % (5.0 +- 0.2 %)
    days = int(round(0.05 / (acc * float(has_code.mean())) * 731))
    flt = native.Filter().add_range(4, 100, 100 + days - 1).add_set(1, allowed, 200).add_range(5, 1, 1)
    words, n_pass = idx.filter_mask(flt)
    sel = n_pass / rows
    want = (ts >= 100) & (ts <= 100 + days - 1) & np.isin(proj, np.asarray(allowed, np.int32)) & (has_code == 1)
    padded = np.zeros((rows + 31) // 32 * 32, np.uint8)
    padded[:rows] = want
    want_words = np.packbits(padded.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).reshape(-1)
    mask_exact = bool(np.array_equal(words, want_words)) and int(want.sum()) == n_pass
    assert mask_exact, "device filter mask differs from the numpy evaluation of the predicate"
    q = np.random.default_rng(43).standard_normal((n_queries, D)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True) + 1e-8
    qd = torch.from_numpy(q[:4]).to(dev)
    ref_s, ref_i = brute_force_topk(torch, dev, gen_tiles(torch, dev, rows, 42), qd, K + 8, mask=torch.from_numpy(want).to(dev))
    Dc, Ic = idx.search(q[:4], K, flt)
    ok, why = check_topk(Dc, Ic, ref_s.cpu().numpy(), ref_i.cpu().numpy(), K)
    assert ok, f"filtered 10M: {why}"
    # every query with a filter the index has not just evaluated (two windows alternate: clauses compiled, staged,
    # evaluated over all 10M rows, then the scan) -- the p50 BASELINE configs[4] asks for ...
    flt_b = native.Filter().add_range(4, 101, 101 + days - 1).add_set(1, allowed, 200).add_range(5, 1, 1)
    for i in range(6):
        idx.search(q[i:i + 1], K, flt if i % 2 == 0 else flt_b)
    lat = []
    for i in range(n_queries):
        f = flt if i % 2 == 0 else flt_b
        t0 = time.perf_counter()
        idx.search(q[i:i + 1], K, f)
        lat.append((time.perf_counter() - t0) * 1e3)
    # ... and with the same filter as the previous query (its mask is still on the device: evaluation skipped)
    lat_same = []
    for i in range(n_queries):
        t0 = time.perf_counter()
        idx.search(q[i:i + 1], K, flt)
        lat_same.append((time.perf_counter() - t0) * 1e3)
    lat_u = []
    for i in range(50):
        t0 = time.perf_counter()
        idx.search(q[i:i + 1], K)
        lat_u.append((time.perf_counter() - t0) * 1e3)
    # device-only time of the filtered scan with the mask already evaluated (what an unchanged filter costs per query)
    sp = torch.cuda.current_stream(dev).cuda_stream
    mptr, _ = idx.filter_mask_device(flt, sp)
    Dd = torch.empty((1, K), device=dev, dtype=torch.float32)
    Id = torch.empty((1, K), device=dev, dtype=torch.int64)
    qall = torch.from_numpy(q).to(dev)
    def scan_step(i):
        idx.search_device(qall[i % n_queries].data_ptr(), 1, K, Dd.data_ptr(), Id.data_ptr(), mptr, 0, sp)
    for i in range(5):   # the first search on a stream allocates that stream's scratch: not part of the scan
        scan_step(i)
    ms_scan, _ = time_region(torch, dev, scan_step, 200)
    ms_scan /= 200
    idx.close()
    torch.cuda.empty_cache()
    p50 = float(np.median(lat))
    dense = rows * BYTES_PER_ROW + rows / 8
    selective = sel * rows * BYTES_PER_ROW + rows / 8 + 3 * 4 * rows   # + the 3 int32 columns the predicate reads
    selective_bf16 = sel * rows * D * 2 + rows / 8 + 3 * 4 * rows
    selective_int8 = sel * rows * INT8_ROW_BYTES + rows / 8 + 3 * 4 * rows
    return {"filtered_10M": {"rows": rows, "selectivity": sel, "p50_ms": p50, "p99_ms": float(np.percentile(lat, 99)),
                             "repeated_filter_p50_ms": float(np.median(lat_same)), "repeated_filter_p99_ms": float(np.percentile(lat_same, 99)),
                             "unfiltered_p50_ms": float(np.median(lat_u)), "device_scan_ms_mask_cached": ms_scan,
                             "mask_bit_exact_vs_numpy_all_rows": mask_exact,
                             "check": "4 filtered results == brute force within 1e-4",
                             "dense_roofline_ms": dense / (pk["hbm_gbs"] * 1e9) * 1e3,
                             "selective_roofline_ms": selective / (pk["hbm_gbs"] * 1e9) * 1e3,
                             "selective_bf16_roofline_ms": selective_bf16 / (pk["hbm_gbs"] * 1e9) * 1e3,
                             "selective_int8_roofline_ms": selective_int8 / (pk["hbm_gbs"] * 1e9) * 1e3,
                             "device_scan_frac_of_hbm_peak_selective_int8_denominator":
                                 (sel * rows * INT8_ROW_BYTES + rows / 8) / (ms_scan * 1e-3) / 1e9 / pk["hbm_gbs"],
                             "achieved_gbs_dense_denominator": dense / (p50 * 1e-3) / 1e9,
                             "achieved_gbs_selective_denominator": selective / (p50 * 1e-3) / 1e9,
                             "frac_of_hbm_peak_selective_bf16_denominator": selective_bf16 / (p50 * 1e-3) / 1e9 / pk["hbm_gbs"],
                             "device_scan_frac_of_hbm_peak_selective_bf16_denominator": selective_bf16 / (ms_scan * 1e-3) / 1e9 / pk["hbm_gbs"],
                             "api": "css_index_search with css_filter (3 clauses), host buffers"}}


def bench_batched(torch, native, dev, idx, qs, pk, world, rank, rows, dist):
    """Batch-1024 exact top-10 on the same shard (tensor-core bound)."""
    nq = 1024
    sp = torch.cuda.current_stream(dev).cuda_stream
    Db = torch.empty((nq, K), device=dev, dtype=torch.float32)
    Ib = torch.empty((nq, K), device=dev, dtype=torch.int64)

    def step(i):
        idx.search_device(qs.data_ptr(), nq, K, Db.data_ptr(), Ib.data_ptr(), 0, rank * rows, sp)
    for _ in range(3):
        step(0)
    ms, _ = time_region(torch, dev, step, 10, dist)
    ms /= 10
    flops = 2.0 * nq * rows * D
    tf = flops / (ms * 1e-3) / 1e12
    return {"batch1024": {"qps": nq / (ms * 1e-3), "ms_per_call": ms, "achieved_tflops": tf,
                          "frac_of_bf16_peak": tf / pk["bf16_tflops"], "peak": pk["bf16_tflops"],
                          "algorithmic_flops": flops}}


def bench_config4(torch, native, dev, pk, world, rank, dist, rows):
    """BASELINE configs[3]: exact top-10 over world x rows x 768 fp32 (100M on 8 GPUs: 12.5M rows = 38.4 GB fp32 +
    19.2 GB bf16 per GPU), row-sharded, batch-1 (two-phase scan and fp32 sweep, in-kernel exchange) and batch-1024
    (tensor cores, NCCL gather of the lists).  Verification inside the run: for 64 queries, 10 rows spread over ALL
    shards are overwritten with normalise(q + sigma_r noise) at decreasing similarity, so the exact global top-10
    (ids and order) is known a priori; both paths must return exactly those, and agree on needle-free queries."""
    from claude_semantic_search_b200.sharded import ShardedSearch
    N = rows * world
    gq = torch.Generator(device="cpu").manual_seed(43)
    q = torch.randn((1024, D), generator=gq)
    q = (q / q.norm(dim=1, keepdim=True)).to(dev)
    rng = np.random.default_rng(7)
    n_needle_q = 64
    needle_ids = np.sort(rng.choice(N, size=(n_needle_q, K), replace=False), axis=1)
    rng.shuffle(needle_ids, axis=1)
    noise = torch.randn((n_needle_q, K, D), generator=gq)
    noise = noise / noise.norm(dim=2, keepdim=True)
    sig = (0.05 + 0.05 * torch.arange(K)).view(1, K, 1)
    needles = q[:n_needle_q].cpu().unsqueeze(1) + sig * noise
    needles = needles / needles.norm(dim=2, keepdim=True)          # [64, 10, 768], similarity decreasing in r
    lo = rank * rows

    def plant(r0, blk):
        sel = (needle_ids >= lo + r0) & (needle_ids < lo + r0 + blk.shape[0])
        for qi, r in zip(*np.nonzero(sel)):
            blk[int(needle_ids[qi, r]) - lo - r0] = needles[qi, r].to(dev) * 3.0   # any scale: add() normalises

    t0 = time.perf_counter()
    idx = build_shard(torch, native, dev, rows, seed=42 + rank, plant=plant)
    build_s = time.perf_counter() - t0
    ss = ShardedSearch(idx, id_offset=lo)

    def timed(fn, n):
        ms, _ = time_region(torch, dev, fn, n, dist)
        return max_over_ranks(torch, dev, dist, ms) / n

    def batch1_ids(lo_q, hi_q):
        out = []
        for i in range(lo_q, hi_q):
            _, Id = ss.search_device(q[i:i + 1], K)
            torch.cuda.synchronize(dev)
            out.append(Id.cpu().numpy()[0].copy())
        return np.stack(out)

    res = {}
    I1 = batch1_ids(0, n_needle_q)
    ok_scan = bool((I1 == needle_ids).all())
    Db, Ib = ss.search_device(q, K)
    torch.cuda.synchronize(dev)
    Ib = Ib.cpu().numpy().copy()
    ok_batched = bool((Ib[:n_needle_q] == needle_ids).all())
    I2 = batch1_ids(n_needle_q, n_needle_q + 32)
    paths_agree = bool((I2 == Ib[n_needle_q:n_needle_q + 32]).all())
    native.set_option("scan_bf16", 0)
    try:
        I1f = batch1_ids(0, n_needle_q)
        ok_fp32 = bool((I1f == needle_ids).all())
        for i in range(3):
            ss.search_device(q[i:i + 1], K)
        ms_f = timed(lambda i: ss.search_device(q[i % 1024:i % 1024 + 1], K), 60)
    finally:
        native.set_option("scan_bf16", 1)
    # second tier: bf16 shadow sweep (needles again, then timing)
    int8_tier = os.environ.get("CSS_SCAN_BF16", "1") != "0" and os.environ.get("CSS_SCAN_INT8", "1") != "0"
    ok_bf16, ms_b16 = None, None
    if int8_tier:
        native.set_option("scan_int8", 0)
        try:
            ok_bf16 = bool((batch1_ids(0, n_needle_q) == needle_ids).all())
            for i in range(3):
                ss.search_device(q[i:i + 1], K)
            ms_b16 = timed(lambda i: ss.search_device(q[i % 1024:i % 1024 + 1], K), 100)
        finally:
            native.set_option("scan_int8", 1)
    row_bytes = INT8_ROW_BYTES if int8_tier else D * 2
    for i in range(5):
        ss.search_device(q[i:i + 1], K)
    ms1 = timed(lambda i: ss.search_device(q[i % 1024:i % 1024 + 1], K), 200)
    lat = per_query_latency(torch, dev, lambda i: ss.search_device(q[i % 1024:i % 1024 + 1], K), 300, dist)
    qh = q.cpu().numpy()
    for i in range(3):
        ss.search_host(qh[i:i + 1], K)
    _, wall = time_region(torch, dev, lambda i: ss.search_host(qh[i % 1024:i % 1024 + 1], K), 100, dist)
    e2e_ms = max_over_ranks(torch, dev, dist, wall) / 100
    ss.search_device(q, K)
    msb = timed(lambda i: ss.search_device(q, K), 3)
    stats = idx.scan_stats()
    hbm = pk["hbm_gbs"]
    res = {
        "workload": f"exact top-10 over {N} x {D} fp32 rows, row-sharded over {world} GPU(s) ({rows} rows per GPU)",
        "rows_per_gpu": rows, "n_gpus": world, "build_s": build_s,
        "batch1": {"ms_per_query": ms1, "qps": 1e3 / ms1, "p50_ms": lat["p50_ms"], "p99_ms": lat["p99_ms"], "e2e_ms_per_query": e2e_ms,
                   "path": f"two-phase exact scan ({'int8' if int8_tier else 'bf16'} shadow sweep, {row_bytes} B per row) + in-kernel exchange",
                   "per_gpu_gbs_on_bytes_read": rows * row_bytes / (ms1 * 1e-3) / 1e9,
                   "frac_of_hbm_peak_on_bytes_read": rows * row_bytes / (ms1 * 1e-3) / 1e9 / hbm,
                   "aggregate_gbs_on_bytes_read": world * rows * row_bytes / (ms1 * 1e-3) / 1e9,
                   "last_tier": stats["last_tier"],
                   "per_gpu_gbs_fp32_row_definition": rows * BYTES_PER_ROW / (ms1 * 1e-3) / 1e9,
                   "unproven_queries": stats["unproven_queries"], "two_phase_queries": stats["two_phase_queries"]},
        "batch1_bf16_tier": None if ms_b16 is None else {
            "ms_per_query": ms_b16, "qps": 1e3 / ms_b16, "per_gpu_gbs_on_bytes_read": rows * D * 2 / (ms_b16 * 1e-3) / 1e9,
            "frac_of_hbm_peak_on_bytes_read": rows * D * 2 / (ms_b16 * 1e-3) / 1e9 / hbm, "needles_exact": ok_bf16},
        "batch1_fp32_sweep": {"ms_per_query": ms_f, "qps": 1e3 / ms_f, "per_gpu_gbs": rows * BYTES_PER_ROW / (ms_f * 1e-3) / 1e9,
                              "frac_of_hbm_peak": rows * BYTES_PER_ROW / (ms_f * 1e-3) / 1e9 / hbm,
                              "aggregate_gbs": world * rows * BYTES_PER_ROW / (ms_f * 1e-3) / 1e9},
        "batch1024": {"ms_per_call": msb, "qps": 1024 / (msb * 1e-3),
                      "tflops_per_gpu": 2.0 * 1024 * rows * D / (msb * 1e-3) / 1e12,
                      "frac_of_bf16_peak": 2.0 * 1024 * rows * D / (msb * 1e-3) / 1e12 / pk["bf16_tflops"]},
        "needles_exact_two_phase": ok_scan, "needles_exact_fp32_sweep": ok_fp32, "needles_exact_batched": ok_batched,
        "scan_and_batched_paths_agree_32_queries": paths_agree}
    assert ok_scan and ok_fp32 and ok_batched and paths_agree, (ok_scan, ok_fp32, ok_batched, paths_agree)
    ss.close()
    idx.close()
    torch.cuda.empty_cache()
    return res


def bench_sharded_handle(torch, native, dev, dist, world, rank, rows):
    """The single-process multi-device index (css_index_create_sharded: what HybridStorage uses with
    StorageConfig.devices): rank 0 builds ONE index over all `world` GPUs (rows per GPU as in the headline) and
    times css_index_search with host buffers; the other ranks wait.  Checked against brute force."""
    out = None
    # the other ranks wait on the HOST (a store key), not in an NCCL barrier: a spinning barrier kernel of another
    # process on a GPU this index uses would time-slice with the index's kernels (2 ms slices)
    store = None
    if dist is not None:
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
        from torch.distributed.distributed_c10d import _get_default_store
        store = _get_default_store()
    if rank == 0:
        devs = list(range(world))
        idx = native.Index(D, native.METRIC_INNER_PRODUCT, devices=devs)
        total = rows * world
        idx.reserve(total)
        g = torch.Generator(device=dev).manual_seed(4242)
        gq = torch.Generator(device=dev).manual_seed(43)
        q = torch.randn((256, D), generator=gq, device=dev)
        q = (q / q.norm(dim=1, keepdim=True)).contiguous()
        best_s = best_i = None
        t0 = time.perf_counter()
        for r0 in range(0, total, TILE):
            blk = torch.randn((min(TILE, total - r0), D), generator=g, device=dev)
            blk = blk / (blk.norm(dim=1, keepdim=True) + 1e-8)
            idx.add(blk.cpu().numpy(), normalize=False)
            s = q[:8] @ blk.T
            ts, ti = torch.topk(s, K + 8, dim=1)
            ti = ti + r0
            if best_s is not None:
                ts, ti = torch.cat([best_s, ts], 1), torch.cat([best_i, ti], 1)
                ts, sel = torch.topk(ts, K + 8, dim=1)
                ti = torch.gather(ti, 1, sel)
            best_s, best_i = ts, ti
        build_s = time.perf_counter() - t0
        qh = q.cpu().numpy()
        Dg, Ig = [], []
        for i in range(8):
            d_, i_ = idx.search(qh[i:i + 1], K)
            Dg.append(d_)
            Ig.append(i_)
        ok, why = check_topk(np.concatenate(Dg), np.concatenate(Ig), best_s.cpu().numpy(), best_i.cpu().numpy(), K)
        assert ok, f"sharded handle: {why}"
        for i in range(20):
            idx.search(qh[i:i + 1], K)
        lat = wall_latency(lambda i: idx.search(qh[i % 256:i % 256 + 1], K), 500)
        t0 = time.perf_counter()
        idx.search(qh, K)
        batch_ms = (time.perf_counter() - t0) * 1e3
        idx.close()
        out = {"n_devices": world, "rows": total, "build_s": build_s, "batch1_e2e": lat, "batch1_e2e_qps": 1e3 / lat["mean_ms"],
               "shard_scans_per_s": world * 1e3 / lat["mean_ms"], "batch256_e2e_ms": batch_ms,
               "check": "8 queries == brute force within 1e-4",
               "api": "css_index_create_sharded + css_index_search (one process, host buffers, in-kernel NVLink exchange)"}
        if store is not None:
            store.set("css_sharded_handle_done", "1")
    elif store is not None:
        store.wait(["css_sharded_handle_done"])
    if dist is not None:
        dist.barrier()
    return out


if __name__ == "__main__":
    main()
