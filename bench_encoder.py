"""Encoder leg of bench.py (BASELINE configs[2]: MPNet encode at seq len 384, bf16,
data-parallel over the ranks with no collective)."""
from __future__ import annotations

import os
import time

import numpy as np

SEQ_LEN = 384
FLOP_PER_CHUNK = 12 * (24 * SEQ_LEN * 768 ** 2 + 4 * SEQ_LEN ** 2 * 768)  # 70.67 GFLOP


def synthetic_batch(n_seq: int, seed: int):
    rng = np.random.default_rng(seed)
    ids = rng.integers(4, 30526, size=(n_seq, SEQ_LEN), dtype=np.int64).astype(np.int32)
    ids[:, 0] = 0
    ids[:, -1] = 2
    cu = (np.arange(n_seq + 1, dtype=np.int64) * SEQ_LEN).astype(np.int32)
    return np.ascontiguousarray(ids.reshape(-1)), cu


def cpu_encoder_baseline(n_seq: int = 16, device: int = None):
    """The reference's CPU path for this half: transformers MPNetModel fp32 + the restated
    sentence-transformers pooling (oracle/encoder_oracle.py) on a bounded sample.  With `device` the same 16
    chunks (384 tokens, 12 layers, the prescribed random-init weights) also go through the CUDA encoder and the
    embeddings are compared: the benchmarked configuration is checked against the oracle inside the bench run."""
    import torch
    from oracle import encoder_oracle as eo
    torch.set_num_threads(os.cpu_count() or 1)
    model = eo.build_model(seed=0)
    seqs = eo.synthetic_ids(n_seq, [SEQ_LEN], seed=7)
    eo.st_encode_ids(model, seqs[:2])
    t0 = time.perf_counter()
    want = eo.st_encode_ids(model, seqs, batch_size=16)
    dt = time.perf_counter() - t0
    out = {"value": n_seq / dt, "unit": "chunks/s", "cores": os.cpu_count(), "kind": "port",
           "sample": f"{n_seq} chunks x {SEQ_LEN} tokens, batch 16, fp32 transformers MPNetModel + ST pooling on the host"}
    if device is not None:
        from claude_semantic_search_b200.encoder import MPNetEncoder
        enc = MPNetEncoder.from_hf_model(model, device=device, max_tokens=n_seq * SEQ_LEN)
        got = enc.encode_ids(seqs)
        enc.close()
        cos = eo.cosine_rows(want, got)
        out["parity"] = {"chunks": n_seq, "min_cosine_vs_fp32_cpu": float(cos.min()), "bar": 0.9999,
                         "max_abs_component_diff": float(np.abs(want - got).max())}
        assert cos.min() >= 0.9999, f"encoder parity at the benchmarked shape: min cosine {cos.min():.6f}"
    return out


def bench_text_to_embedding(enc, n_seq: int, reps: int = 3):
    """Text in, embedding out: native WordPiece tokenizer (css_tokenizer_encode_batch, synthetic 30527-entry
    vocabulary written to a temporary vocab.txt) -> packed ids -> css_encoder_encode, all on host buffers."""
    import random
    import tempfile
    from pathlib import Path

    from claude_semantic_search_b200.st_compat import NativeWordPieceTokenizer
    rnd = random.Random(17)
    letters = "abcdefghijklmnopqrstuvwxyz"
    words = sorted({"".join(rnd.choice(letters) for _ in range(rnd.randint(2, 9))) for _ in range(24000)})
    vocab = ["<s>", "<pad>", "</s>", "[UNK]"] + words + ["##" + w[:4] for w in words[:6000]] + list(letters) + \
        ["##" + c for c in letters] + list(".,!?()-:;")
    vocab = list(dict.fromkeys(vocab))[:30527]
    with tempfile.TemporaryDirectory() as d:
        vf = Path(d) / "vocab.txt"
        vf.write_text("\n".join(vocab) + "\n", encoding="utf-8")
        tok = NativeWordPieceTokenizer(vf)
        texts = [" ".join(rnd.choice(words) + rnd.choice(["", "", "", ",", "."]) for _ in range(400)) for _ in range(n_seq)]
        tok.encode_packed(texts[:8], SEQ_LEN)
        t_tok = t_all = 0.0
        for _ in range(reps):
            t0 = time.perf_counter()
            ids, cu = tok.encode_packed(texts, SEQ_LEN)
            t1 = time.perf_counter()
            emb = enc.encode_packed(ids, cu)
            t2 = time.perf_counter()
            t_tok += t1 - t0
            t_all += t2 - t0
        # what SentenceTransformer.encode does for a whole indexing batch: tokeniser one slab ahead of the GPU
        from claude_semantic_search_b200.st_compat import encode_texts_pipelined
        many = texts * max(1, 4096 // n_seq)
        encode_texts_pipelined(tok, enc, many[:2048], SEQ_LEN, True)
        t0 = time.perf_counter()
        emb_p = encode_texts_pipelined(tok, enc, many, SEQ_LEN, True)
        t_pipe = time.perf_counter() - t0
        tok.close()
    assert emb.shape == (n_seq, 768) and int(cu[-1]) == n_seq * SEQ_LEN
    assert np.array_equal(emb_p[:n_seq], emb)
    return {"chunks_per_s": n_seq * reps / t_all, "tokenizer_chunks_per_s": n_seq * reps / t_tok,
            "pipelined_chunks_per_s": len(many) / t_pipe, "pipelined_chunks": len(many),
            "tokenizer_threads": "hardware_concurrency, at most 16", "chars_per_chunk": int(np.mean([len(t) for t in texts])),
            "api": "css_tokenizer_encode_batch + css_encoder_encode (host text -> host embeddings, tokenise and encode not overlapped)"}


def bench_single_query(enc, reps: int = 300):
    """p50 / p99 latency of one short query through css_encoder_encode (host ids in, host
    embedding out): the CUDA-graph path over the 32-row bucket (query_kernels.cuh)."""
    rng = np.random.default_rng(3)
    out = {}
    for L in (16, 48):
        qs = [[0] + rng.integers(4, 30000, size=L - 2).tolist() + [2] for _ in range(reps)]
        for q in qs[:20]:
            enc.encode_ids([q])
        lat = []
        for q in qs:
            t0 = time.perf_counter()
            enc.encode_ids([q])
            lat.append(time.perf_counter() - t0)
        out[f"L{L}"] = {"p50_ms": float(np.percentile(lat, 50) * 1e3), "p99_ms": float(np.percentile(lat, 99) * 1e3)}
    out["api"] = "css_encoder_encode, one sequence (host ids -> host embedding), wall clock"
    return out


def bench_multi_device_handle(n_dev: int, n_seq: int, ids, cu, emb_one, reps: int = 5):
    """north_star: embedding batches split data-parallel over the GPUs of one box behind the unchanged API.  One
    process, one MultiDeviceEncoder (EmbeddingConfig.devices) over `n_dev` GPUs, host ids in -> host embeddings out for
    n_dev x n_seq chunks per call; the rows must equal the single-device result bit for bit."""
    from claude_semantic_search_b200.encoder import MPNetEncoder, MultiDeviceEncoder, random_state_dict
    sd = random_state_dict(0)
    enc = MultiDeviceEncoder.create(lambda d: MPNetEncoder(sd, device=d, max_tokens=n_seq * SEQ_LEN), list(range(n_dev)))
    big_ids = np.tile(ids, n_dev)
    big_cu = (np.arange(n_dev * n_seq + 1, dtype=np.int64) * SEQ_LEN).astype(np.int32)
    got = enc.encode_packed(big_ids, big_cu)           # sizes the staging buffers of every device
    same = all(np.array_equal(got[r * n_seq:(r + 1) * n_seq], emb_one) for r in range(n_dev))
    t0 = time.perf_counter()
    for _ in range(reps):
        enc.encode_packed(big_ids, big_cu)
    dt = (time.perf_counter() - t0) / reps
    enc.close()
    assert same, "multi-device encode differs from the single-device result"
    return {"n_devices": n_dev, "chunks_per_call": n_dev * n_seq, "e2e_chunks_per_s": n_dev * n_seq / dt,
            "check": "every device's rows == the single-device embeddings bit for bit",
            "api": "MultiDeviceEncoder.encode_packed behind EmbeddingConfig.devices (one process, host ids -> host embeddings, "
                   "one css_encoder per GPU, one host thread per GPU, no collective)"}


def bench_encoder(torch, dev, pk, world, rank, dist, args, steps: int = 20, warmup: int = 3):
    from claude_semantic_search_b200 import _native as native
    from claude_semantic_search_b200.encoder import MPNetEncoder, random_state_dict
    n_seq = int(getattr(args, "encode_seqs", 256))
    T = n_seq * SEQ_LEN
    enc = MPNetEncoder(random_state_dict(0), device=dev.index, max_tokens=T)
    ids, cu = synthetic_batch(n_seq, seed=7 + rank)
    ids_d = torch.from_numpy(ids).to(dev)
    cu_d = torch.from_numpy(cu).to(dev)
    out_d = torch.empty((n_seq, 768), device=dev, dtype=torch.float32)
    sp = torch.cuda.current_stream(dev).cuda_stream

    def step(_i):
        enc.encode_device(ids_d.data_ptr(), cu_d.data_ptr(), cu, out_d.data_ptr(), True, sp)
    for i in range(warmup):
        step(i)
    torch.cuda.synchronize(dev)
    from bench import ClockSampler
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    l0 = native.kernel_launch_count()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = native.kernel_launch_count() - l0
    if dist is not None:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    chunks_s = n_seq * steps / (ms * 1e-3) * world
    tf = chunks_s / world * FLOP_PER_CHUNK / 1e12
    # e2e: host ids -> host embeddings through css_encoder_encode
    emb = enc.encode_packed(ids, cu)   # first call sizes the pinned staging buffers
    t0 = time.perf_counter()
    for _ in range(3):
        emb = enc.encode_packed(ids, cu)
    t_e2e = (time.perf_counter() - t0) / 3
    assert np.isfinite(emb).all() and abs(float(np.linalg.norm(emb[0])) - 1) < 1e-3
    text_leg = None
    if rank == 0:
        try:
            text_leg = bench_text_to_embedding(enc, n_seq)
        except Exception as e:  # report, never hide
            text_leg = {"error": repr(e)}
    query_leg = None
    if rank == 0:
        try:
            query_leg = bench_single_query(enc)
        except Exception as e:  # report, never hide
            query_leg = {"error": repr(e)}
    enc.close()
    multi_leg = None
    if world > 1:
        # the single-process multi-device handle (EmbeddingConfig.devices): rank 0 drives all GPUs of the job while the
        # other ranks sleep on a flag file (an NCCL barrier would park a spinning kernel on the GPUs under test)
        import tempfile
        from pathlib import Path
        flag = Path(tempfile.gettempdir()) / f"css_b200_multi_encoder_{os.environ.get('MASTER_PORT', '0')}.done"
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)
        if rank == 0:
            try:
                multi_leg = bench_multi_device_handle(world, n_seq, ids, cu, emb)
            except Exception as e:  # report, never hide
                multi_leg = {"error": repr(e)}
            flag.write_text("done")
        else:
            t_wait = time.time()
            while not flag.exists() and time.time() - t_wait < 600:
                time.sleep(0.05)
        if dist is not None:
            dist.barrier()
        if rank == 0:
            flag.unlink(missing_ok=True)
    res = {"encode": {"chunks_per_s": chunks_s, "ms_per_step": ms / steps, "chunks_per_step_per_gpu": n_seq,
                      "seq_len": SEQ_LEN, "achieved_tflops_per_gpu": tf,
                      "frac_of_bf16_peak": tf / pk["bf16_tflops_sustained"], "peak": pk["bf16_tflops_sustained"],
                      "peak_kind": "sustained bf16 (MEASURED_PEAKS.json)", "flop_per_chunk": FLOP_PER_CHUNK,
                      "gpu_launches": int(launches), "clocks": clocks,
                      "e2e_chunks_per_s": n_seq / t_e2e * world, "h2d_bytes_per_step": int(ids.nbytes + cu.nbytes),
                      "d2h_bytes_per_step": int(emb.nbytes)}}
    if text_leg is not None:
        res["encode"]["e2e_from_text"] = text_leg
    if query_leg is not None:
        res["encode"]["single_query"] = query_leg
    if multi_leg is not None:
        res["encode"]["multi_device_handle"] = multi_leg
    if rank == 0 and not getattr(args, "no_cpu", False):
        res["encode"]["cpu_baseline"] = cpu_encoder_baseline(device=dev.index)
    return res
