"""GPU parity of the MPNet chunk encoder (half A) against the fp32 CPU oracle
(transformers MPNetModel + restated sentence-transformers pooling).

north_star tolerance: cosine >= 0.9999 per chunk against fp32 CPU.  The kernel-level
tests (GEMM, attention) use tighter, bf16-rounding-derived bounds so a layout bug cannot
hide behind the loose end-to-end bar.
"""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"
COS_MIN = 0.9999  # north_star: cosine >= 0.9999 per chunk vs fp32 CPU


@pytest.fixture(scope="module")
def native():
    from claude_semantic_search_b200 import _native
    assert _native.device_count() >= 1
    return _native


def _bf16(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


# ------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K,gelu", [(128, 256, 64, 0), (128, 256, 768, 0), (1, 256, 128, 0), (130, 768, 768, 0),
                                        (1000, 2304, 768, 0), (333, 3072, 768, 1), (777, 768, 3072, 0),
                                        (40000, 768, 768, 0)])
@pytest.mark.parametrize("kernel", [2, 4])   # 2 = single-CTA tiles, 4 = 2-CTA (cta_group::2) tiles
def test_tcgen05_gemm_vs_fp32(native, M, N, K, gelu, kernel):
    import torch
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K), dtype=np.float32)
    B = (rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    out = np.empty((M, N), np.float32)
    native.check(native.load().css_debug_gemm(A.ctypes.data, B.ctypes.data, bias.ctypes.data, M, N, K, gelu | kernel, 0,
                                              out.ctypes.data))
    ref = torch.from_numpy(_bf16(A)).double() @ torch.from_numpy(_bf16(B)).double().T + torch.from_numpy(bias).double()
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    ref = ref.numpy()
    # fp32 accumulation error + one bf16 rounding of the result
    err = np.abs(out - ref)
    tol = 2.0 ** -8 * np.abs(ref) + 2e-3
    assert (err <= tol).all(), f"max err {err.max():.4g} at {np.unravel_index(err.argmax(), err.shape)}"


@pytest.mark.parametrize("M,K", [(1, 768), (127, 768), (256, 768), (1000, 3072), (33000, 768)])
@pytest.mark.parametrize("mode", [0, 1, 2, 3])   # bit 0: 2-CTA tiles, bit 1: LayerNorm inside the epilogue
def test_gemm_resid_layernorm_epilogue_vs_fp64(native, M, K, mode):
    """The fused epilogue of the attention-output / FFN-down projections: LayerNorm(A W^T + b + resid)."""
    import torch
    rng = np.random.default_rng(M + K)
    N = 768
    A = rng.standard_normal((M, K), dtype=np.float32)
    B = (rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    resid = (rng.standard_normal((M, N)) * 1.5 + 0.3).astype(np.float32)
    gamma = (1 + 0.2 * rng.standard_normal(N)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(N)).astype(np.float32)
    out = np.empty((M, N), np.float32)
    native.check(native.load().css_debug_gemm_resid_ln(A.ctypes.data, B.ctypes.data, bias.ctypes.data, resid.ctypes.data,
                                                       gamma.ctypes.data, beta.ctypes.data, M, K, 1e-5, mode, 0,
                                                       out.ctypes.data))
    pre = (torch.from_numpy(_bf16(A)).double() @ torch.from_numpy(_bf16(B)).double().T + torch.from_numpy(bias).double()
           + torch.from_numpy(_bf16(resid)).double())
    ref = torch.nn.functional.layer_norm(pre, (N,), torch.from_numpy(gamma).double(), torch.from_numpy(beta).double(),
                                         1e-5).numpy()
    err = np.abs(out - ref)
    # pre-LayerNorm value rounded to bf16 once (relative 2^-9 of |pre| ~ a few), the result once more
    tol = 2.0 ** -7 * np.abs(ref) + 2.5e-2
    assert (err <= tol).all(), f"max err {err.max():.4g} at {np.unravel_index(err.argmax(), err.shape)}"
    assert err.mean() < 4e-3


# ------------------------------------------------------------- attention
@pytest.mark.parametrize("tc", ["1", "0"])   # tcgen05 kernel / mma.sync kernel
@pytest.mark.parametrize("lens", [[1], [2, 3], [64], [65, 63], [128, 5, 200], [384], [129, 448, 300], [512, 17],
                                  [384, 383, 321, 320, 257, 193, 192, 129, 100, 7] * 20, [384] * 40])
def test_attention_vs_fp32(native, lens, tc, monkeypatch):
    import torch
    monkeypatch.setenv("CSS_ATTN_TC", tc)
    rng = np.random.default_rng(sum(lens))
    T = sum(lens)
    cu = np.zeros(len(lens) + 1, np.int32)
    cu[1:] = np.cumsum(lens)
    qkv = (rng.standard_normal((T, 2304)) * 1.5).astype(np.float32)
    half = 511
    rel = rng.standard_normal((12, 2 * half + 1)).astype(np.float32)
    ctx = np.empty((T, 768), np.float32)
    native.check(native.load().css_debug_attention(qkv.ctypes.data, cu.ctypes.data, len(lens), rel.ctypes.data, half, 0,
                                                   ctx.ctypes.data))
    qb = torch.from_numpy(_bf16(qkv)).double()
    for s, L in enumerate(lens):
        blk = qb[cu[s]:cu[s + 1]]
        q = blk[:, :768].view(L, 12, 64).transpose(0, 1)
        k = blk[:, 768:1536].view(L, 12, 64).transpose(0, 1)
        v = blk[:, 1536:].view(L, 12, 64).transpose(0, 1)
        i = torch.arange(L)
        bias = torch.from_numpy(rel).double()[:, (i[None, :] - i[:, None]) + half]   # [12, L(query), L(key)]
        p = torch.softmax(q @ k.transpose(1, 2) / 8.0 + bias, dim=-1)
        ref = (p @ v).transpose(0, 1).reshape(L, 768).numpy()
        err = np.abs(ctx[cu[s]:cu[s + 1]] - ref)
        # P is rounded to bf16 before the PV product, the output once more
        assert err.max() < 4e-2, f"seq {s} (L={L}): max err {err.max():.4g}"
        assert err.mean() < 3e-3


# ------------------------------------------------------------- end to end
def _golden():
    g = np.load(GOLDEN / "encoder_small.npz")
    cu = g["cu_seqlens"]
    return g, [g["ids"][cu[i]:cu[i + 1]].tolist() for i in range(len(cu) - 1)]


@pytest.mark.parametrize("perturb", [False, True])
def test_encoder_vs_golden(native, perturb):
    from claude_semantic_search_b200.encoder import MPNetEncoder
    from oracle import encoder_oracle as eo
    g, seqs = _golden()
    model = eo.build_model(seed=0, perturb=perturb)
    enc = MPNetEncoder.from_hf_model(model, max_tokens=2048)
    got = enc.encode_ids(seqs)
    want = g["emb_perturbed" if perturb else "emb_plain"]
    cos = eo.cosine_rows(want, got)
    # The north_star bar (0.9999) is asserted on the configuration it names: random-init
    # weights (HF initialisation).  The perturbed model (2.5x linear weights, random LayerNorm
    # gains and biases) is a stress test for layout / bias / gamma bugs: bf16 rounding
    # accumulates ~1e-2 relative error through its 12 layers, so its bar is 0.9998.
    bar = 0.9998 if perturb else COS_MIN
    print(f"encoder golden perturb={perturb}: min cosine {cos.min():.6f}")
    assert cos.min() >= bar, f"min cosine {cos.min():.6f} (per row {np.round(cos, 6)})"
    np.testing.assert_allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-4)
    if perturb:
        raw = enc.encode_ids(seqs, normalize=False)
        want_raw = g["emb_perturbed_unnormalized"]
        assert eo.cosine_rows(want_raw, raw).min() >= bar
        ratio = np.linalg.norm(raw, axis=1) / np.linalg.norm(want_raw, axis=1)
        assert np.abs(ratio - 1).max() < 2e-2
    # small workspace: the same call split into several passes gives the same rows
    enc2 = MPNetEncoder.from_hf_model(model, max_tokens=512)
    got2 = enc2.encode_ids(seqs)
    assert eo.cosine_rows(got, got2).min() > 0.999999
    enc.close()
    enc2.close()


def test_encoder_ragged_batch_vs_oracle(native):
    """SURVEY 8d ragged correctness set (scaled): lengths U[8, 384], perturbed weights."""
    from claude_semantic_search_b200.encoder import MPNetEncoder
    from oracle import encoder_oracle as eo
    rng = np.random.default_rng(5)
    lengths = rng.integers(8, 385, size=48).tolist() + [1, 2, 384, 384]
    seqs = eo.synthetic_ids(len(lengths), lengths, seed=9)
    model = eo.build_model(seed=0, perturb=True, num_layers=4)
    enc = MPNetEncoder.from_hf_model(model)
    got = enc.encode_ids(seqs)
    want = eo.st_encode_ids(model, seqs, batch_size=16)
    cos = eo.cosine_rows(want, got)
    print(f"encoder ragged (perturbed, 4 layers): min cosine {cos.min():.6f}")
    assert cos.min() >= COS_MIN, f"min cosine {cos.min():.6f}"
    # determinism + independence from batch composition
    again = enc.encode_ids(seqs[::-1])[::-1]
    np.testing.assert_array_equal(got, again)
    enc.close()


def test_encoder_at_the_benchmarked_shape_vs_oracle(native):
    """VERDICT r1 item 7 / SURVEY 8d config 3: the configuration bench.py times -- 12 layers, the prescribed random-init
    weights, full 384-token chunks in passes of 296 -- plus the ragged correctness set (lengths U[8, 384]) against the
    fp32 CPU oracle: cosine >= 0.9999 per chunk.  (512 + 1024 sequences: about a minute of CPU.)  Random-init MPNet
    outputs are nearly collinear (pairwise cosine ~0.99), so the embeddings are also compared after removing their
    common mean direction -- the part of the vector that ranks neighbours."""
    from claude_semantic_search_b200.encoder import MPNetEncoder
    from oracle import encoder_oracle as eo
    model = eo.build_model(seed=0)
    enc = MPNetEncoder.from_hf_model(model, max_tokens=296 * 384)
    full = eo.synthetic_ids(512, [384], seed=7)
    rng = np.random.default_rng(17)
    ragged = eo.synthetic_ids(1024, rng.integers(8, 385, size=1024).tolist(), seed=19)
    for name, seqs in (("512 x 384", full), ("1024 ragged", ragged)):
        got = enc.encode_ids(seqs)
        want = eo.st_encode_ids(model, seqs, batch_size=32)
        cos = eo.cosine_rows(want, got)
        mu = want.mean(axis=0, keepdims=True)
        cos_c = eo.cosine_rows(want - mu, got - mu)
        print(f"encoder {name}: min cosine {cos.min():.6f}, centred {cos_c.min():.5f} (median {np.median(cos_c):.5f})")
        assert cos.min() >= COS_MIN, f"{name}: min cosine {cos.min():.6f}"
        assert np.median(cos_c) >= 0.99, f"{name}: centred median cosine {np.median(cos_c):.5f}"
        np.testing.assert_allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-4)
    enc.close()


def test_single_query_graph_path_matches_batch_path(native):
    """a single short sequence is served by a captured CUDA graph over a token
    bucket (32 / 64 / 128 / 256 / 384 rows).  Buckets of 128+ rows run the batch kernels and must return
    exactly what the batch path returns; the 32 / 64-row buckets run the weight-streaming query kernels
    (query_kernels.cuh: different summation order) and are held to the oracle bar instead, and to
    CSS_QUERY_SKINNY=0-style agreement with the batch path within bf16 noise."""
    from claude_semantic_search_b200.encoder import MPNetEncoder
    from oracle import encoder_oracle as eo
    lengths = [1, 5, 31, 32, 33, 64, 100, 128, 129, 256, 300, 384]
    seqs = eo.synthetic_ids(len(lengths), lengths, seed=21)
    model = eo.build_model(seed=0, perturb=True, num_layers=3)
    enc = MPNetEncoder.from_hf_model(model)
    batch = enc.encode_ids(seqs)                      # n_seq > 1: plain launches
    want = eo.st_encode_ids(model, seqs, batch_size=16)
    assert eo.cosine_rows(want, batch).min() >= COS_MIN
    first = {}
    for rep in range(2):                              # first call captures, second replays
        for i, q in enumerate(seqs):
            one = enc.encode_ids([q])
            if lengths[i] > 64:
                np.testing.assert_array_equal(one[0], batch[i], err_msg=f"L={lengths[i]} rep={rep}")
            else:
                assert eo.cosine_rows(want[i:i + 1], one).min() >= COS_MIN, f"L={lengths[i]} rep={rep}"
                np.testing.assert_allclose(one[0], batch[i], atol=2e-3, err_msg=f"L={lengths[i]} rep={rep}")
                if rep == 1:
                    np.testing.assert_array_equal(one[0], first[i])   # replay == capture run
            if rep == 0:
                first[i] = one[0].copy()
    raw = enc.encode_ids([seqs[4]], normalize=False)  # separate graph per normalize flag
    np.testing.assert_allclose(raw[0] / np.linalg.norm(raw[0]), first[4], atol=2e-6)
    raw = enc.encode_ids([seqs[7]], normalize=False)
    np.testing.assert_allclose(raw[0] / np.linalg.norm(raw[0]), batch[7], atol=2e-6)
    enc.close()


def test_many_short_sequences_one_pass(native):
    """Thousands of 1..30-token sequences in one pass: cu_seqlens beyond the attention kernel's shared-memory
    cache, one (sequence, head) unit per few CTA iterations, token rows far below a GEMM tile.  The result
    must not depend on how the sequences are grouped into passes, and must match the oracle."""
    from claude_semantic_search_b200.encoder import MPNetEncoder
    from oracle import encoder_oracle as eo
    rng = np.random.default_rng(11)
    lengths = rng.integers(1, 31, size=3000).tolist()
    seqs = eo.synthetic_ids(len(lengths), lengths, seed=13)
    model = eo.build_model(seed=0, perturb=True, num_layers=2)
    enc = MPNetEncoder.from_hf_model(model)
    one = enc.encode_ids(seqs)
    parts = np.concatenate([enc.encode_ids(seqs[i:i + 400]) for i in range(0, len(seqs), 400)])
    np.testing.assert_array_equal(one, parts)
    want = eo.st_encode_ids(model, seqs[:64], batch_size=16)
    assert eo.cosine_rows(want, one[:64]).min() >= COS_MIN
    enc.close()


def test_encoder_errors(native):
    from claude_semantic_search_b200.encoder import MPNetEncoder
    from oracle import encoder_oracle as eo
    enc = MPNetEncoder.from_hf_model(eo.build_model(seed=0, num_layers=1), max_tokens=1024)
    assert enc.encode_ids([]).shape == (0, 768)
    with pytest.raises(native.NativeError):
        enc.encode_ids([[0, 2], []])                 # empty sequence
    with pytest.raises(native.NativeError):
        enc.encode_ids([[0] * 600])                  # longer than max_seq_len
    enc.close()
