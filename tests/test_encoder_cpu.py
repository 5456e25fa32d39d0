"""CPU-only checks of half A: the encoder oracle against its committed fixture, the
relative-position bucket function of the C library against transformers, and the
host-side helpers (safetensors reader, packing)."""
import json
import struct
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).parent / "golden"


def test_relative_bucket_matches_transformers():
    import torch
    from transformers.models.mpnet.modeling_mpnet import MPNetEncoder as HFEncoder

    from claude_semantic_search_b200.encoder import relative_bucket
    rel = torch.arange(-700, 701)
    want = HFEncoder.relative_position_bucket(rel, num_buckets=32, max_distance=128).numpy()
    got = np.array([relative_bucket(int(r)) for r in rel])
    np.testing.assert_array_equal(got, want)


def test_oracle_reproduces_golden_fixture():
    """The fixture was written by oracle/make_golden_encoder.py from the real transformers
    MPNetModel; the weights are regenerated from the seed."""
    from oracle import encoder_oracle as eo
    g = np.load(GOLDEN / "encoder_small.npz")
    cu = g["cu_seqlens"]
    seqs = [g["ids"][cu[i]:cu[i + 1]].tolist() for i in range(len(cu) - 1)]
    assert [len(s) for s in seqs] == [2, 3, 9, 31, 64, 65, 127, 200, 384]
    # a subset keeps the CPU suite short; the GPU suite checks every row
    pick = [0, 2, 4, 5]
    out = eo.st_encode_ids(eo.build_model(seed=0, perturb=False), [seqs[i] for i in pick])
    np.testing.assert_allclose(out, g["emb_plain"][pick], atol=2e-6)
    np.testing.assert_allclose(np.linalg.norm(out, axis=1), 1.0, atol=1e-5)


def test_st_batching_is_order_invariant():
    """Padding to the longest sequence of a batch + attention mask does not change a
    sequence's embedding (what lets the device path run unpadded)."""
    from oracle import encoder_oracle as eo
    model = eo.build_model(seed=0, perturb=True, num_layers=2)
    seqs = eo.synthetic_ids(6, [5, 40, 17, 3, 29, 64], seed=11)
    a = eo.st_encode_ids(model, seqs, batch_size=16)
    b = eo.st_encode_ids(model, seqs, batch_size=1)
    assert eo.cosine_rows(a, b).min() > 0.999999


def test_safetensors_reader(tmp_path):
    from claude_semantic_search_b200.encoder import read_safetensors
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    b = (np.arange(6, dtype=np.float32) / 3).astype(np.float16)
    header = {"a": {"dtype": "F32", "shape": [3, 4], "data_offsets": [0, 48]},
              "b": {"dtype": "F16", "shape": [6], "data_offsets": [48, 60]}, "__metadata__": {"format": "pt"}}
    hj = json.dumps(header).encode()
    p = tmp_path / "model.safetensors"
    p.write_bytes(struct.pack("<Q", len(hj)) + hj + a.tobytes() + b.tobytes())
    sd = read_safetensors(p)
    np.testing.assert_array_equal(sd["a"], a)
    np.testing.assert_allclose(sd["b"], b.astype(np.float32))


def test_pack():
    from claude_semantic_search_b200.encoder import MPNetEncoder
    ids, cu = MPNetEncoder.pack([[0, 5, 2], [0, 2], [0, 7, 8, 9, 2]])
    assert ids.dtype == np.int32 and cu.tolist() == [0, 3, 5, 10]
    assert ids.tolist() == [0, 5, 2, 0, 2, 0, 7, 8, 9, 2]


def test_encoder_fails_loudly_without_gpu(gpu_available):
    if gpu_available:
        pytest.skip("GPU present")
    from claude_semantic_search_b200 import _native
    from claude_semantic_search_b200.encoder import MPNetEncoder
    H = 768
    sd = {"embeddings.word_embeddings.weight": np.zeros((8, H), np.float32),
          "embeddings.position_embeddings.weight": np.zeros((16, H), np.float32),
          "embeddings.LayerNorm.weight": np.ones(H, np.float32), "embeddings.LayerNorm.bias": np.zeros(H, np.float32),
          "encoder.relative_attention_bias.weight": np.zeros((32, 12), np.float32)}
    with pytest.raises(_native.NoDeviceError):
        MPNetEncoder(sd, dict(vocab_size=8, max_position_embeddings=16, num_hidden_layers=0 + 1) | {}, 0) \
            if False else MPNetEncoder(_tiny_sd(sd), dict(vocab_size=8, max_position_embeddings=16,
                                                          num_hidden_layers=1), 0)


def _tiny_sd(sd):
    H, F = 768, 3072
    z = lambda *s: np.zeros(s, np.float32)
    p = "encoder.layer.0."
    sd = dict(sd)
    for n in ("q", "k", "v", "o"):
        sd[p + f"attention.attn.{n}.weight"] = z(H, H)
        sd[p + f"attention.attn.{n}.bias"] = z(H)
    sd[p + "attention.LayerNorm.weight"] = z(H)
    sd[p + "attention.LayerNorm.bias"] = z(H)
    sd[p + "intermediate.dense.weight"] = z(F, H)
    sd[p + "intermediate.dense.bias"] = z(F)
    sd[p + "output.dense.weight"] = z(H, F)
    sd[p + "output.dense.bias"] = z(H)
    sd[p + "output.LayerNorm.weight"] = z(H)
    sd[p + "output.LayerNorm.bias"] = z(H)
    return sd
