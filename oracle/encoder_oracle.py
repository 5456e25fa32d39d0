"""oracle/encoder_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of half A of the hot path: what
`SentenceTransformer("all-mpnet-base-v2").encode(texts, batch_size=..., normalize_embeddings=...)`
computes after tokenisation (reference call sites: src/embeddings.py:184-188, :216-222).

The arithmetic lives in third-party code that is not vendored in the reference:
  * `transformers` MPNetModel (installed here, 5.5.0: models/mpnet/modeling_mpnet.py)
    -- used directly, fp32, eval mode: this IS the reference's encoder.
  * `sentence-transformers` (>= 5.0, not installable here): only its thin wrapper is
    restated below from its published behaviour --
      - SentenceTransformer.encode: sort texts by length (descending), run batches of
        `batch_size`, pad each batch to its longest sequence with the pad id and an
        attention mask, restore the original order;
      - models.Pooling(pooling_mode_mean_tokens): sum(tok * mask) / clamp(sum(mask), min=1e-9);
      - models.Normalize: F.normalize(x, p=2, dim=1) (eps 1e-12).

Parity pin: the reference's own tests mock SentenceTransformer everywhere (SURVEY.md
section 8c: "Encoder: none"), and the pretrained weights / vocabulary are not available
offline, so the model-level pin is the real HF implementation itself plus the fixtures
under tests/golden/encoder_*.npz generated from it by oracle/make_golden_encoder.py.
Weights: random init under torch.manual_seed(seed) as north_star prescribes; `perturb`
additionally randomises every bias and LayerNorm parameter so those code paths are
exercised (HF initialises them to 0 / 1).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

PAD_ID = 1
BOS_ID = 0
EOS_ID = 2

CONFIG = dict(vocab_size=30527, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
              intermediate_size=3072, max_position_embeddings=514, layer_norm_eps=1e-5,
              relative_attention_num_buckets=32, pad_token_id=1)


def build_model(seed: int = 0, perturb: bool = False, num_layers: int = 12):
    """MPNetModel(add_pooling_layer=False).eval() with random-init weights."""
    import torch
    from transformers import MPNetConfig, MPNetModel
    cfg = MPNetConfig(**{**CONFIG, "num_hidden_layers": num_layers})
    torch.manual_seed(seed)
    model = MPNetModel(cfg, add_pooling_layer=False).eval()
    if perturb:
        g = torch.Generator().manual_seed(seed + 1000)
        with torch.no_grad():
            for name, p in model.named_parameters():
                if name.endswith("LayerNorm.weight"):
                    p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
                elif name.endswith(".bias"):
                    p.copy_(0.1 * torch.randn(p.shape, generator=g))
                elif "relative_attention_bias" in name:
                    p.copy_(0.5 * torch.randn(p.shape, generator=g))
                elif name.endswith("dense.weight") or ".attn." in name:
                    # larger linear weights: makes attention peaky and activations non-trivial
                    p.mul_(2.5)
    for p in model.parameters():
        p.requires_grad_(False)
    return model


def state_dict_numpy(model) -> Dict[str, np.ndarray]:
    return {k: v.detach().cpu().numpy().astype(np.float32) for k, v in model.state_dict().items()}


def st_encode_ids(model, seqs: Sequence[Sequence[int]], batch_size: int = 16, normalize: bool = True,
                  max_seq_length: int = 384) -> np.ndarray:
    """SentenceTransformer.encode restated for pre-tokenised input (see module docstring)."""
    import torch
    seqs = [list(s)[:max_seq_length] for s in seqs]
    n = len(seqs)
    out = np.zeros((n, model.config.hidden_size), np.float32)
    order = np.argsort([-len(s) for s in seqs], kind="stable")
    with torch.no_grad():
        for b0 in range(0, n, batch_size):
            idxs = order[b0:b0 + batch_size]
            L = max(len(seqs[i]) for i in idxs)
            ids = torch.full((len(idxs), L), PAD_ID, dtype=torch.long)
            mask = torch.zeros((len(idxs), L), dtype=torch.long)
            for r, i in enumerate(idxs):
                ids[r, :len(seqs[i])] = torch.tensor(seqs[i], dtype=torch.long)
                mask[r, :len(seqs[i])] = 1
            tok = model(input_ids=ids, attention_mask=mask).last_hidden_state
            m = mask.unsqueeze(-1).to(tok.dtype)
            pooled = (tok * m).sum(1) / torch.clamp(m.sum(1), min=1e-9)
            if normalize:
                pooled = torch.nn.functional.normalize(pooled, p=2, dim=1)
            out[idxs] = pooled.numpy()
    return out


def synthetic_ids(n_seq: int, lengths: Sequence[int], seed: int = 7, vocab: int = 30527) -> List[List[int]]:
    """ids uniform in [4, vocab-1), <s> first, </s> last."""
    rng = np.random.default_rng(seed)
    seqs = []
    for i in range(n_seq):
        L = int(lengths[i % len(lengths)])
        s = rng.integers(4, vocab - 1, size=L).tolist()
        s[0] = BOS_ID
        if L > 1:
            s[-1] = EOS_ID
        seqs.append(s)
    return seqs


def cosine_rows(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return (a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1) + 1e-30)
