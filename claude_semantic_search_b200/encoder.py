"""MPNetEncoder: the all-mpnet-base-v2 forward pass + mean pooling + L2 normalisation on
the B200 path (css_encoder_* in include/css_b200.h).

This is what `SentenceTransformer.encode` computes after tokenisation (reference call
sites src/embeddings.py:184-188, 216-222).  Weights come from a Hugging Face MPNet
state dict (names of transformers' MPNetModel); input is token ids.  There is no CPU
fallback: construction raises if libcss_b200.so or an sm_100 device is missing.
"""
from __future__ import annotations

import ctypes
import json
import struct
from ctypes import c_int64, c_void_p
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _native

DEFAULT_CONFIG = dict(vocab_size=30527, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                      intermediate_size=3072, max_position_embeddings=514, layer_norm_eps=1e-5,
                      relative_attention_num_buckets=32, pad_token_id=1)

_LAYER_NAMES = {
    "q_w": "attention.attn.q.weight", "q_b": "attention.attn.q.bias",
    "k_w": "attention.attn.k.weight", "k_b": "attention.attn.k.bias",
    "v_w": "attention.attn.v.weight", "v_b": "attention.attn.v.bias",
    "o_w": "attention.attn.o.weight", "o_b": "attention.attn.o.bias",
    "ln1_w": "attention.LayerNorm.weight", "ln1_b": "attention.LayerNorm.bias",
    "ffn1_w": "intermediate.dense.weight", "ffn1_b": "intermediate.dense.bias",
    "ffn2_w": "output.dense.weight", "ffn2_b": "output.dense.bias",
    "ln2_w": "output.LayerNorm.weight", "ln2_b": "output.LayerNorm.bias",
}


def _np32(t) -> np.ndarray:
    if hasattr(t, "detach"):
        t = t.detach().cpu().float().numpy()
    return np.ascontiguousarray(t, dtype=np.float32)


def read_safetensors(path: Union[str, Path]) -> Dict[str, np.ndarray]:
    """Minimal safetensors reader (F32 / F16 / BF16 tensors), no third-party import."""
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as fh:
        (hlen,) = struct.unpack("<Q", fh.read(8))
        header = json.loads(fh.read(hlen))
        base = 8 + hlen
        data = np.memmap(path, dtype=np.uint8, mode="r", offset=base)
    for name, meta in header.items():
        if name == "__metadata__":
            continue
        a, b = meta["data_offsets"]
        raw = np.asarray(data[a:b])
        dt = meta["dtype"]
        if dt == "F32":
            arr = raw.view(np.float32)
        elif dt == "F16":
            arr = raw.view(np.float16).astype(np.float32)
        elif dt == "BF16":
            arr = (raw.view(np.uint16).astype(np.uint32) << 16).view(np.float32)
        else:
            continue
        out[name] = arr.reshape(meta["shape"])
    return out


class MPNetEncoder:
    def __init__(self, state_dict: Dict[str, "np.ndarray"], config: Optional[dict] = None, device: int = 0,
                 max_tokens: int = 0):
        cfg = dict(DEFAULT_CONFIG)
        cfg.update(config or {})
        self.config = cfg
        self.device = device
        self._lib = _native.load()
        sd = self._strip_prefix(state_dict)
        keep: List[np.ndarray] = []

        def get(name: str, shape: Tuple[int, ...]) -> int:
            if name not in sd:
                raise KeyError(f"MPNet state dict has no '{name}'")
            a = _np32(sd[name])
            if tuple(a.shape) != tuple(shape):
                raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
            keep.append(a)
            return a.ctypes.data

        H, F, Ln = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"]
        heads = cfg["num_attention_heads"]
        layers = (_native.css_mpnet_layer * Ln)()
        shapes = {"q_w": (H, H), "k_w": (H, H), "v_w": (H, H), "o_w": (H, H), "ffn1_w": (F, H), "ffn2_w": (H, F),
                  "ffn1_b": (F,)}
        for l in range(Ln):
            for field, suffix in _LAYER_NAMES.items():
                setattr(layers[l], field, get(f"encoder.layer.{l}.{suffix}", shapes.get(field, (H,))))
        w = _native.css_mpnet_weights()
        w.word_emb = get("embeddings.word_embeddings.weight", (cfg["vocab_size"], H))
        w.pos_emb = get("embeddings.position_embeddings.weight", (cfg["max_position_embeddings"], H))
        w.emb_ln_w = get("embeddings.LayerNorm.weight", (H,))
        w.emb_ln_b = get("embeddings.LayerNorm.bias", (H,))
        w.rel_bias = get("encoder.relative_attention_bias.weight", (cfg["relative_attention_num_buckets"], heads))
        w.layers = ctypes.cast(layers, ctypes.POINTER(_native.css_mpnet_layer))
        c = _native.css_mpnet_config(
            vocab_size=cfg["vocab_size"], hidden_size=H, num_layers=Ln, num_heads=heads, intermediate_size=F,
            max_position=cfg["max_position_embeddings"], rel_buckets=cfg["relative_attention_num_buckets"],
            rel_max_distance=128, pad_token_id=cfg["pad_token_id"], layer_norm_eps=cfg["layer_norm_eps"])
        self._h = c_void_p()
        _native.check(self._lib.css_encoder_create(ctypes.byref(c), ctypes.byref(w), device, c_int64(max_tokens),
                                                   ctypes.byref(self._h)))
        del keep  # the library copied everything to the device
        self.dim = int(self._lib.css_encoder_dim(self._h))
        self.max_tokens = int(self._lib.css_encoder_max_tokens(self._h))
        self.max_seq_len = int(self._lib.css_encoder_max_seq_len(self._h))

    # ------------------------------------------------------------------ loading
    @staticmethod
    def _strip_prefix(sd: Dict[str, "np.ndarray"]) -> Dict[str, "np.ndarray"]:
        for prefix in ("", "mpnet.", "0.auto_model.", "auto_model."):
            if prefix + "embeddings.word_embeddings.weight" in sd:
                return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
        raise KeyError("not an MPNet state dict (no embeddings.word_embeddings.weight)")

    @classmethod
    def from_hf_model(cls, model, device: int = 0, max_tokens: int = 0) -> "MPNetEncoder":
        """From an in-memory transformers MPNetModel (or anything with state_dict()/config)."""
        cfg = {k: getattr(model.config, k) for k in DEFAULT_CONFIG if hasattr(model.config, k)}
        return cls(model.state_dict(), cfg, device, max_tokens)

    @classmethod
    def from_pretrained(cls, path: Union[str, Path], device: int = 0, max_tokens: int = 0) -> "MPNetEncoder":
        """From a local model directory (config.json + model.safetensors | pytorch_model.bin),
        the layout sentence-transformers caches all-mpnet-base-v2 in."""
        path = Path(path)
        cfg = {}
        if (path / "config.json").exists():
            raw = json.loads((path / "config.json").read_text())
            cfg = {k: raw[k] for k in DEFAULT_CONFIG if k in raw}
        if (path / "model.safetensors").exists():
            sd = read_safetensors(path / "model.safetensors")
        elif (path / "pytorch_model.bin").exists():
            import torch
            sd = torch.load(path / "pytorch_model.bin", map_location="cpu", weights_only=True)
        else:
            raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin under {path}")
        return cls(sd, cfg, device, max_tokens)

    # ------------------------------------------------------------------- encode
    @staticmethod
    def pack(seqs: Iterable[Sequence[int]]) -> Tuple[np.ndarray, np.ndarray]:
        seqs = [np.asarray(s, dtype=np.int32).reshape(-1) for s in seqs]
        cu = np.zeros(len(seqs) + 1, dtype=np.int32)
        if seqs:
            np.cumsum([len(s) for s in seqs], out=cu[1:])
        ids = np.concatenate(seqs) if seqs else np.zeros(0, np.int32)
        return np.ascontiguousarray(ids, dtype=np.int32), cu

    def encode_ids(self, seqs: Iterable[Sequence[int]], normalize: bool = True) -> np.ndarray:
        """[n, 768] float32 embeddings of pre-tokenised sequences (no padding, any lengths
        <= max_seq_len).  Order of the result == order of the input."""
        ids, cu = self.pack(seqs)
        return self.encode_packed(ids, cu, normalize)

    def encode_packed(self, ids: np.ndarray, cu_seqlens: np.ndarray, normalize: bool = True) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        cu = np.ascontiguousarray(cu_seqlens, dtype=np.int32)
        n = cu.shape[0] - 1
        out = np.empty((max(n, 0), self.dim), np.float32)
        if n <= 0:
            return out
        if int(cu[-1]) != ids.shape[0]:
            raise ValueError("cu_seqlens[-1] != len(ids)")
        _native.check(self._lib.css_encoder_encode(self._h, ids.ctypes.data, cu.ctypes.data, n,
                                                   1 if normalize else 0, out.ctypes.data))
        return out

    def encode_device(self, ids_ptr: int, cu_ptr: int, cu_host: np.ndarray, out_ptr: int, normalize: bool = True,
                      stream: int = 0) -> None:
        cu = np.ascontiguousarray(cu_host, dtype=np.int32)
        _native.check(self._lib.css_encoder_encode_device(self._h, c_void_p(ids_ptr), c_void_p(cu_ptr),
                                                          cu.ctypes.data, cu.shape[0] - 1, 1 if normalize else 0,
                                                          c_void_p(out_ptr), c_void_p(stream)))

    # ---------------------------------------------------------------- lifecycle
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.css_encoder_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def split_by_tokens(cu_seqlens: np.ndarray, parts: int) -> List[Tuple[int, int]]:
    """Contiguous sequence ranges [a, b) with (nearly) equal token counts, one per device; the order of the
    sequences is kept, a range may be empty.  SURVEY.md 8(e): "static contiguous split ... weights replicated;
    each rank writes rows [start_r, end_r)"."""
    cu = np.asarray(cu_seqlens, dtype=np.int64)
    n = cu.shape[0] - 1
    total = int(cu[-1]) if n > 0 else 0
    cuts = [0]
    for r in range(1, parts):
        # first sequence boundary at or after r / parts of the tokens, never before the previous cut
        target = total * r / parts
        cuts.append(max(cuts[-1], int(np.searchsorted(cu, target, side="left"))))
    cuts.append(n)
    return [(min(cuts[r], n), min(cuts[r + 1], n)) for r in range(parts)]


class MultiDeviceEncoder:
    """One encoder over several GPUs of this box, inside one process: the weights are replicated (one
    css_encoder per device), a batch of sequences is split into contiguous ranges of equal token count and every
    range is encoded on its own device from its own host thread (ctypes releases the GIL; no collective -- SURVEY.md
    8(e)).  Results are bit-identical to a single device: a sequence's embedding does not depend on what else is
    in the pass.  Small calls (a query, a handful of chunks) stay on the first device."""

    def __init__(self, encoders: List[MPNetEncoder], min_seqs_per_device: int = 8):
        if not encoders:
            raise ValueError("MultiDeviceEncoder needs at least one encoder")
        self._encoders = list(encoders)
        self.devices = [e.device for e in encoders]
        self.device = self.devices[0]
        self.dim = encoders[0].dim
        self.max_tokens = encoders[0].max_tokens
        self.max_seq_len = encoders[0].max_seq_len
        self.config = encoders[0].config
        self.min_seqs_per_device = min_seqs_per_device
        self._pool = None

    @classmethod
    def create(cls, make_one, devices: Sequence[int]) -> "MultiDeviceEncoder":
        """`make_one(device) -> MPNetEncoder`; the per-device loads run concurrently."""
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=len(devices)) as pool:
            return cls(list(pool.map(make_one, devices)))

    pack = staticmethod(MPNetEncoder.pack)

    def encode_ids(self, seqs: Iterable[Sequence[int]], normalize: bool = True) -> np.ndarray:
        ids, cu = self.pack(seqs)
        return self.encode_packed(ids, cu, normalize)

    def encode_packed(self, ids: np.ndarray, cu_seqlens: np.ndarray, normalize: bool = True) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        cu = np.ascontiguousarray(cu_seqlens, dtype=np.int32)
        n = cu.shape[0] - 1
        parts = min(len(self._encoders), max(1, n // self.min_seqs_per_device))
        if parts <= 1:
            return self._encoders[0].encode_packed(ids, cu, normalize)
        out = np.empty((n, self.dim), np.float32)
        ranges = [(e, a, b) for e, (a, b) in zip(self._encoders, split_by_tokens(cu, parts)) if b > a]

        def run(job):
            enc, a, b = job
            out[a:b] = enc.encode_packed(ids[cu[a]:cu[b]], cu[a:b + 1] - cu[a], normalize)

        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=len(self._encoders))
        list(self._pool.map(run, ranges))   # re-raises the first failure
        return out

    def close(self) -> None:
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        for e in self._encoders:
            e.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def relative_bucket(rel: int, num_buckets: int = 32, max_distance: int = 128) -> int:
    return int(_native.load().css_mpnet_relative_bucket(int(rel), num_buckets, max_distance))


def random_state_dict(seed: int = 0, config: Optional[dict] = None) -> Dict[str, np.ndarray]:
    """Random-init MPNet weights of the all-mpnet-base-v2 architecture (normal(0, 0.02)
    linear / embedding weights, zero biases, unit LayerNorm), for synthetic benchmarks
    where the pretrained checkpoint is unavailable."""
    cfg = dict(DEFAULT_CONFIG)
    cfg.update(config or {})
    rng = np.random.default_rng(seed)
    H, F = cfg["hidden_size"], cfg["intermediate_size"]

    def w(*shape):
        return (rng.standard_normal(shape, dtype=np.float32) * 0.02).astype(np.float32)

    sd = {"embeddings.word_embeddings.weight": w(cfg["vocab_size"], H),
          "embeddings.position_embeddings.weight": w(cfg["max_position_embeddings"], H),
          "embeddings.LayerNorm.weight": np.ones(H, np.float32), "embeddings.LayerNorm.bias": np.zeros(H, np.float32),
          "encoder.relative_attention_bias.weight": w(cfg["relative_attention_num_buckets"],
                                                      cfg["num_attention_heads"])}
    sd["embeddings.word_embeddings.weight"][cfg["pad_token_id"]] = 0
    sd["embeddings.position_embeddings.weight"][cfg["pad_token_id"]] = 0
    shapes = {"q_w": (H, H), "k_w": (H, H), "v_w": (H, H), "o_w": (H, H), "ffn1_w": (F, H), "ffn2_w": (H, F)}
    for l in range(cfg["num_hidden_layers"]):
        for field, suffix in _LAYER_NAMES.items():
            name = f"encoder.layer.{l}.{suffix}"
            if field in shapes:
                sd[name] = w(*shapes[field])
            elif field.startswith("ln") and field.endswith("_w"):
                sd[name] = np.ones(H, np.float32)
            else:
                sd[name] = np.zeros(F if field == "ffn1_b" else H, np.float32)
    return sd
