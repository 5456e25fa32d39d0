"""Inner seam: the slice of `sentence_transformers.SentenceTransformer` the reference calls
(src/embeddings.py:86-97,117,184-188,216-222; kwargs pinned by tests/test_embeddings.py:164-166,
193-199), served by css_encoder_*.  Install as `sys.modules["sentence_transformers"]` to run the
reference's own src/embeddings.py unmodified on the B200 path.

Model files are read from a local directory only (no network): config.json,
model.safetensors | pytorch_model.bin, and tokenizer.json | vocab.txt.  When the checkpoint is
absent, `CSS_B200_SYNTHETIC_MODEL=1` (or model name "synthetic-mpnet") selects random-init
weights of the same architecture and a deterministic stand-in tokenizer -- for benchmarks and
tests, never silently.
"""
from __future__ import annotations

import os
import re
import unicodedata
import zlib
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from .encoder import DEFAULT_CONFIG, MPNetEncoder, MultiDeviceEncoder, random_state_dict

BOS, PAD, EOS, UNK = 0, 1, 2, 3


# ---------------------------------------------------------------------------- tokenizers
class StandInTokenizer:
    """ids = 4 + crc32(lower-cased word) mod (vocab - 6), <s> ... </s>, truncated
    (the real vocabulary is not available offline)."""

    def __init__(self, vocab_size: int = 30527):
        self.mod = vocab_size - 6
        self._word = re.compile(r"\w+|[^\w\s]", re.UNICODE)

    def encode_batch(self, texts: Sequence[str], max_length: int) -> List[List[int]]:
        out = []
        for t in texts:
            words = self._word.findall(t.lower())[: max(max_length - 2, 0)]
            out.append([BOS] + [4 + zlib.crc32(w.encode("utf-8")) % self.mod for w in words] + [EOS])
        return out


class WordPieceTokenizer:
    """Plain-Python BERT-style basic + WordPiece tokenisation over vocab.txt (lower-case, strip accents, split
    punctuation and the main CJK block, greedy longest-match with '##').  Readable cross-check of the native
    tokenizer on ordinary text (tests) and holder of the vocabulary; the product path is
    NativeWordPieceTokenizer, which alone reproduces the reference's pipeline on all of Unicode."""

    def __init__(self, vocab_file: Union[str, Path], do_lower_case: bool = True):
        self.vocab: Dict[str, int] = {}
        with open(vocab_file, encoding="utf-8") as fh:
            for i, line in enumerate(fh):
                self.vocab[line.rstrip("\n")] = i
        self.lower = do_lower_case
        self.bos = self.vocab.get("<s>", BOS)
        self.eos = self.vocab.get("</s>", EOS)
        self.unk = self.vocab.get("[UNK]", self.vocab.get("<unk>", UNK))

    @staticmethod
    def _is_punct(ch: str) -> bool:
        cp = ord(ch)
        if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
            return True
        return unicodedata.category(ch).startswith("P")

    def _basic(self, text: str) -> List[str]:
        if self.lower:
            text = text.lower()
            text = "".join(c for c in unicodedata.normalize("NFD", text) if unicodedata.category(c) != "Mn")
        out, cur = [], []
        for ch in text:
            if ch.isspace():
                if cur:
                    out.append("".join(cur))
                    cur = []
            elif self._is_punct(ch) or 0x4E00 <= ord(ch) <= 0x9FFF:
                if cur:
                    out.append("".join(cur))
                    cur = []
                out.append(ch)
            else:
                cur.append(ch)
        if cur:
            out.append("".join(cur))
        return out

    def _wordpiece(self, word: str) -> List[int]:
        if len(word) > 100:
            return [self.unk]
        ids, start = [], 0
        while start < len(word):
            end, cur = len(word), None
            while start < end:
                sub = word[start:end]
                if start > 0:
                    sub = "##" + sub
                if sub in self.vocab:
                    cur = self.vocab[sub]
                    break
                end -= 1
            if cur is None:
                return [self.unk]
            ids.append(cur)
            start = end
        return ids

    def encode_batch(self, texts: Sequence[str], max_length: int) -> List[List[int]]:
        out = []
        for t in texts:
            ids: List[int] = []
            for w in self._basic(t):
                ids.extend(self._wordpiece(w))
                if len(ids) >= max_length - 2:
                    break
            out.append([self.bos] + ids[: max(max_length - 2, 0)] + [self.eos])
        return out


MPNET_SPECIALS = ("<s>", "<pad>", "</s>", "<unk>", "[UNK]", "<mask>")


def native_tokenizer_settings(tokenizer_json: Union[str, Path]) -> Optional[Dict]:
    """Read a `tokenizers` tokenizer.json and return {"lower": bool, "specials": {literal: id}} when it
    is the pipeline css_tokenizer_encode_batch implements (BertNormalizer with clean_text +
    handle_chinese_chars and strip_accents following lowercase, BertPreTokenizer, WordPiece with "##"
    and 100 characters per word, <s> ... </s> framing, added tokens matched on the raw text); None
    otherwise (the caller then keeps the `tokenizers` package)."""
    import json
    try:
        cfg = json.loads(Path(tokenizer_json).read_text(encoding="utf-8"))
    except (OSError, ValueError):
        return None
    nz, pt, model = cfg.get("normalizer") or {}, cfg.get("pre_tokenizer") or {}, cfg.get("model") or {}
    if nz.get("type") != "BertNormalizer" or not nz.get("clean_text", True) or not nz.get("handle_chinese_chars", True):
        return None
    lower = bool(nz.get("lowercase", True))
    if nz.get("strip_accents") not in (None, lower):
        return None
    if pt.get("type") != "BertPreTokenizer" or model.get("type", "WordPiece") != "WordPiece":
        return None
    if model.get("continuing_subword_prefix", "##") != "##" or model.get("max_input_chars_per_word", 100) != 100:
        return None
    vocab = model.get("vocab") or {}
    if model.get("unk_token", "[UNK]") not in ("[UNK]", "<unk>") or vocab.get("<s>") is None or vocab.get("</s>") is None:
        return None
    post = cfg.get("post_processor") or {}
    if post.get("type") == "RobertaProcessing":
        if post.get("cls", ["<s>"])[0] != "<s>" or post.get("sep", ["</s>"])[0] != "</s>":
            return None
    elif post.get("type") == "TemplateProcessing":
        single = [next(iter(x.values())).get("id") for x in post.get("single", [])]
        if single != ["<s>", "A", "</s>"]:
            return None
    else:
        return None
    specials = {}
    for at in cfg.get("added_tokens") or []:
        if at.get("single_word") or (at.get("normalized") and not at.get("special")):
            return None
        specials[at["content"]] = int(at["id"])
    return {"lower": lower, "specials": specials, "vocab": vocab}


class NativeWordPieceTokenizer:
    """The reference's fast tokenizer pipeline (BertNormalizer -> BertPreTokenizer -> WordPiece, <s> ... </s>,
    truncation) through libcss_b200's multi-threaded C++ implementation (css_tokenizer_encode_batch),
    producing the packed ids / cu_seqlens the encoder consumes.  Covers all of Unicode (tables generated
    from the `tokenizers` package, scripts/gen_unicode_tables.py) and the added special tokens, which are
    cut out of the raw text like the reference does (`specials`: literal -> id; default: the MPNet
    specials present in the vocabulary)."""

    def __init__(self, vocab_file: Union[str, Path], do_lower_case: bool = True, n_threads: int = 0,
                 specials: Optional[Dict[str, int]] = None):
        import ctypes

        from . import _native
        self._native = _native
        self._py = WordPieceTokenizer(vocab_file, do_lower_case)
        self._lib = _native.load()
        self._h = ctypes.c_void_p()
        _native.check(self._lib.css_tokenizer_create(str(vocab_file).encode(), 1 if do_lower_case else 0,
                                                     ctypes.byref(self._h)))
        if specials is None:
            specials = {s: self._py.vocab[s] for s in MPNET_SPECIALS if s in self._py.vocab}
        self.specials = dict(specials)
        for lit, tid in self.specials.items():
            _native.check(self._lib.css_tokenizer_add_special(self._h, lit.encode("utf-8"), int(tid)))
        self.n_threads = n_threads

    def encode_packed(self, texts: Sequence[str], max_length: int):
        import ctypes
        n = len(texts)
        cu = np.zeros(n + 1, np.int32)
        if n == 0:
            return np.zeros(0, np.int32), cu
        raw = [t.encode("utf-8") for t in texts]
        ptrs = (ctypes.c_char_p * n)(*raw)
        lens = np.fromiter((len(b) for b in raw), dtype=np.int64, count=n)
        ids = np.empty(n * max_length, np.int32)
        fb = np.zeros(n, np.uint8)
        self._native.check(self._lib.css_tokenizer_encode_batch(
            self._h, ctypes.cast(ptrs, ctypes.c_void_p), lens.ctypes.data, n, max_length, ids.ctypes.data,
            cu.ctypes.data, fb.ctypes.data, self.n_threads))
        if fb.any():
            # the library flags malformed UTF-8 only; str.encode cannot produce it (lone surrogates raise above,
            # as they do in the reference's tokenizer): never answer with an approximation
            bad = int(np.flatnonzero(fb)[0])
            raise UnicodeError(f"text {bad} is not valid UTF-8")
        return ids[:cu[-1]], cu

    def encode_batch(self, texts: Sequence[str], max_length: int) -> List[List[int]]:
        ids, cu = self.encode_packed(texts, max_length)
        return [ids[cu[i]:cu[i + 1]].tolist() for i in range(len(texts))]

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.css_tokenizer_destroy(self._h)
            self._h.value = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FastTokenizer:
    """tokenizer.json through the `tokenizers` package (multi-threaded)."""

    def __init__(self, path: Union[str, Path]):
        from tokenizers import Tokenizer
        self.tok = Tokenizer.from_file(str(path))
        self.tok.no_padding()

    def encode_batch(self, texts: Sequence[str], max_length: int) -> List[List[int]]:
        self.tok.enable_truncation(max_length=max_length)
        return [e.ids for e in self.tok.encode_batch(list(texts))]


def _find_model_dir(name: str, cache_folder: Optional[str]) -> Optional[Path]:
    cands = []
    p = Path(name).expanduser()
    if p.is_dir():
        cands.append(p)
    roots = [cache_folder, os.environ.get("SENTENCE_TRANSFORMERS_HOME"),
             os.path.join(os.environ.get("HF_HOME", os.path.expanduser("~/.cache/huggingface")), "hub")]
    short = name.split("/")[-1]
    for r in roots:
        if not r:
            continue
        r = Path(r).expanduser()
        cands += [r / name, r / short, r / f"sentence-transformers_{short}"]
        hub = r / f"models--sentence-transformers--{short}" / "snapshots"
        if hub.is_dir():
            cands += sorted(hub.iterdir())
    for c in cands:
        if c.is_dir() and ((c / "model.safetensors").exists() or (c / "pytorch_model.bin").exists()):
            return c
    return None


# ---------------------------------------------------------------------- SentenceTransformer
def _device_index_of(device) -> int:
    """'cuda:3' -> 3; 'cuda', 'auto', None, 'cpu' -> 0 (torch's current-device convention)."""
    s = str(device) if device is not None else ""
    if s.startswith("cuda:"):
        try:
            return int(s.split(":", 1)[1])
        except ValueError:
            return 0
    return 0


def encode_texts_pipelined(tokenizer, encoder, texts: Sequence[str], max_len: int, normalize: bool,
                           slab: int = 1024) -> np.ndarray:
    """Text in, embeddings out with the host tokenizer one slab ahead of the GPU: a worker thread
    tokenises texts[i+1] (css_tokenizer_encode_batch, GIL released) while the caller's thread runs
    css_encoder_encode on texts[i].  The batch kernels' results do not depend on how sequences are
    grouped into passes, so the output is bit-identical to the one-shot call; a trailing slab of a
    single text is merged into its predecessor (one sequence alone would take the query-graph path)."""
    n = len(texts)
    if n <= slab + slab // 2:
        ids, cu = tokenizer.encode_packed(texts, max_len)
        return encoder.encode_packed(ids, cu, normalize=normalize)
    from concurrent.futures import ThreadPoolExecutor
    starts = list(range(0, n, slab))
    if n - starts[-1] < max(2, slab // 2):
        starts.pop()
    bounds = list(zip(starts, starts[1:] + [n]))
    out = np.empty((n, encoder.dim), np.float32)
    with ThreadPoolExecutor(max_workers=1) as pool:
        fut = pool.submit(tokenizer.encode_packed, texts[bounds[0][0]:bounds[0][1]], max_len)
        for i, (b0, b1) in enumerate(bounds):
            ids, cu = fut.result()
            if i + 1 < len(bounds):
                fut = pool.submit(tokenizer.encode_packed, texts[bounds[i + 1][0]:bounds[i + 1][1]], max_len)
            out[b0:b1] = encoder.encode_packed(ids, cu, normalize=normalize)
    return out


class SentenceTransformer:
    def __init__(self, model_name_or_path: str = "all-mpnet-base-v2", cache_folder: Optional[str] = None,
                 device: Optional[str] = None, devices: Optional[Sequence[int]] = None, **_unused):
        """`devices` (extension, EmbeddingConfig.devices): several GPUs of this box -- the weights are replicated and
        every encode() call is split data-parallel over them (MultiDeviceEncoder).  The encoder is built on first
        use, so that `.to("cuda:N")` (src/embeddings.py:94) decides where without loading the weights twice."""
        self.model_name = model_name_or_path
        self.max_seq_length = 384
        self.tokenize_slab = 1024   # texts per tokeniser slab when tokenising overlaps encoding
        self._device_index = _device_index_of(device)
        self._devices = [int(d) for d in devices] if devices else None
        if self._devices:
            self.tokenize_slab *= len(self._devices)   # every device still gets a full pass per slab
        self._enc = None
        synthetic = model_name_or_path == "synthetic-mpnet" or os.environ.get("CSS_B200_SYNTHETIC_MODEL") == "1"
        model_dir = None if model_name_or_path == "synthetic-mpnet" else _find_model_dir(model_name_or_path, cache_folder)
        self._model_dir = model_dir
        if model_dir is not None:
            if not ((model_dir / "model.safetensors").exists() or (model_dir / "pytorch_model.bin").exists()):
                raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin under {model_dir}")
            self.tokenizer = self._load_tokenizer(model_dir)
            if self.tokenizer is None and (model_dir / "tokenizer.json").exists():
                self.tokenizer = FastTokenizer(model_dir / "tokenizer.json")   # a pipeline the native one does not implement
            if self.tokenizer is None:
                raise FileNotFoundError(f"{model_dir}: neither tokenizer.json nor vocab.txt")
            self.synthetic = False
        elif synthetic:
            # benchmarks of the text path: random-init weights with a real WordPiece vocabulary file
            vocab = os.environ.get("CSS_B200_SYNTHETIC_VOCAB")
            self.tokenizer = NativeWordPieceTokenizer(vocab) if vocab else StandInTokenizer(DEFAULT_CONFIG["vocab_size"])
            self.synthetic = True
        else:
            raise FileNotFoundError(
                f"model '{model_name_or_path}' not found locally (cache_folder={cache_folder!r}); this build does "
                "not download.  Point cache_folder at a directory holding the checkpoint, or set "
                "CSS_B200_SYNTHETIC_MODEL=1 for random-init weights + stand-in tokenizer (benchmarks only).")
        self._encoder   # fail now, not at the first encode(), when the device or the checkpoint is unusable

    def _make_encoder(self, device: int) -> MPNetEncoder:
        if self._model_dir is not None:
            return MPNetEncoder.from_pretrained(self._model_dir, device=device)
        return MPNetEncoder(random_state_dict(0), device=device)

    @property
    def _encoder(self):
        if self._enc is None:
            if self._devices and len(self._devices) > 1:
                self._enc = MultiDeviceEncoder.create(self._make_encoder, self._devices)
            else:
                self._enc = self._make_encoder(self._devices[0] if self._devices else self._device_index)
        return self._enc

    @staticmethod
    def _load_tokenizer(model_dir: Path):
        """The native tokenizer whenever the checkpoint's tokenizer is the pipeline it implements
        (tokenizer.json inspected when present; vocab.txt alone means MPNetTokenizer defaults)."""
        vocab_txt, tj = model_dir / "vocab.txt", model_dir / "tokenizer.json"
        if os.environ.get("CSS_B200_NATIVE_TOKENIZER", "1") == "0" and tj.exists():
            return None
        if tj.exists():
            st = native_tokenizer_settings(tj)
            if st is None or not vocab_txt.exists():
                return None
            with open(vocab_txt, encoding="utf-8") as fh:
                same = all(st["vocab"].get(line.rstrip("\n")) == i for i, line in enumerate(fh))
            return NativeWordPieceTokenizer(vocab_txt, st["lower"], specials=st["specials"]) if same else None
        if vocab_txt.exists():
            lower = True
            tc = model_dir / "tokenizer_config.json"
            if tc.exists():
                import json
                lower = bool(json.loads(tc.read_text(encoding="utf-8")).get("do_lower_case", True))
            return NativeWordPieceTokenizer(vocab_txt, lower)
        return None

    # -- API surface the reference uses -----------------------------------------------
    def to(self, device) -> "SentenceTransformer":
        s = str(device)
        if s.startswith("cpu"):
            # the reference moves to "cpu" when use_gpu is False; this path has no CPU encoder
            return self
        idx = _device_index_of(s)
        if not self._devices and idx != self._device_index:
            self._device_index = idx
            if self._enc is not None:   # already resident elsewhere: move = reload there
                self._enc.close()
                self._enc = None
            self._encoder
        return self

    @property
    def device(self) -> str:
        return f"cuda:{self._devices[0] if self._devices else self._device_index}"

    def get_sentence_embedding_dimension(self) -> int:
        return self._encoder.dim

    def tokenize_ids(self, sentences: Sequence[str]) -> List[List[int]]:
        max_len = min(int(self.max_seq_length), self._encoder.max_seq_len)
        return self.tokenizer.encode_batch(sentences, max_len)

    def encode(self, sentences, batch_size: int = 32, show_progress_bar=None, convert_to_numpy: bool = True,
               normalize_embeddings: bool = False, **_unused):
        single = isinstance(sentences, str)
        texts = [sentences] if single else list(sentences)
        if not texts:
            return np.zeros((0, self._encoder.dim), np.float32)
        # batch_size is a host-memory knob of the reference; the device path packs whole
        # passes of up to max_tokens tokens, results do not depend on it
        if hasattr(self.tokenizer, "encode_packed"):   # native tokenizer: straight to packed ids, no Python lists
            max_len = min(int(self.max_seq_length), self._encoder.max_seq_len)
            emb = encode_texts_pipelined(self.tokenizer, self._encoder, texts, max_len, normalize_embeddings,
                                         self.tokenize_slab)
        else:
            emb = self._encoder.encode_ids(self.tokenize_ids(texts), normalize=normalize_embeddings)
        return emb[0] if single else emb

    def close(self) -> None:
        if self._enc is not None:
            self._enc.close()
            self._enc = None
