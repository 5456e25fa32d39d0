"""CPU-only host-logic tests for the outer seam of half A (EmbeddingGenerator mirror,
tokenisers, capability heuristics)."""
import numpy as np
import pytest


def test_calculate_optimal_batch_size_matches_reference_formula():
    # reference src/gpu_utils.py:169-192
    from claude_semantic_search_b200.gpu_utils import calculate_optimal_batch_size as f
    assert f(0.5) == 8
    assert f(1.0) == 8
    assert f(1.0 + 100 * 768 * 16 / 1024 ** 3) in (99, 100)
    assert f(80.0) == 256
    assert f(80.0, backend="mps") == 64


def test_wordpiece_matches_transformers_bert(tmp_path):
    from transformers import BertTokenizer

    from claude_semantic_search_b200.st_compat import WordPieceTokenizer
    vocab = ["<s>", "<pad>", "</s>", "[UNK]", "hello", "world", "##s", "un", "##believ", "##able", ",", "!", "cafe",
             "new", "york", "##er", "a", "b", "##c", "1", "##2", "##3", "."]
    vf = tmp_path / "vocab.txt"
    vf.write_text("\n".join(vocab) + "\n", encoding="utf-8")
    mine = WordPieceTokenizer(vf)
    hf = BertTokenizer(str(vf), do_lower_case=True, unk_token="[UNK]", cls_token="<s>", sep_token="</s>",
                       pad_token="<pad>")
    texts = ["Hello, worlds!", "Unbelievable  café", "New Yorker abc 123.", "zzz hello", ""]
    for t in texts:
        want = hf.encode(t, add_special_tokens=True)
        got = mine.encode_batch([t], max_length=64)[0]
        assert got == want, (t, got, want)
    assert len(mine.encode_batch(["hello " * 100], max_length=16)[0]) == 16


def test_native_wordpiece_matches_python_and_transformers(tmp_path):
    """css_tokenizer_encode_batch (C++, multi-threaded) == WordPieceTokenizer == transformers BertTokenizer
    on ASCII and simple accented / CJK text (the full Unicode sweep is the next test)."""
    import random
    import time

    from transformers import BertTokenizer

    from claude_semantic_search_b200.st_compat import NativeWordPieceTokenizer, WordPieceTokenizer
    rnd = random.Random(5)
    letters = "abcdefghijklmnopqrstuvwxyz0123456789"
    words = {"".join(rnd.choice(letters) for _ in range(rnd.randint(1, 7))) for _ in range(3000)}
    pieces = {"##" + "".join(rnd.choice(letters) for _ in range(rnd.randint(1, 4))) for _ in range(2000)}
    vocab = ["<s>", "<pad>", "</s>", "[UNK]"] + sorted(words) + sorted(pieces) + list("abcdefghijklmnopqrstuvwxyz") + \
        ["##" + c for c in letters] + list("!\"#$%&'()*+,-./:;<=>?@[]^_`{|}~") + ["cafe"]
    vocab = list(dict.fromkeys(vocab))
    vf = tmp_path / "vocab.txt"
    vf.write_text("\n".join(vocab) + "\n", encoding="utf-8")
    native = NativeWordPieceTokenizer(vf, n_threads=4)
    py = WordPieceTokenizer(vf)
    hf = BertTokenizer(str(vf), do_lower_case=True, unk_token="[UNK]", cls_token="<s>", sep_token="</s>",
                       pad_token="<pad>")
    wl = sorted(words)

    def text():
        out = []
        for _ in range(rnd.randint(0, 60)):
            r = rnd.random()
            if r < 0.6:
                w = rnd.choice(wl)
            elif r < 0.8:
                w = rnd.choice(wl) + rnd.choice(wl)[:3]
            elif r < 0.9:
                w = "".join(rnd.choice(letters + "ABCXYZ") for _ in range(rnd.randint(1, 12)))
            else:
                w = rnd.choice("!,.;:()[]{}-_/\\'\"@#")
            out.append(w)
            out.append(rnd.choice([" ", " ", " ", "", "\n", "\t", "  "]))
        return "".join(out)

    texts = [text() for _ in range(400)] + ["", " ", "x" * 150, "Hello, World!", "a" * 101 + " b", "tab\tsep\r\nline"]
    texts += ["Unbelievable café", "naïve résumé", "日本語 text"]
    ctrl = ["ctrl\x0bchar", "zero\x00byte"]   # control characters are dropped (the Python stand-in keeps them)
    for max_len in (16, 64, 384):
        got = native.encode_batch(texts, max_len)
        assert got == py.encode_batch(texts, max_len)
        got_c = native.encode_batch(ctrl, max_len)
        for t, g in zip(texts[:120] + texts[-9:] + ctrl, got[:120] + got[-9:] + got_c):
            assert g == hf.encode(t, add_special_tokens=True, truncation=True, max_length=max_len), t
    ids, cu = native.encode_packed(texts, 384)
    assert cu[0] == 0 and cu[-1] == ids.shape[0] and (np.diff(cu) >= 2).all()
    big = [text() * 8 for _ in range(4000)]
    t0 = time.perf_counter()
    ids, cu = native.encode_packed(big, 384)
    dt = time.perf_counter() - t0
    print(f"native tokenizer: {len(big) / dt:.0f} texts/s, {ids.shape[0] / dt / 1e6:.1f} M tokens/s")
    native.close()


def _reference_fast_tokenizer(vocab, lower, specials):
    """The pipeline behind MPNetTokenizerFast (what sentence-transformers' AutoTokenizer gives the
    reference, src/embeddings.py:136-151), assembled from the `tokenizers` package."""
    from tokenizers import AddedToken, Tokenizer
    from tokenizers.models import WordPiece
    from tokenizers.normalizers import BertNormalizer
    from tokenizers.pre_tokenizers import BertPreTokenizer
    from tokenizers.processors import TemplateProcessing
    ids = {w: i for i, w in enumerate(vocab)}
    tok = Tokenizer(WordPiece(ids, unk_token="[UNK]", max_input_chars_per_word=100))
    tok.normalizer = BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=None, lowercase=lower)
    tok.pre_tokenizer = BertPreTokenizer()
    tok.post_processor = TemplateProcessing(single="<s> $A </s>", special_tokens=[("<s>", ids["<s>"]), ("</s>", ids["</s>"])])
    tok.add_special_tokens([AddedToken(s, lstrip=(s == "<mask>")) for s in specials])
    return tok


def test_native_tokenizer_equals_fast_pipeline_on_all_of_unicode(tmp_path):
    """Every code point (inside a word, doubled, and after a space), random mixed-script strings, the
    NFD reordering corner, added special tokens in the raw text, truncation -- identical ids to the
    `tokenizers` BertNormalizer / BertPreTokenizer / WordPiece pipeline, in both case modes; then the
    tokenizer.json inspection that decides whether a checkpoint may use the native path."""
    import random

    from tokenizers.normalizers import BertNormalizer

    from claude_semantic_search_b200.st_compat import (MPNET_SPECIALS, NativeWordPieceTokenizer, SentenceTransformer,
                                                       native_tokenizer_settings)
    rnd = random.Random(1)
    allcps = [cp for cp in range(0x110000) if not 0xD800 <= cp <= 0xDFFF]
    norm = BertNormalizer(lowercase=True)
    chars = set()
    for cp in allcps[::7] + list(range(0x20, 0x3000)):
        chars.update(norm.normalize_str(chr(cp)))
    chars.discard(" ")
    keep = [c for c in sorted(chars) if rnd.random() < 0.5]
    vocab = ["<s>", "<pad>", "</s>", "<unk>"] + [f"[unused{i}]" for i in range(10)] + ["[UNK]"] + keep + \
        ["##" + c for c in keep if rnd.random() < 0.7] + ["cafe", "##fe", "naive", "resume", "hello", "world", "##llo", "<mask>"]
    vocab = list(dict.fromkeys(vocab))
    vf = tmp_path / "vocab.txt"
    vf.write_text("\n".join(vocab) + "\n", encoding="utf-8")
    pools = [range(0x20, 0x7F), range(0xA0, 0x250), range(0x300, 0x370), range(0x370, 0x530), range(0x590, 0x700),
             range(0x900, 0xA00), range(0xE00, 0xE80), range(0x1100, 0x1200), range(0x2000, 0x2070),
             range(0x3000, 0x3100), range(0x4E00, 0x4F00), range(0xAC00, 0xAD00), range(0xF900, 0xFA00),
             range(0xFF00, 0xFFF0), range(0x1F600, 0x1F650), range(0x1D160, 0x1D175), range(0x302A, 0x3030),
             range(0, 0x20), range(0x1B00, 0x1C00), range(0x20000, 0x20100)]
    mixed = []
    for _ in range(6000):
        s = []
        for _ in range(rnd.randint(0, 40)):
            pool = pools[0] if rnd.random() < 0.4 else rnd.choice(pools)
            s.append(chr(rnd.choice(pool)))
            if rnd.random() < 0.15:
                s.append(" ")
            if rnd.random() < 0.03:
                s.append(rnd.choice(MPNET_SPECIALS + ("<S>", "<mas", "[unk]")))
        mixed.append("".join(s))
    mixed += ["", " ", "Unbelievable café", "naïve résumé", "日本語 text", "ctrl\x0bchar", "zero\x00byte", "x" * 150,
              "é" * 101, "é" * 100 + " ok", "İstanbul ΣΊΣΥΦΟΣ ß ǅ", "a〮ୖ\U0001d165b", "á〮\U0001d165b",
              "한국어 텍스트", "😀 emoji 👍🏽 zwj 👨‍👩‍👧", "a<s>b</s>c", "x <mask> y<mask>z", "<<s>>", "[UNK][UNK] <pad>", "<unk>hello",
              "hello<mask", "\ufeffbom\u200bzero\u00adwidth", "\u2028line\u2029para\u00a0nbsp\u3000wide"]
    sweep = [" ".join("a" + chr(c) + "b " + chr(c) + chr(c) for c in allcps[i:i + 16]) for i in range(0, len(allcps), 16)]
    for lower in (True, False):
        ref = _reference_fast_tokenizer(vocab, lower, MPNET_SPECIALS)
        nat = NativeWordPieceTokenizer(vf, do_lower_case=lower, n_threads=4)
        assert nat.specials == {s: vocab.index(s) for s in MPNET_SPECIALS}
        for texts, lens in ((sweep, (384,)), (mixed, (8, 64, 384))):
            for ml in lens:
                ref.enable_truncation(max_length=ml)
                want = [e.ids for e in ref.encode_batch(texts)]
                ids, cu = nat.encode_packed(texts, ml)
                bad = [i for i, w in enumerate(want) if ids[cu[i]:cu[i + 1]].tolist() != w]
                assert not bad, (lower, ml, len(bad), repr(texts[bad[0]]), ids[cu[bad[0]]:cu[bad[0] + 1]].tolist(), want[bad[0]])
        nat.close()
        # tokenizer.json inspection: the same pipeline is accepted, with its case mode and added tokens
        mdir = tmp_path / f"ckpt_{int(lower)}"
        mdir.mkdir()
        ref.no_truncation()
        ref.save(str(mdir / "tokenizer.json"))
        st = native_tokenizer_settings(mdir / "tokenizer.json")
        assert st is not None and st["lower"] is lower and st["specials"] == {s: vocab.index(s) for s in MPNET_SPECIALS}
        (mdir / "vocab.txt").write_text("\n".join(vocab) + "\n", encoding="utf-8")
        tk = SentenceTransformer._load_tokenizer(mdir)
        assert isinstance(tk, NativeWordPieceTokenizer)
        assert tk.encode_batch(["Héllo <mask> wörld"], 16) == [ref.encode("Héllo <mask> wörld").ids]
        tk.close()
    # malformed UTF-8 (overlong, lone continuation, surrogate, truncated, > U+10FFFF) is flagged by the library, never tokenised
    import ctypes
    nat = NativeWordPieceTokenizer(vf, n_threads=1)
    raws = [b"ok text", b"bad \xc0\xaf", b"\x80lone", b"sur \xed\xa0\x80", b"cut \xe2\x82", b"big \xf4\x90\x80\x80", "fine é 日本".encode()]
    n = len(raws)
    ptrs = (ctypes.c_char_p * n)(*raws)
    lens = np.asarray([len(b) for b in raws], np.int64)
    ids, cu, fb = np.empty(n * 16, np.int32), np.zeros(n + 1, np.int32), np.zeros(n, np.uint8)
    assert nat._lib.css_tokenizer_encode_batch(nat._h, ctypes.cast(ptrs, ctypes.c_void_p), lens.ctypes.data, n, 16,
                                               ids.ctypes.data, cu.ctypes.data, fb.ctypes.data, 1) == 0
    assert fb.tolist() == [0, 1, 1, 1, 1, 1, 0]
    assert [int(cu[i + 1] - cu[i]) for i in (1, 2, 3, 4, 5)] == [0] * 5 and cu[1] - cu[0] >= 2
    nat.close()
    # a different pipeline (no CJK padding) must NOT be taken over by the native tokenizer
    import json
    cfg = json.loads((tmp_path / "ckpt_1" / "tokenizer.json").read_text(encoding="utf-8"))
    cfg["normalizer"]["handle_chinese_chars"] = False
    other = tmp_path / "other"
    other.mkdir()
    (other / "tokenizer.json").write_text(json.dumps(cfg), encoding="utf-8")
    (other / "vocab.txt").write_text("\n".join(vocab) + "\n", encoding="utf-8")
    assert native_tokenizer_settings(other / "tokenizer.json") is None
    assert SentenceTransformer._load_tokenizer(other) is None


def test_tokenizer_json_of_the_mpnet_layout_is_accepted(tmp_path):
    """The layout all-mpnet-base-v2 ships (RobertaProcessing post-processor, six added special tokens with
    lstrip on <mask>) is recognised as the native pipeline; near misses are not."""
    import json

    from claude_semantic_search_b200.st_compat import native_tokenizer_settings
    vocab = {"<s>": 0, "<pad>": 1, "</s>": 2, "<unk>": 3, "[UNK]": 4, "hello": 5, "<mask>": 6}
    added = [{"id": i, "content": c, "single_word": False, "lstrip": c == "<mask>", "rstrip": False, "normalized": False,
              "special": True} for c, i in (("<s>", 0), ("<pad>", 1), ("</s>", 2), ("<unk>", 3), ("[UNK]", 4), ("<mask>", 6))]
    cfg = {"version": "1.0", "truncation": {"max_length": 384}, "padding": None, "added_tokens": added,
           "normalizer": {"type": "BertNormalizer", "clean_text": True, "handle_chinese_chars": True, "strip_accents": None,
                          "lowercase": True},
           "pre_tokenizer": {"type": "BertPreTokenizer"},
           "post_processor": {"type": "RobertaProcessing", "sep": ["</s>", 2], "cls": ["<s>", 0], "trim_offsets": True,
                              "add_prefix_space": False},
           "decoder": {"type": "WordPiece", "prefix": "##", "cleanup": True},
           "model": {"type": "WordPiece", "unk_token": "[UNK]", "continuing_subword_prefix": "##",
                     "max_input_chars_per_word": 100, "vocab": vocab}}
    f = tmp_path / "tokenizer.json"

    def settings(mut=None):
        c = json.loads(json.dumps(cfg))
        if mut:
            mut(c)
        f.write_text(json.dumps(c), encoding="utf-8")
        return native_tokenizer_settings(f)
    st = settings()
    assert st is not None and st["lower"] is True and st["specials"] == {a["content"]: a["id"] for a in added}
    assert settings(lambda c: c["normalizer"].update(lowercase=False))["lower"] is False
    assert settings(lambda c: c["normalizer"].update(strip_accents=False)) is None       # accents kept while lower-casing
    assert settings(lambda c: c["normalizer"].update(type="NFKC")) is None
    assert settings(lambda c: c["pre_tokenizer"].update(type="Whitespace")) is None
    assert settings(lambda c: c["model"].update(max_input_chars_per_word=200)) is None
    assert settings(lambda c: c["model"].update(continuing_subword_prefix="@@")) is None
    assert settings(lambda c: c["post_processor"].update(cls=["[CLS]", 0])) is None
    assert settings(lambda c: c["added_tokens"][0].update(single_word=True)) is None
    assert settings(lambda c: c["added_tokens"].append({"id": 5, "content": "hello", "single_word": False, "lstrip": False,
                                                        "rstrip": False, "normalized": True, "special": False})) is None
    f.write_text("{not json", encoding="utf-8")
    assert native_tokenizer_settings(f) is None


def test_pipelined_text_encoding_slabs():
    """encode_texts_pipelined: every text is tokenised and encoded exactly once, in order, whatever the slab size;
    a short tail is merged into the previous slab (a slab of one text would take the single-query path)."""
    from claude_semantic_search_b200.st_compat import encode_texts_pipelined

    class Tok:
        def __init__(self):
            self.calls = []

        def encode_packed(self, texts, max_len):
            self.calls.append(len(texts))
            ids = np.asarray([int(t) for t in texts], np.int32)
            return ids, np.arange(len(texts) + 1, dtype=np.int32)

    class Enc:
        dim = 3

        def encode_packed(self, ids, cu, normalize=True):
            assert cu[-1] == ids.shape[0]
            return np.stack([ids, ids * 2, ids * 3], axis=1).astype(np.float32)
    for n, slab in ((0, 4), (1, 4), (5, 4), (6, 4), (7, 4), (9, 4), (100, 7), (1000, 64), (1025, 1024), (4096, 1024)):
        tok = Tok()
        texts = [str(i) for i in range(n)]
        out = encode_texts_pipelined(tok, Enc(), texts, 384, True, slab)
        assert out.shape == (n, 3)
        np.testing.assert_array_equal(out[:, 0], np.arange(n, dtype=np.float32))
        assert sum(tok.calls) == n and (n < 2 or min(tok.calls) >= 2), (n, slab, tok.calls)


def test_standin_tokenizer_is_deterministic_and_bounded():
    from claude_semantic_search_b200.st_compat import StandInTokenizer
    t = StandInTokenizer()
    a = t.encode_batch(["Some text, with punctuation.", "x " * 1000], 384)
    assert a == t.encode_batch(["Some text, with punctuation.", "x " * 1000], 384)
    assert a[0][0] == 0 and a[0][-1] == 2 and len(a[1]) == 384
    assert all(4 <= i < 30527 for s in a for i in s[1:-1])


def test_embedding_generator_surface_without_gpu(gpu_available):
    from claude_semantic_search_b200 import Chunk, EmbeddingConfig, EmbeddingGenerator, _native
    g = EmbeddingGenerator(EmbeddingConfig(model_name="synthetic-mpnet", use_gpu=True))
    assert not g.is_model_loaded and g.embedding_dimension is None and not g.is_using_gpu
    assert g.get_model_info() == {}
    assert g.get_embedding_stats([]).total_chunks == 0
    a, b = np.array([1.0, 0.0]), np.array([1.0, 1.0])
    assert abs(g.compute_similarity(a, b) - 2 ** -0.5) < 1e-12
    assert g.find_similar_chunks(a, [b, a], top_k=1)[0][0] == 1
    c = Chunk(id="c", text="t", metadata={}, embedding=[0.6, 0.8])
    v = g.validate_embeddings([c, Chunk(id="d", text="u")])
    assert v["chunks_with_embeddings"] == 1 and v["issues"] == ["Missing embedding for chunk d"]
    if not gpu_available:
        with pytest.raises(_native.NativeError):
            g.load_model()                       # no CPU fallback
        with pytest.raises(_native.NativeError):
            g.generate_embeddings([c])


def test_missing_checkpoint_is_an_error_not_a_download(tmp_path, monkeypatch):
    from claude_semantic_search_b200.st_compat import _find_model_dir
    monkeypatch.delenv("CSS_B200_SYNTHETIC_MODEL", raising=False)
    assert _find_model_dir("all-mpnet-base-v2", str(tmp_path)) is None
    d = tmp_path / "all-mpnet-base-v2"
    d.mkdir()
    (d / "model.safetensors").write_bytes(b"")
    assert _find_model_dir("sentence-transformers/all-mpnet-base-v2", str(tmp_path)) == d


def test_embedding_helpers_accept_ndarray_views(tmp_path):
    """save / validate follow `if c.embedding` of the reference (src/embeddings.py:249,274) and must not trip
    over the ndarray-view embeddings of EmbeddingConfig.embedding_as_ndarray."""
    from claude_semantic_search_b200 import Chunk, EmbeddingGenerator
    g = EmbeddingGenerator()
    v = np.ones(768, dtype=np.float32) / np.sqrt(768.0)
    chunks = [Chunk(id="a", text="x", metadata={}, embedding=v),
              Chunk(id="b", text="y", metadata={}, embedding=v.tolist()),
              Chunk(id="c", text="z", metadata={}, embedding=None)]
    res = g.validate_embeddings(chunks)
    assert res["chunks_with_embeddings"] == 2 and res["embedding_dimension"] == 768
    assert any("Missing embedding for chunk c" in s for s in res["issues"])
    path = str(tmp_path / "e.npz")
    g.save_embeddings(chunks, path)
    assert [c.id for c in g.load_embeddings(path)] == ["a", "b"]


def test_split_by_tokens_is_contiguous_ordered_and_balanced():
    """static contiguous split of the sequences over the devices, by token count."""
    from claude_semantic_search_b200.encoder import split_by_tokens
    rng = np.random.default_rng(5)
    for n, parts in [(0, 4), (1, 4), (3, 8), (100, 1), (100, 2), (257, 8), (4096, 8)]:
        lens = rng.integers(1, 385, size=n)
        cu = np.zeros(n + 1, np.int64)
        cu[1:] = np.cumsum(lens)
        ranges = split_by_tokens(cu, parts)
        assert len(ranges) == parts and ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a <= b for a, b in ranges) and all(ranges[i][1] == ranges[i + 1][0] for i in range(parts - 1))
        if n >= 64 * parts:   # balanced to within one sequence
            tok = [int(cu[b] - cu[a]) for a, b in ranges]
            assert max(tok) - min(tok) <= 2 * 384


def test_multi_device_encoder_splits_and_reassembles_in_order():
    """MultiDeviceEncoder over stand-in per-device encoders: every sequence is encoded exactly once, on the device
    that owns its range, the result rows come back in input order, and small calls stay on the first device."""
    from claude_semantic_search_b200.encoder import MultiDeviceEncoder

    class Fake:
        dim, max_tokens, max_seq_len, config = 4, 1 << 20, 384, {}

        def __init__(self, device):
            self.device, self.calls = device, []

        def encode_packed(self, ids, cu, normalize=True):
            assert cu[0] == 0 and cu[-1] == ids.shape[0]
            self.calls.append(cu.shape[0] - 1)
            out = np.empty((cu.shape[0] - 1, 4), np.float32)
            for i in range(cu.shape[0] - 1):
                seg = ids[cu[i]:cu[i + 1]]
                out[i] = (seg.sum(), seg.shape[0], seg[0], self.device)
            return out

        def close(self):
            pass

    fakes = [Fake(d) for d in range(4)]
    enc = MultiDeviceEncoder(fakes, min_seqs_per_device=8)
    rng = np.random.default_rng(0)
    seqs = [rng.integers(4, 30000, size=int(l)).astype(np.int32) for l in rng.integers(2, 385, size=203)]
    got = enc.encode_ids(seqs)
    assert got.shape == (203, 4)
    for i, s in enumerate(seqs):
        assert got[i, 0] == np.float32(s.sum()) and got[i, 1] == len(s) and got[i, 2] == s[0]
    assert (np.diff(got[:, 3]) >= 0).all() and set(got[:, 3]) == {0.0, 1.0, 2.0, 3.0}   # contiguous ranges, all devices
    assert sum(sum(f.calls) for f in fakes) == 203
    few = enc.encode_ids(seqs[:5])                       # a handful of sequences: first device only
    assert (few[:, 3] == 0).all()
    enc.close()
