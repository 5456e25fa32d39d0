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
    on ASCII text; anything else is flagged and tokenised by the Python path (still == transformers)."""
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
    texts += ["Unbelievable café", "naïve résumé", "日本語 text", "ctrl\x0bchar", "zero\x00byte"]   # out of the native scope
    for max_len in (16, 64, 384):
        got = native.encode_batch(texts, max_len)
        assert got == py.encode_batch(texts, max_len)
        for t, g in zip(texts[:120] + texts[-11:], got[:120] + got[-11:]):
            if "\x0b" in t or "\x00" in t:
                continue   # transformers drops these control characters; the Python path keeps its documented behaviour
            assert g == hf.encode(t, add_special_tokens=True, truncation=True, max_length=max_len), t
    ids, cu = native.encode_packed(texts, 384)
    assert cu[0] == 0 and cu[-1] == ids.shape[0] and (np.diff(cu) >= 2).all()
    big = [text() * 8 for _ in range(4000)]
    t0 = time.perf_counter()
    ids, cu = native.encode_packed(big, 384)
    dt = time.perf_counter() - t0
    print(f"native tokenizer: {len(big) / dt:.0f} texts/s, {ids.shape[0] / dt / 1e6:.1f} M tokens/s")
    native.close()


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
