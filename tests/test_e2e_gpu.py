"""BASELINE config 1 end to end on the GPU: synthetic Claude-style chunks -> EmbeddingGenerator
(MPNet on the device, random-init weights + stand-in tokenizer) -> HybridStorage.add_chunks ->
100 queries top-10, compared with the oracle pipeline (transformers MPNetModel fp32 on the CPU
with the same weights + the flat-IP / filter oracle)."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WORDS = ("lorem ipsum dolor sit amet consectetur adipiscing elit sed do eiusmod tempor incididunt ut labore et dolore "
         "magna aliqua function class return import def async await error trace index vector search kernel stream "
         "memory cache token query result project session file path config test build deploy").split()


def _make_chunks(n, seed=1234):
    from claude_semantic_search_b200 import Chunk
    r = random.Random(seed)
    chunks = []
    for i in range(n):
        text = " ".join(r.choice(WORDS) for _ in range(r.randint(8, 160)))
        has_code = r.random() < 0.4
        if has_code:
            text += "\n```python\nprint('x')\n```"
        chunks.append(Chunk(id=f"chunk_{i:05d}", text=text, metadata=dict(
            session_id=f"sess-{i % 40}", project_name=r.choice(["/home/u/alpha", "/home/u/Beta-Proj", "/srv/gamma"]),
            file_path=f"/f/{i % 40}.jsonl", chunk_type=r.choice(["qa_pair", "code_block", "context_segment"]),
            timestamp=f"2024-{1 + i % 12:02d}-{1 + i % 28:02d}T10:00:00+00:00", has_code=has_code, has_tools=False,
            message_count=2, char_count=len(text), word_count=len(text.split()))))
    return chunks


def test_config1_pipeline_vs_oracle(tmp_path):
    import torch
    from transformers import MPNetConfig, MPNetModel

    from claude_semantic_search_b200 import (EmbeddingConfig, EmbeddingGenerator, HybridStorage, SearchConfig,
                                             StorageConfig)
    from claude_semantic_search_b200.encoder import random_state_dict
    from oracle import encoder_oracle as eo
    from oracle import search_oracle as so

    n = 1000                                   # BASELINE configs[0]: ~1k chunks, 100 queries
    chunks = _make_chunks(n)
    gen = EmbeddingGenerator(EmbeddingConfig(model_name="synthetic-mpnet", use_gpu=True, show_progress=False))
    emb = gen.generate_embeddings(chunks)
    assert emb.shape == (n, 768) and emb.dtype == np.float32
    assert gen.is_model_loaded and gen.embedding_dimension == 768 and gen.is_using_gpu
    assert all(isinstance(c.embedding, list) and len(c.embedding) == 768 for c in chunks)

    # oracle encoder with the same weights and the same token ids
    model = MPNetModel(MPNetConfig(**eo.CONFIG), add_pooling_layer=False).eval()
    sd = {k: torch.from_numpy(v) for k, v in random_state_dict(0).items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in m for m in missing)
    ids = gen.model.tokenize_ids([c.text for c in chunks])
    sub = list(range(0, n, 8))                 # 125 chunks through the fp32 CPU oracle
    ref = eo.st_encode_ids(model, [ids[i] for i in sub])
    cos = eo.cosine_rows(ref, emb[sub])
    assert cos.min() >= 0.9999, f"min cosine {cos.min():.6f}"   # north_star bar

    st = HybridStorage(StorageConfig(data_dir=str(tmp_path), use_gpu=True))
    st.initialize()
    st.add_chunks(chunks)
    assert st.faiss_index.ntotal == n and st.total_chunks == n
    x = so.normalize_rows(emb)
    rows = [dict(id=c.id, **{k: (int(v) if isinstance(v, bool) else v) for k, v in c.metadata.items()}) for c in chunks]
    r = random.Random(5)
    flt = {"project_name": "beta", "has_code": True}
    for qi in r.sample(range(n), 100):
        q = gen.generate_single_embedding(chunks[qi].text[:80])
        assert q.shape == (768,) and q.dtype == np.float32
        got = st.search(q, SearchConfig(top_k=10))
        want = so.storage_search(x, rows, q, top_k=10)
        assert [g.chunk_id for g in got] == [chunks[i].id for i, _ in want] or \
            np.allclose([g.similarity for g in got], [s for _, s in want], atol=1e-4)
        np.testing.assert_allclose([g.similarity for g in got], [s for _, s in want], atol=1e-4)
        gotf = st.search(q, SearchConfig(top_k=10), filters=flt)
        wantf = so.prefilter_search(x, rows, q, top_k=10, filters=flt)
        np.testing.assert_allclose([g.similarity for g in gotf], [s for _, s in wantf], atol=1e-4)
        assert all("beta" in g.metadata["project_name"].lower() and g.metadata["has_code"] for g in gotf)
    # ndarray-view embeddings (EmbeddingConfig.embedding_as_ndarray): same vectors, same search
    gen.config.embedding_as_ndarray = True
    chunks_b = _make_chunks(n)
    emb_b = gen.generate_embeddings(chunks_b)
    np.testing.assert_array_equal(emb_b, emb)
    assert all(isinstance(c.embedding, np.ndarray) and c.embedding.dtype == np.float32 and c.embedding.base is not None
               for c in chunks_b)
    st_b = HybridStorage(StorageConfig(data_dir=str(tmp_path / "views"), use_gpu=True))
    st_b.initialize()
    st_b.add_chunks(chunks_b)
    q = gen.generate_single_embedding(chunks[7].text[:80])
    assert [(g.chunk_id, g.similarity) for g in st_b.search(q, SearchConfig(top_k=10))] == \
        [(g.chunk_id, g.similarity) for g in st.search(q, SearchConfig(top_k=10))]
    st_b.close()
    st.close()
    # persistence: a new instance sees the same index (reference tests/test_storage.py:541-558)
    st2 = HybridStorage(StorageConfig(data_dir=str(tmp_path), use_gpu=True))
    st2.initialize()
    assert st2.faiss_index.ntotal == n
    st2.close()
    gen.model.close()


def test_optimize_compacts_orphans_on_device(tmp_path):
    """remove_chunks_for_file / delete_chunk orphan rows (src/storage.py:836-846); optimize() gathers
    the survivors on the device (css_index_compact) -- searches before and after return the same
    chunks and scores, ntotal shrinks to the live count, and the compacted index persists."""
    from claude_semantic_search_b200 import HybridStorage, SearchConfig, StorageConfig

    n = 600
    chunks = _make_chunks(n, seed=77)
    rng = np.random.default_rng(3)
    emb = rng.standard_normal((n, 768)).astype(np.float32)
    for c, e in zip(chunks, emb):
        c.embedding = e
    st = HybridStorage(StorageConfig(data_dir=str(tmp_path), use_gpu=True))
    st.initialize()
    st.add_chunks(chunks)
    removed = st.remove_chunks_for_file("/f/3.jsonl") + st.remove_chunks_for_file("/f/17.jsonl")
    assert removed == 30
    assert st.delete_chunk("chunk_00000")
    live = n - removed - 1
    assert st.faiss_index.ntotal == n
    qs = rng.standard_normal((12, 768)).astype(np.float32)
    flt = {"project_name": "alpha"}
    before = [[(r.chunk_id, r.similarity) for r in st.search(q, SearchConfig(top_k=10))] for q in qs]
    before_f = [[(r.chunk_id, r.similarity) for r in st.search(q, SearchConfig(top_k=10), filters=flt)] for q in qs]
    assert all(len(b) == 10 for b in before)
    st.optimize()
    assert st.faiss_index.ntotal == live and len(st.faiss_id_to_chunk_id) == live
    assert sorted(st.faiss_id_to_chunk_id) == list(range(live))
    after = [[(r.chunk_id, r.similarity) for r in st.search(q, SearchConfig(top_k=10))] for q in qs]
    after_f = [[(r.chunk_id, r.similarity) for r in st.search(q, SearchConfig(top_k=10), filters=flt)] for q in qs]
    assert after == before and after_f == before_f        # same rows, same fp32 arithmetic: bit-identical
    # appending after compaction continues from the compacted row count
    extra = _make_chunks(5, seed=5)
    for i, c in enumerate(extra):
        c.id = f"extra_{i}"
        c.embedding = qs[i]
    st.add_chunks(extra)
    assert st.faiss_index.ntotal == live + 5
    top = st.search(qs[2], SearchConfig(top_k=1))
    assert top[0].chunk_id == "extra_2" and abs(top[0].similarity - 1.0) < 1e-4
    st.close()
    st2 = HybridStorage(StorageConfig(data_dir=str(tmp_path), use_gpu=True))
    st2.initialize()
    assert st2.faiss_index.ntotal == live + 5
    again = [[(r.chunk_id, r.similarity) for r in st2.search(q, SearchConfig(top_k=10))] for q in qs[5:]]
    assert again == before[5:]
    st2.close()


def test_local_checkpoint_text_to_embedding_vs_hf(tmp_path):
    """Text in, embedding out through the inner seam with a LOCAL checkpoint directory (config.json +
    pytorch_model.bin + vocab.txt): native WordPiece tokenizer -> packed ids -> B200 encoder, against
    transformers' BertTokenizer-style tokenisation + fp32 MPNetModel + sentence-transformers pooling."""
    import json
    import random

    import torch
    from transformers import BertTokenizer

    from claude_semantic_search_b200.st_compat import NativeWordPieceTokenizer, SentenceTransformer
    from oracle import encoder_oracle as eo
    rnd = random.Random(3)
    letters = "abcdefghijklmnopqrstuvwxyz"
    words = sorted({"".join(rnd.choice(letters) for _ in range(rnd.randint(2, 8))) for _ in range(1500)})
    vocab = ["<s>", "<pad>", "</s>", "[UNK]"] + words + ["##" + w[:3] for w in words[:400]] + list(letters) + \
        ["##" + c for c in letters] + list(".,!?()-")
    vocab = list(dict.fromkeys(vocab))
    model = eo.build_model(seed=0, perturb=True, num_layers=2)   # vocabulary ids stay below the model's 30527 rows
    d = tmp_path / "all-mpnet-base-v2"
    d.mkdir()
    (d / "vocab.txt").write_text("\n".join(vocab) + "\n", encoding="utf-8")
    cfg = model.config.to_dict()
    (d / "config.json").write_text(json.dumps({k: v for k, v in cfg.items() if isinstance(v, (int, float, str, bool))}))
    torch.save(model.state_dict(), d / "pytorch_model.bin")
    st = SentenceTransformer(str(d))
    assert isinstance(st.tokenizer, NativeWordPieceTokenizer)
    texts = [" ".join(rnd.choice(words) + rnd.choice(["", "", ",", ".", "!"]) for _ in range(rnd.randint(1, 120)))
             for _ in range(40)] + ["Hello (world)!", "", "naïve café " + words[0]]
    got = st.encode(texts, normalize_embeddings=True)
    hf = BertTokenizer(str(d / "vocab.txt"), do_lower_case=True, unk_token="[UNK]", cls_token="<s>", sep_token="</s>",
                       pad_token="<pad>")
    ids = [hf.encode(t, add_special_tokens=True, truncation=True, max_length=384) for t in texts]
    assert st.tokenize_ids(texts) == ids
    want = eo.st_encode_ids(model, ids, batch_size=16)
    cos = eo.cosine_rows(want, got)
    assert cos.min() >= 0.9999, cos.min()
    one = st.encode(texts[0], normalize_embeddings=True)
    assert one.shape == (768,) and eo.cosine_rows(want[:1], one[None])[0] >= 0.9999
    # tokeniser one slab ahead of the GPU (worker thread): bit-identical to the one-shot call, whatever the
    # slab size, including a single-text tail (merged into the previous slab)
    for slab in (4, 6, 14):
        st.tokenize_slab = slab
        np.testing.assert_array_equal(st.encode(texts, normalize_embeddings=True), got, err_msg=f"slab={slab}")
    st.close()


def test_embedding_generator_over_several_devices_matches_one_device():
    """EmbeddingConfig.devices (north_star: embedding batches split data-parallel over the GPUs of one box behind
    the unchanged generate_embeddings API): replicated weights, contiguous ranges of equal token count,
    no collective.  A sequence's embedding does not depend on which pass it travels in, so the result must equal the
    single-device one bit for bit.  On a one-GPU box the two handles share the device (the split / thread / reassembly
    logic is what is under test); `device="cuda:N"` must be honoured as well."""
    from claude_semantic_search_b200 import EmbeddingConfig, EmbeddingGenerator, _native

    n_dev = min(_native.device_count(), 4)
    devices = list(range(n_dev)) if n_dev > 1 else [0, 0]
    chunks = _make_chunks(333, seed=77)
    one = EmbeddingGenerator(EmbeddingConfig(model_name="synthetic-mpnet", use_gpu=True, show_progress=False))
    ref = np.array(one.generate_embeddings(chunks))
    many = EmbeddingGenerator(EmbeddingConfig(model_name="synthetic-mpnet", use_gpu=True, show_progress=False,
                                              devices=devices, embedding_as_ndarray=True))
    got = many.generate_embeddings(chunks)
    assert got.shape == ref.shape and got.dtype == np.float32
    assert np.array_equal(got, ref)
    assert all(isinstance(c.embedding, np.ndarray) for c in chunks)
    q1 = one.generate_single_embedding("where is the vector index kept")
    qn = many.generate_single_embedding("where is the vector index kept")
    assert np.array_equal(q1, qn)                       # a single query stays on the first device
    assert many.model.device == f"cuda:{devices[0]}" and many.is_using_gpu
    last = _native.device_count() - 1
    pinned = EmbeddingGenerator(EmbeddingConfig(model_name="synthetic-mpnet", use_gpu=True, show_progress=False,
                                                device=f"cuda:{last}"))
    pinned.load_model()
    assert pinned.model.device == f"cuda:{last}"
    assert np.array_equal(np.array(pinned.generate_embeddings(chunks[:40])), ref[:40])
