"""The reference's indexing flow on the hot path (src/cli.py:217-237): chunks with text ->
EmbeddingGenerator.generate_embeddings -> HybridStorage.add_chunks, host objects in, device index out.
Random-init weights + a synthetic WordPiece vocabulary (CSS_B200_SYNTHETIC_VOCAB); reports chunks/s for the
reference's list embeddings and for EmbeddingConfig.embedding_as_ndarray."""
import json
import os
import random
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    rnd = random.Random(17)
    letters = "abcdefghijklmnopqrstuvwxyz"
    words = sorted({"".join(rnd.choice(letters) for _ in range(rnd.randint(2, 9))) for _ in range(24000)})
    vocab = ["<s>", "<pad>", "</s>", "[UNK]"] + words + ["##" + w[:4] for w in words[:6000]] + list(letters) + \
        ["##" + c for c in letters] + list(".,!?()-:;")
    vocab = list(dict.fromkeys(vocab))[:30527]
    out = {}
    with tempfile.TemporaryDirectory() as d:
        vf = Path(d) / "vocab.txt"
        vf.write_text("\n".join(vocab) + "\n", encoding="utf-8")
        os.environ["CSS_B200_SYNTHETIC_VOCAB"] = str(vf)
        from claude_semantic_search_b200 import Chunk, EmbeddingConfig, EmbeddingGenerator, HybridStorage, StorageConfig
        texts = [" ".join(rnd.choice(words) + rnd.choice(["", "", "", ",", "."]) for _ in range(400)) for _ in range(n)]

        def chunks():
            return [Chunk(id=f"c{i:06d}", text=t, metadata=dict(
                session_id=f"s{i % 50}", project_name=f"/p/{i % 7}", file_path=f"/f/{i % 200}.jsonl", chunk_type="qa_pair",
                timestamp=f"2024-{1 + i % 12:02d}-{1 + i % 28:02d}T10:00:00+00:00", has_code=bool(i & 1), has_tools=False,
                message_count=2, char_count=len(t), word_count=400)) for i, t in enumerate(texts)]
        gen = EmbeddingGenerator(EmbeddingConfig(model_name="synthetic-mpnet", use_gpu=True, show_progress=False))
        gen.generate_embeddings(chunks()[:512])   # load + warm
        for mode in (False, True):
            gen.config.embedding_as_ndarray = mode
            st = HybridStorage(StorageConfig(data_dir=str(Path(d) / f"st{int(mode)}"), use_gpu=True, auto_save=False))
            st.initialize()
            cs = chunks()
            t0 = time.perf_counter()
            gen.generate_embeddings(cs)
            t1 = time.perf_counter()
            st.add_chunks(cs)
            t2 = time.perf_counter()
            assert st.faiss_index.ntotal == n
            out["ndarray_views" if mode else "lists (reference behaviour)"] = {
                "generate_embeddings_chunks_per_s": n / (t1 - t0), "add_chunks_chunks_per_s": n / (t2 - t1),
                "pipeline_chunks_per_s": n / (t2 - t0)}
            st.close()
    out["chunks"] = n
    out["chars_per_chunk"] = sum(len(t) for t in texts) // n
    print(json.dumps(out))


if __name__ == "__main__":
    main()
