"""EmbeddingGenerator: the reference's embedding API (src/embeddings.py) on the B200 path.

Outer seam of the drop-in: same dataclasses, constructor, method
names, return types and side effects (`chunk.embedding = row.tolist()`) as the reference,
so SemanticSearchCLI / the watcher / the MCP server call it unchanged.  The model is the
css_encoder_* MPNet forward behind st_compat.SentenceTransformer; there is no CPU fallback:
`load_model()` raises when no sm_100 device (or no local checkpoint) is present.
"""
from __future__ import annotations

import logging
import os
import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np

from . import _native
from .chunk import Chunk
from .gpu_utils import assess_gpu_capability, calculate_optimal_batch_size, log_gpu_status
from .st_compat import SentenceTransformer


@dataclass
class EmbeddingConfig:
    """src/embeddings.py:28-40."""

    model_name: str = "all-mpnet-base-v2"
    batch_size: int = 16
    max_seq_length: int = 384
    device: str = "auto"
    use_gpu: bool = False
    auto_batch_size: bool = True
    normalize_embeddings: bool = True
    show_progress: bool = True
    cache_dir: Optional[str] = None
    # Extension: False = the reference's `chunk.embedding = row.tolist()`
    # (src/embeddings.py:174-175); True = `chunk.embedding` is a float32 ndarray view of the result row.
    # 768 Python floats per chunk cost 0.76 s per 10 k chunks (tolist + the asarray in add_chunks) against
    # 0.83 s of GPU time: the list round trip halves the indexing rate.  HybridStorage.add_chunks takes both.
    embedding_as_ndarray: bool = False
    # Extension (north_star: embedding batches split data-parallel over the GPUs of one box, same API): the GPUs to
    # use, e.g. [0, 1, ..., 7].  The weights are replicated, every generate_embeddings call is split into contiguous
    # ranges of equal token count, one per device, no collective; None = the one device of `device`.
    devices: Optional[List[int]] = None


@dataclass
class EmbeddingStats:
    """src/embeddings.py:43-52."""

    total_chunks: int = 0
    total_tokens: int = 0
    generation_time: float = 0.0
    average_chunk_length: float = 0.0
    throughput_chunks_per_second: float = 0.0
    model_info: Dict[str, Any] = field(default_factory=dict)


def _has_embedding(c: Chunk) -> bool:
    """`if c.embedding` of the reference (src/embeddings.py:249,274), also for ndarray-view embeddings
    (EmbeddingConfig.embedding_as_ndarray), whose truth value is ambiguous."""
    return c.embedding is not None and len(c.embedding) > 0


class EmbeddingGenerator:
    def __init__(self, config: Optional[EmbeddingConfig] = None):
        self.config = config or EmbeddingConfig()
        self.model: Optional[SentenceTransformer] = None
        self.logger = logging.getLogger(__name__)
        self._embedding_dim: Optional[int] = None
        self._gpu_capability = None
        # the device probe is deferred to load_model(): constructing must work on a box
        # without a GPU (the reference's tests construct generators freely)

    # ------------------------------------------------------------------ model
    def load_model(self) -> None:
        try:
            self.logger.info("Loading model: %s", self.config.model_name)
            cache_dir = self.config.cache_dir
            if cache_dir:
                os.environ["SENTENCE_TRANSFORMERS_HOME"] = cache_dir
            self._gpu_capability = assess_gpu_capability()
            if not self._gpu_capability.can_use_gpu:
                raise _native.NoDeviceError(_native.CSS_ERR_NO_DEVICE, self._gpu_capability.status_message)
            target = self._determine_target_device()
            # the shim loads the weights where they will live (the reference constructs on the CPU and moves)
            self.model = SentenceTransformer(self.config.model_name, cache_folder=cache_dir, device=target,
                                             devices=self.config.devices)
            self.model.to(target)
            self.model.max_seq_length = self.config.max_seq_length
            if self.config.use_gpu and self.config.auto_batch_size and self._gpu_capability.gpu_memory_free:
                self.config.batch_size = calculate_optimal_batch_size(self._gpu_capability.gpu_memory_free / 1024 ** 3)
                self.logger.info("Auto-adjusted batch size for GPU (cuda): %d", self.config.batch_size)
            self._embedding_dim = self.model.get_sentence_embedding_dimension()
            self.logger.info("Model loaded successfully on %s. Embedding dimension: %s", self.model.device,
                             self._embedding_dim)
            if self.config.use_gpu:
                log_gpu_status(self._gpu_capability, self.logger)
        except Exception as e:
            self.logger.error("Failed to load model %s: %s", self.config.model_name, e)
            raise

    def _determine_target_device(self) -> str:
        if self.config.device not in ("auto", "cpu"):
            return self.config.device
        return "cuda"

    # -------------------------------------------------------------- embeddings
    def generate_embeddings(self, chunks: List[Chunk]):
        if not self.model:
            self.load_model()
        if not chunks:
            return []
        embeddings = self._generate_embeddings_batch([c.text for c in chunks])
        if self.config.embedding_as_ndarray:
            for chunk, row in zip(chunks, embeddings):
                chunk.embedding = row
        else:
            for chunk, row in zip(chunks, embeddings):
                chunk.embedding = row.tolist()
        return embeddings

    def generate_single_embedding(self, text: str) -> np.ndarray:
        if not self.model:
            self.load_model()
        return self.model.encode(text, normalize_embeddings=self.config.normalize_embeddings,
                                 show_progress_bar=False)

    def _generate_embeddings_batch(self, texts: List[str]):
        start = time.time()
        clean = []
        for i, text in enumerate(texts):
            if text is None:
                self.logger.warning("Skipping chunk %d: text is None", i)
                clean.append("")
            elif not isinstance(text, str):
                self.logger.warning("Skipping chunk %d: text is not string (type: %s)", i, type(text))
                clean.append(str(text) if text else "")
            elif not text.strip():
                self.logger.warning("Skipping chunk %d: text is empty or whitespace only", i)
                clean.append("empty")
            else:
                clean.append(text)
        embeddings = self.model.encode(clean, batch_size=self.config.batch_size,
                                       normalize_embeddings=self.config.normalize_embeddings,
                                       show_progress_bar=self.config.show_progress, convert_to_numpy=True)
        dt = time.time() - start
        if self.config.show_progress:
            self.logger.info("Generated %d embeddings in %.2fs (%.1f chunks/s, avg length: %.0f chars)", len(texts), dt,
                             len(texts) / dt if dt > 0 else 0.0, float(np.mean([len(t or "") for t in texts])))
        return embeddings

    # ---------------------------------------------------- pure-numpy helpers
    def compute_similarity(self, embedding1: np.ndarray, embedding2: np.ndarray) -> float:
        return np.dot(embedding1, embedding2) / (np.linalg.norm(embedding1) * np.linalg.norm(embedding2))

    def compute_similarity_matrix(self, embeddings: List[np.ndarray]) -> np.ndarray:
        n = len(embeddings)
        m = np.zeros((n, n))
        for i in range(n):
            for j in range(i, n):
                m[i, j] = m[j, i] = self.compute_similarity(embeddings[i], embeddings[j])
        return m

    def find_similar_chunks(self, query_embedding: np.ndarray, chunk_embeddings: List[np.ndarray],
                            top_k: int = 5) -> List[tuple]:
        sims = [(i, self.compute_similarity(query_embedding, e)) for i, e in enumerate(chunk_embeddings)]
        sims.sort(key=lambda x: x[1], reverse=True)
        return sims[:top_k]

    def get_embedding_stats(self, chunks: List[Chunk]) -> EmbeddingStats:
        if not chunks:
            return EmbeddingStats()
        info = {}
        if self.model:
            info = {"model_name": self.config.model_name, "embedding_dimension": self._embedding_dim,
                    "max_seq_length": self.config.max_seq_length, "device": str(self.model.device)}
        return EmbeddingStats(total_chunks=len(chunks), total_tokens=sum(len(c.text.split()) for c in chunks),
                              average_chunk_length=float(np.mean([len(c.text) for c in chunks])), model_info=info)

    def save_embeddings(self, chunks: List[Chunk], file_path: str) -> None:
        data = [{"chunk_id": c.id, "embedding": c.embedding, "text": c.text, "metadata": c.metadata}
                for c in chunks if _has_embedding(c)]
        np.savez_compressed(file_path, embeddings=data)
        self.logger.info("Saved %d embeddings to %s", len(data), file_path)

    def load_embeddings(self, file_path: str) -> List[Chunk]:
        data = np.load(file_path, allow_pickle=True)
        chunks = [Chunk(id=it["chunk_id"], text=it["text"], metadata=it["metadata"], embedding=it["embedding"])
                  for it in data["embeddings"]]
        self.logger.info("Loaded %d embeddings from %s", len(chunks), file_path)
        return chunks

    def validate_embeddings(self, chunks: List[Chunk]) -> Dict[str, Any]:
        res: Dict[str, Any] = {"total_chunks": len(chunks), "chunks_with_embeddings": 0, "embedding_dimension": None,
                               "embedding_stats": {}, "issues": []}
        embs = []
        for c in chunks:
            if _has_embedding(c):
                res["chunks_with_embeddings"] += 1
                embs.append(np.array(c.embedding))
                if res["embedding_dimension"] is None:
                    res["embedding_dimension"] = len(c.embedding)
                elif res["embedding_dimension"] != len(c.embedding):
                    res["issues"].append(f"Inconsistent embedding dimension for chunk {c.id}")
            else:
                res["issues"].append(f"Missing embedding for chunk {c.id}")
        if embs:
            if len({len(e) for e in embs}) == 1:
                a = np.array(embs)
                norms = np.linalg.norm(a, axis=1)
                res["embedding_stats"] = {"mean": a.mean(0).tolist(), "std": a.std(0).tolist(),
                                          "min": a.min(0).tolist(), "max": a.max(0).tolist(),
                                          "norm_mean": norms.mean(), "norm_std": norms.std()}
            else:
                norms = [np.linalg.norm(e) for e in embs]
                res["embedding_stats"] = {"norm_mean": np.mean(norms), "norm_std": np.std(norms),
                                          "note": "Embeddings have different dimensions, limited stats computed"}
        return res

    def benchmark_model(self, test_texts: List[str], warmup_runs: int = 3) -> Dict[str, Any]:
        if not self.model:
            self.load_model()
        for _ in range(warmup_runs):
            self.model.encode(test_texts[: min(5, len(test_texts))], show_progress_bar=False)
        perf = {}
        for bs in (1, 4, 8, 16, 32):
            if bs > len(test_texts):
                continue
            t0 = time.time()
            for i in range(0, len(test_texts), bs):
                self.model.encode(test_texts[i:i + bs], show_progress_bar=False)
            dt = time.time() - t0
            perf[f"batch_size_{bs}"] = {"total_time": dt, "throughput": len(test_texts) / dt,
                                        "avg_time_per_text": dt / len(test_texts)}
        info = _native.device_info(0)
        return {"model_name": self.config.model_name, "device": str(self.model.device),
                "embedding_dimension": self._embedding_dim, "test_texts_count": len(test_texts), "performance": perf,
                "memory_info": {"allocated": info["hbm_total"] - info["hbm_free"], "reserved": 0, "max_allocated": 0}}

    # ------------------------------------------------------------- properties
    @property
    def embedding_dimension(self) -> Optional[int]:
        return self._embedding_dim

    @property
    def is_model_loaded(self) -> bool:
        return self.model is not None

    def get_model_info(self) -> Dict[str, Any]:
        if not self.model:
            return {}
        info = {"model_name": self.config.model_name, "embedding_dimension": self._embedding_dim,
                "max_seq_length": self.config.max_seq_length, "device": str(self.model.device),
                "batch_size": self.config.batch_size, "use_gpu": self.config.use_gpu,
                "gpu_available": bool(self._gpu_capability and self._gpu_capability.can_use_gpu)}
        cap = self._gpu_capability
        if cap and cap.can_use_gpu:
            info["gpu_info"] = {"gpu_count": cap.gpu_count, "gpu_names": cap.gpu_names,
                                "gpu_memory_total_gb": (cap.gpu_memory_total or 0) / 1024 ** 3,
                                "gpu_memory_free_gb": (cap.gpu_memory_free or 0) / 1024 ** 3,
                                "recommended_batch_size": cap.recommended_batch_size}
        return info

    @property
    def is_using_gpu(self) -> bool:
        return self.model is not None
