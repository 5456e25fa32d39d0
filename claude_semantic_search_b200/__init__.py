"""B200-native hot path for claude-semantic-search.

Public surface mirrors the reference package's re-exports for this path
(src/__init__.py:11-31): EmbeddingGenerator / EmbeddingConfig / EmbeddingStats,
HybridStorage / StorageConfig / SearchConfig / SearchResult, Chunk.
Everything computes through libcss_b200.so (hand-written sm_100a CUDA behind
the C ABI in include/css_b200.h); there is no CPU fallback.
"""
from .chunk import Chunk
from .hybrid_storage import HybridStorage, SearchConfig, SearchResult, StorageConfig

__version__ = "0.1.0"

__all__ = [
    "Chunk",
    "HybridStorage",
    "StorageConfig",
    "SearchConfig",
    "SearchResult",
]

from .embedding_generator import EmbeddingConfig, EmbeddingGenerator, EmbeddingStats  # noqa: E402
from .encoder import MPNetEncoder  # noqa: E402

__all__ += ["EmbeddingGenerator", "EmbeddingConfig", "EmbeddingStats", "MPNetEncoder"]
