"""Chunk record exchanged with the hot path (mirror of the reference's
src/chunker.py:16-23 dataclass: id, text, metadata, embedding)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional


@dataclass
class Chunk:
    id: str
    text: str
    metadata: Dict[str, Any] = field(default_factory=dict)
    embedding: Optional[List[float]] = None
