"""Capability probe + batch heuristic of the reference (src/gpu_utils.py:169-267) for the
B200 path.  The probe asks libcss_b200.so (css_device_count / css_device_info), not
torch / faiss."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

from . import _native


@dataclass
class GPUCapability:
    """Same fields the reference's callers read (src/gpu_utils.py GPUCapability)."""

    torch_cuda_available: bool = False
    faiss_gpu_available: bool = False
    gpu_count: int = 0
    gpu_memory_total: Optional[int] = None
    gpu_memory_free: Optional[int] = None
    gpu_names: List[str] = field(default_factory=list)
    can_use_gpu: bool = False
    recommended_batch_size: int = 32
    status_message: str = ""


def calculate_optimal_batch_size(available_memory_gb: float, embedding_dim: int = 768, backend: str = "cuda") -> int:
    """src/gpu_utils.py:169-192: (free_GB - 1) / (dim * 4 * 4 B), clamped to [8, 256] (64 on mps)."""
    working = available_memory_gb - 1.0
    if working <= 0:
        return 8
    per_item = (embedding_dim * 4 * 4) / (1024 ** 3)
    bs = int(working / per_item)
    return max(8, min(bs, 64 if backend == "mps" else 256))


def assess_gpu_capability(target_chunks: int = 10000, embedding_dim: int = 768) -> GPUCapability:
    cap = GPUCapability()
    try:
        n = _native.device_count()
        info = _native.device_info(0)
    except _native.NativeError as e:
        cap.status_message = f"❌ GPU unavailable: {e}"
        return cap
    cap.torch_cuda_available = True
    cap.faiss_gpu_available = True          # the flat index is device-resident in this build
    cap.gpu_count = n
    cap.gpu_memory_total = int(info["hbm_total"])
    cap.gpu_memory_free = int(info["hbm_free"])
    cap.gpu_names = ["NVIDIA B200 (sm_%d%d)" % info["cc"]] * n
    cap.recommended_batch_size = calculate_optimal_batch_size(cap.gpu_memory_free / 1024 ** 3, embedding_dim)
    cap.can_use_gpu = True
    cap.status_message = f"✅ sm_100 GPU ready (Free: {cap.gpu_memory_free / 1024 ** 3:.1f}GB)"
    return cap


def log_gpu_status(capability: GPUCapability, logger) -> None:
    logger.info("GPU status: %s", capability.status_message)
    for i, name in enumerate(capability.gpu_names):
        logger.info("  GPU %d: %s", i, name)
