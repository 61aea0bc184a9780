"""Memory accounting helpers (reference: src/core/utils.py:10-71); `mb` divides by 1024**2."""

from __future__ import annotations

import os
from typing import Optional

import torch


def get_cpu_mem_mb() -> float:
    """Resident set size of this process in MB."""
    import psutil

    return psutil.Process(os.getpid()).memory_info().rss / (1024 ** 2)


def _cuda(device: str) -> bool:
    return str(device).startswith("cuda") and torch.cuda.is_available()


def reset_gpu_peak(device: str = "cuda") -> None:
    if _cuda(device):
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()


def get_gpu_peak_mb(device: str = "cuda") -> Optional[float]:
    if _cuda(device):
        return torch.cuda.max_memory_allocated() / (1024 ** 2)
    return None


def tensor_bytes(tensor: torch.Tensor) -> int:
    return tensor.numel() * tensor.element_size()


def mb(num_bytes: int) -> float:
    return num_bytes / (1024 ** 2)


def kv_bytes_fp(k: torch.Tensor, v: torch.Tensor) -> int:
    return tensor_bytes(k) + tensor_bytes(v)
