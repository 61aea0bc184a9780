"""Configuration records with the reference's fields and defaults
(/root/reference/src/core/config.py:10-84): eps 1e-8, window 256, block 64, chunk 64, keep_last 256."""

from __future__ import annotations

import random
from dataclasses import dataclass, field
from typing import List, Literal

import torch


def _default_device() -> str:
    return "cuda" if torch.cuda.is_available() else "cpu"


def _default_dtype() -> torch.dtype:
    return torch.float16 if torch.cuda.is_available() else torch.float32


@dataclass
class Config:
    """Run-level settings; constructing one seeds python, torch and CUDA RNGs (reference :32-37)."""

    model_name: str = "gpt2"
    device: str = field(default_factory=_default_device)
    dtype: torch.dtype = field(default_factory=_default_dtype)
    seed: int = 42
    max_new_tokens: int = 64
    batch_size: int = 1

    def __post_init__(self):
        random.seed(self.seed)
        torch.manual_seed(self.seed)
        if self.device == "cuda" and torch.cuda.is_available():
            torch.cuda.manual_seed_all(self.seed)


@dataclass
class QuantizationConfig:
    mode: Literal["int8", "int4", "mixed"] = "int8"
    eps: float = 1e-8


@dataclass
class CacheConfig:
    window_size: int = 256
    block_size: int = 64
    chunk_size: int = 64
    keep_last: int = 256


@dataclass
class BenchmarkConfig:
    methods: List[str] = field(default_factory=lambda: ["no_cache", "full_cache", "sliding_window"])
    window_sizes: List[int] = field(default_factory=lambda: [64, 128, 256, 512])
    block_sizes: List[int] = field(default_factory=lambda: [32, 64, 128])
    chunk_sizes: List[int] = field(default_factory=lambda: [32, 64, 128])
