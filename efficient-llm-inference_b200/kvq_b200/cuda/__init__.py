"""Native-extension access (reference: src/cuda/__init__.py:5-8)."""

from .extensions import build_cuda_extension, get_cuda_extension

__all__ = ["build_cuda_extension", "get_cuda_extension"]
