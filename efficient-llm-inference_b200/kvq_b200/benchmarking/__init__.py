"""Benchmark harness (reference: src/benchmarking/__init__.py:7 -- KVCacheBenchmarker only; the MMLU and
summarisation task harnesses are out of the hot path)."""

from .benchmarker import KVCacheBenchmarker

__all__ = ["KVCacheBenchmarker"]
