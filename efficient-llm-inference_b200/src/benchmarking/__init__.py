"""`src.benchmarking` of the reference, served by kvq_b200 (KVCacheBenchmarker only; the task harnesses are out of scope)."""

from kvq_b200.benchmarking import KVCacheBenchmarker

__all__ = ["KVCacheBenchmarker"]


def __getattr__(name):
    if name in ("MMLUBenchmark", "SummarizationBenchmark"):
        raise ImportError(f"src.benchmarking.{name} is outside the KV-cache hot path rebuilt here (SURVEY.md section 8)")
    raise AttributeError(name)
