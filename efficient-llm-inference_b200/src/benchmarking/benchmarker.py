"""`src.benchmarking.benchmarker` of the reference, served by kvq_b200."""

from kvq_b200.benchmarking.benchmarker import KVCacheBenchmarker  # noqa: F401
