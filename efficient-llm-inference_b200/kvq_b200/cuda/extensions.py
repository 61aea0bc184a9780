"""The native extension, reference-compatible view.

The reference JIT-builds a pybind module `kvq_ext` from an inline CUDA string on first use
(/root/reference/src/cuda/extensions.py:12-147) with two functions:
    dequant_int8_to_fp16(q, scale) and dequant_int4_packed_to_fp16(packed, scale, orig_last_dim).
Here the extension is the prebuilt libkvq_sm100.so (C ABI, include/kvq.h).  `get_cuda_extension()`
returns an object with those two methods (same signatures and results) so code written against
`kvq_ext` keeps working, plus `.lib` for the raw ctypes handle.  Unlike the reference, a missing
library or a missing GPU is an error, not a silent `None` + torch fallback.
"""

from __future__ import annotations

import os
import subprocess
import threading
from typing import Optional

import torch

from .. import _lib as L

_ext: Optional["KvqExtension"] = None
_ext_lock = threading.Lock()


class KvqExtension:
    """`kvq_ext`-shaped facade over libkvq_sm100.so (reference extensions.py:116-119)."""

    def __init__(self):
        self.lib = L.lib()
        self.path = L.LIB_PATH

    def dequant_int8_to_fp16(self, q: torch.Tensor, scale: float) -> torch.Tensor:
        if not q.is_cuda:
            raise RuntimeError("q must be CUDA")          # CHECK_CUDA, extensions.py:33
        if not q.is_contiguous():
            raise RuntimeError("q must be contiguous")    # CHECK_CONTIGUOUS, extensions.py:34
        from ..quantization.ops import dequantize_int8_per_tensor
        s = torch.tensor(float(scale), dtype=torch.float32, device=q.device)
        return dequantize_int8_per_tensor(q, s, torch.float16)

    def dequant_int4_packed_to_fp16(self, packed: torch.Tensor, scale: float, orig_last_dim: int) -> torch.Tensor:
        """Returns the full 2*P columns with the padded tail zeroed, like the reference kernel
        (extensions.py:65-66); callers slice [..., :orig_last_dim] themselves (ops.py:119)."""
        if not packed.is_cuda:
            raise RuntimeError("packed must be CUDA")
        if not packed.is_contiguous():
            raise RuntimeError("packed must be contiguous")
        from ..quantization.ops import dequantize_int4_per_tensor_packed
        s = torch.tensor(float(scale), dtype=torch.float32, device=packed.device)
        full = 2 * packed.size(-1)
        out = dequantize_int4_per_tensor_packed(packed, s, full, torch.float16)
        if orig_last_dim < full:
            out[..., orig_last_dim:] = 0
        return out


def build_cuda_extension(verbose: bool = True) -> KvqExtension:
    """Compile libkvq_sm100.so in-tree (nvcc, sm_100a) if it is not there yet and return the facade.
    Cross-compiles without a GPU.  Reference: build_cuda_extension, extensions.py:12-135."""
    if not os.path.exists(L.LIB_PATH):
        csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(L.__file__))), "csrc")
        if verbose:
            print(f"building {L.LIB_PATH} from {csrc}")
        subprocess.run(["make", "-C", csrc, "-j8"], check=True, stdout=None if verbose else subprocess.DEVNULL)
    ext = KvqExtension()
    if verbose:
        print(f"CUDA extension loaded: {ext.path} (kvq ABI v{ext.lib.kvq_version()})")
    return ext


def get_cuda_extension() -> KvqExtension:
    """Process-wide handle, created on first use (thread-safe)."""
    global _ext
    if _ext is None:
        with _ext_lock:
            if _ext is None:
                _ext = build_cuda_extension(verbose=False)
    return _ext
