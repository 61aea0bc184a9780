"""Helpers to read tests/golden/*.npz (written by tests/golden/make_golden.py from the reference)."""

import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DT = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}
TAG = {v: k for k, v in DT.items()}


def load(name: str):
    return np.load(os.path.join(GOLDEN_DIR, name))


def tensor(npz, key_prefix: str):
    """Find `key_prefix_<tag>` (or the bare key) and return it as a torch tensor of the right dtype."""
    for tag, dt in DT.items():
        k = f"{key_prefix}_{tag}"
        if k in npz.files:
            a = npz[k]
            if tag == "f32":
                return torch.from_numpy(a.copy())
            return torch.from_numpy(a.copy()).view(dt)
    return torch.from_numpy(np.asarray(npz[key_prefix]).copy())


def typed(npz, key: str, tag: str):
    a = torch.from_numpy(np.asarray(npz[key]).copy())
    return a if tag == "f32" else a.view(DT[tag])


def bits_equal(a: torch.Tensor, b: torch.Tensor) -> bool:
    """Bit-for-bit equality (distinguishes -0.0 from 0.0, treats equal NaN payloads as equal)."""
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    a, b = a.contiguous().cpu(), b.contiguous().cpu()
    if a.dtype in (torch.float16, torch.bfloat16):
        return torch.equal(a.view(torch.int16), b.view(torch.int16))
    if a.dtype == torch.float32:
        return torch.equal(a.view(torch.int32), b.view(torch.int32))
    return torch.equal(a, b)
