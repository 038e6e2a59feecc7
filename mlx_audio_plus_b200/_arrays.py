"""Array interop: numpy / torch / mlx (and anything speaking DLPack or __cuda_array_interface__).

The library computes on device buffers.  A torch CUDA tensor is used in place (zero copy, result is a
torch CUDA tensor on the same device).  Host arrays (numpy, torch-cpu, mlx via the buffer protocol) go
through the C-ABI host entry points, which pipeline H2D / compute / D2H; the result comes back in the
caller's own family so the reference's call sites (`.abs().square()`, `[:-1, :]`, `@ filters.T`,
`.swapaxes`) keep working unchanged.
"""
from __future__ import annotations

import numpy as np


class DspArray(np.ndarray):
    """ndarray with the handful of mx.array *methods* the reference's call sites chain on results
    (whisper/audio.py:77 `.abs().square()`, :82 `.log10()`; vocos/mel.py:32 `.log()`)."""

    def abs(self):
        return np.abs(self)

    def square(self):
        return np.square(self)

    def log(self):
        return np.log(self)

    def log10(self):
        return np.log10(self)

    def exp(self):
        return np.exp(self)

    def sqrt(self):
        return np.sqrt(self)

    def moveaxis(self, source, destination):
        return np.moveaxis(self, source, destination)


def _torch():
    import torch

    return torch


def _is_torch(x):
    return type(x).__module__.split(".")[0] == "torch" and hasattr(x, "data_ptr")


def _is_mlx(x):
    return type(x).__module__.split(".")[0] == "mlx"


class Ingested:
    """family: 'numpy' | 'torch' | 'mlx';  on_device: torch CUDA tensor (zero copy) else host ndarray."""

    __slots__ = ("family", "on_device", "data", "orig_dtype", "device")

    def __init__(self, family, on_device, data, orig_dtype=None, device=None):
        self.family, self.on_device, self.data = family, on_device, data
        self.orig_dtype, self.device = orig_dtype, device


def ingest(x, dtype="float32") -> Ingested:
    """dtype: 'float32' or 'complex64' (what the kernels read)."""
    if _is_torch(x):
        torch = _torch()
        tdt = torch.float32 if dtype == "float32" else torch.complex64
        if x.is_cuda:
            return Ingested("torch", True, x.detach().to(tdt).contiguous(), x.dtype, x.device)
        return Ingested("torch", False, np.ascontiguousarray(x.detach().to(tdt).numpy()), x.dtype)
    if _is_mlx(x):
        # mlx arrays export the buffer protocol (CPU-visible unified memory)
        return Ingested("mlx", False, np.ascontiguousarray(np.asarray(x), dtype=dtype), getattr(x, "dtype", None))
    if hasattr(x, "__cuda_array_interface__") or (hasattr(x, "__dlpack__") and not isinstance(x, np.ndarray)):
        torch = _torch()
        t = torch.as_tensor(x, device="cuda") if hasattr(x, "__cuda_array_interface__") else torch.from_dlpack(x)
        r = ingest(t, dtype)
        return r
    return Ingested("numpy", False, np.ascontiguousarray(np.asarray(x), dtype=dtype))


def emit(ing: Ingested, result):
    """result: torch CUDA tensor (device path) or ndarray (host path) -> caller's family."""
    if ing.family == "torch":
        if ing.on_device:
            return result
        return _torch().from_numpy(np.ascontiguousarray(result))
    if ing.family == "mlx":
        import mlx.core as mx  # only reachable when the caller handed us an mlx array

        return mx.array(result)
    return np.asarray(result).view(DspArray)


def host_window(window) -> np.ndarray:
    """An array window (numpy / torch / mlx) as a float32 host vector."""
    if _is_torch(window):
        window = window.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(window), dtype=np.float32).reshape(-1)
