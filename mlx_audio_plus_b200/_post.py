"""Plumbing for the steps right after the path (csrc/post.cu): device buffers in, device buffers out."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from ._arrays import _is_torch


def _to_cuda_f32(x):
    import torch

    if L.lib.b2a_device_count() < 1:
        raise L.B2AError("b200audio: no CUDA device — there is no CPU fallback")
    t = x if _is_torch(x) else torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.float32))
    t = t.to(torch.float32).contiguous()
    return t if t.is_cuda else t.cuda()


def _back(x, out):
    if _is_torch(x):
        return out if x.is_cuda else out.cpu()
    import torch

    if out.dtype == torch.bfloat16:  # NumPy has no bfloat16
        raise TypeError("bfloat16 output needs a torch input")
    return out.cpu().numpy()


def _stream():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_DTYPES = {"float32": L.DTYPE_F32, "float16": L.DTYPE_F16, "bfloat16": L.DTYPE_BF16}


def rows_pad_cast(x, row_begin: int, rows_valid: int, rows_out: int, dtype="float16"):
    """x: (T, M) or (B, T, M) float32 -> (rows_out, M) / (B, rows_out, M) in `dtype`: rows [row_begin, row_begin + rows_valid)
    followed by zeros (whisper/whisper.py:990-996)."""
    import torch

    name = str(dtype).split(".")[-1]
    if name not in _DTYPES:
        raise ValueError(f"unsupported dtype {dtype}")
    t = _to_cuda_f32(x)
    one = t.ndim == 2
    if one:
        t = t[None]
    B, T, M = t.shape
    row_begin = max(int(row_begin), 0)
    rows_valid = max(0, min(int(rows_valid), T - row_begin, int(rows_out)))
    out = torch.empty((B, rows_out, M), dtype=getattr(torch, name), device=t.device)
    with torch.cuda.device(t.device):
        L.check(L.lib.b2a_rows_pad_cast(t.data_ptr(), T * M, row_begin, rows_valid, M, out.data_ptr(), int(rows_out), _DTYPES[name], B,
                                        _stream()))
    return _back(x, out[0] if one else out)


def transpose_pad(features, rows_out: int = None):
    """(B, T, M) -> (B, M, rows_out) float32 with zeros behind column T (Sortformer pad_to, sortformer.py:112-118), one pass"""
    import torch

    t = _to_cuda_f32(features)
    B, T, M = t.shape
    rows_out = T if rows_out is None else int(rows_out)
    out = torch.empty((B, M, rows_out), dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        L.check(L.lib.b2a_transpose_pad(t.data_ptr(), out.data_ptr(), T, M, rows_out, B, _stream()))
    return _back(features, out)


def cmvn_utterance(features, eps: float = 1e-6):
    """(x - mean) / (std + eps) over the frames of every feature column (funasr/audio.py:160-164); (T, M) or (B, T, M)"""
    import torch

    t = _to_cuda_f32(features)
    one = t.ndim == 2
    if one:
        t = t[None]
    B, T, M = t.shape
    out = torch.empty_like(t)
    ws = torch.empty((B, M, 2), dtype=torch.float64, device=t.device)
    with torch.cuda.device(t.device):
        L.check(L.lib.b2a_cmvn_utterance(t.data_ptr(), out.data_ptr(), 0, T, M, float(eps), ws.data_ptr(), B, _stream()))
    return _back(features, out[0] if one else out)


def lfr(features, lfr_m: int, lfr_n: int, cmvn_shift=None, cmvn_scale=None):
    """features: (T, M) or (B, T, M) float32 -> (ceil(T / lfr_n), lfr_m * M) [batched alike] (funasr/audio.py:84-139)"""
    import torch

    t = _to_cuda_f32(features)
    one = t.ndim == 2
    if one:
        t = t[None]
    B, T, M = t.shape
    if T < 1:
        raise ValueError("apply_lfr needs at least one frame")
    t_lfr = -(-T // int(lfr_n))
    out = torch.empty((B, t_lfr, int(lfr_m) * M), dtype=torch.float32, device=t.device)
    sh = sc = None
    if cmvn_shift is not None:
        sh = _to_cuda_f32(cmvn_shift).reshape(-1)
        sc = _to_cuda_f32(cmvn_scale).reshape(-1)
        if sh.numel() != lfr_m * M or sc.numel() != lfr_m * M:
            raise ValueError(f"CMVN vectors must have {lfr_m * M} entries")
    with torch.cuda.device(t.device):
        L.check(L.lib.b2a_lfr(t.data_ptr(), T * M, T, M, int(lfr_m), int(lfr_n), sh.data_ptr() if sh is not None else None,
                              sc.data_ptr() if sc is not None else None, out.data_ptr(), 0, B, _stream()))
    return _back(features, out[0] if one else out)


def rows_normalize(x, valid=None, *, den_kind: int = 0, eps: float = 1e-7, pad_value: float = 0.0):
    """Row-wise zero-mean / unit-variance of waveforms, one launch: x (L,) or (B, L) float32 -> same shape.  `valid`: per-row
    sample counts (the rest of a row becomes `pad_value`).  den_kind 0 = sqrt(var + eps) (transformers' zero_mean_unit_var_norm),
    1 = max(std, eps) (smart_turn.py:196-199)."""
    import torch

    t = _to_cuda_f32(x)
    one = t.ndim == 1
    if one:
        t = t[None]
    B, n = t.shape
    out = torch.empty_like(t)
    v = None
    if valid is not None:
        v = torch.as_tensor(valid, dtype=torch.int64, device=t.device).contiguous()
    if n > 0:
        with torch.cuda.device(t.device):
            L.check(L.lib.b2a_rows_normalize(t.data_ptr(), out.data_ptr(), B, n, v.data_ptr() if v is not None else None, int(den_kind),
                                             float(eps), float(pad_value), _stream()))
    return _back(x, out[0] if one else out)


def unwrap(p, discont: float, period: float, axis: int = -1):
    """Phase unwrap along `axis` (kokoro/istftnet.py:418-452), one launch over rows of the last axis."""
    import torch

    t = _to_cuda_f32(p)
    moved = axis not in (-1, t.ndim - 1)
    if moved:
        t = t.movedim(axis, -1).contiguous()
    shape = t.shape
    n = int(shape[-1]) if t.ndim else 0
    out = torch.empty_like(t)
    rows = t.numel() // n if n else 0
    if rows > 0:
        with torch.cuda.device(t.device):
            L.check(L.lib.b2a_unwrap(t.data_ptr(), out.data_ptr(), rows, n, float(discont), float(period), _stream()))
    if moved:
        out = out.movedim(-1, axis)
    return _back(p, out)
