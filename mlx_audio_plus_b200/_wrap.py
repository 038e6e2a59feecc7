"""Shared plumbing for the per-model front-end wrappers: one fused-kernel plan per parameter set."""
from __future__ import annotations

import numpy as np

from . import _lib as L
from ._arrays import Ingested, emit, ingest
from .frontend import FrontendPlan, _device_index, cached_plan


def as_batch(audio):
    """-> (Ingested with (B, L) data, was_1d)"""
    ing = ingest(audio, "float32")
    was_1d = ing.data.ndim == 1
    if was_1d:
        ing.data = ing.data.reshape(1, -1)
    elif ing.data.ndim != 2:
        raise ValueError("expected a 1-D waveform or a (B, L) batch")
    return ing, was_1d


def run_frontend(ing: Ingested, window, filterbank, *, length=None, pad_value=0.0, **plan_kw):
    plan = cached_plan(FrontendPlan, _device_index(ing), np.asarray(window, dtype=np.float32),
                       None if filterbank is None else np.asarray(filterbank, dtype=np.float32), **plan_kw)
    return plan.run(ing, length=length, pad_value=pad_value)


def reflect_pad_rows(ing: Ingested, pad: int) -> Ingested:
    """Caller-side reflect padding `x[1:p+1][::-1] | x | x[-(p+1):-1][::-1]` per row (qwen3_tts.py:71-73)."""
    x = ing.data
    if ing.on_device:
        import torch

        y = torch.cat([x[:, 1 : pad + 1].flip(1), x, x[:, -(pad + 1) : -1].flip(1)], dim=1).contiguous()
    else:
        y = np.ascontiguousarray(np.concatenate([x[:, 1 : pad + 1][:, ::-1], x, x[:, -(pad + 1) : -1][:, ::-1]], axis=1))
    return Ingested(ing.family, ing.on_device, y, ing.orig_dtype, ing.device)
