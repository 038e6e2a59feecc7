"""Drop-in for HiFTGenerator._stft / ._istft of mlx_audio/tts/models/chatterbox_turbo/models/s3gen/hifigan.py:418-537.

Unlike the S3Gen pair, the forward transform here does NOT centre-pad: frames = (T - n_fft) // hop + 1 over the raw
signal (438), a signal shorter than n_fft is right-padded with zeros to one frame (441-446), window applied as given.
The inverse is the S3Gen one: magnitude clipped to <= 1e2 (487), sum w^2 envelope floored at 1e-8 (527-528), n_fft // 2
samples stripped in front and (T - 1) * hop kept (532-535)."""
from __future__ import annotations

from ......_arrays import _is_torch
from ......codec.models.s3gen.hifigan import hann_window_periodic, istft  # noqa: F401
from ......dsp import stft as _stft


def stft(x, n_fft: int, hop_len: int, window):
    """x: (B, T) -> (real, imag), each (B, n_fft // 2 + 1, frames)"""
    if x.shape[1] < n_fft:
        if _is_torch(x):
            import torch

            x = torch.nn.functional.pad(x, (0, n_fft - x.shape[1]))
        else:
            import numpy as np

            x = np.pad(np.asarray(x), ((0, 0), (0, n_fft - x.shape[1])))
    spec = _stft(x, n_fft=n_fft, hop_length=hop_len, win_length=n_fft, window=window, center=False).swapaxes(1, 2)
    return spec.real, spec.imag
