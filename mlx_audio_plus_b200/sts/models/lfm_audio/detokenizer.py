"""Drop-in for the waveform step of mlx_audio/sts/models/lfm_audio/detokenizer.py:462-507 (`LFM2AudioDetokenizer._istft`):
exp'd magnitude and phase planes (B, T, F) -> per-item istft(center=False, normalized=True) with the checkpoint's window,
then the "same" trim of (n_fft - hop) / 2 samples at both ends -> (B, T * hop).  One fused polar-input launch for the
batch (cos / sin, inverse FFT, window, overlap-add, window**2 envelope division) instead of the reference's Python loop."""
from __future__ import annotations

from ...._arrays import _is_torch
from ....dsp import istft_polar


def istft_same(mag, phase, window, n_fft: int = 1280, hop_length: int = 320):
    if _is_torch(mag):
        m, p = mag.transpose(1, 2).contiguous(), phase.transpose(1, 2).contiguous()
    else:
        import numpy as np

        m = np.ascontiguousarray(np.swapaxes(np.asarray(mag, dtype=np.float32), 1, 2))
        p = np.ascontiguousarray(np.swapaxes(np.asarray(phase, dtype=np.float32), 1, 2))
    y = istft_polar(m, p, n_fft, hop_length, window, center=False, normalized=True)
    pad = (n_fft - hop_length) // 2
    return y[:, pad:-pad] if pad > 0 else y
