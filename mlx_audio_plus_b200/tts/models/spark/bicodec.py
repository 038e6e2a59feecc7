"""Drop-in for mel_spectrogram of mlx_audio/tts/models/spark/bicodec.py:20-49: periodic Hann of win_length taps
(hanning(win_length + 1)[:-1]) right-padded to n_fft by dsp.stft, hop 320, magnitude, Slaney/slaney filterbank from f_min,
no logarithm, (1, T, n_mels)."""
from __future__ import annotations

from typing import Optional

from ...._arrays import emit
from ...._wrap import as_batch, run_frontend
from .... import _lib as L
from ....dsp import hanning, mel_filters


def mel_spectrogram(audio, sample_rate: int = 16_000, n_mels: int = 128, n_fft: int = 1024, f_min: int = 10,
                    f_max: Optional[int] = None, hop_length: int = 320, win_length: int = 640, padding: int = 0):
    ing, _ = as_batch(audio)
    fb = mel_filters(sample_rate=sample_rate, n_fft=n_fft, n_mels=n_mels, f_min=f_min, f_max=f_max, norm="slaney",
                     mel_scale="slaney")
    out = run_frontend(
        ing, hanning(win_length + 1)[:-1], fb, length=ing.data.shape[1] + max(int(padding), 0), n_fft=n_fft,
        hop=hop_length, center=True, pad_mode="reflect", spec_kind=L.SPEC_MAGNITUDE)
    return emit(ing, out)  # (1, T, M) / (B, T, M)
