"""Drop-in for the model-local STFT pair of CosyVoice3's HiFT generator, mlx_audio/tts/models/cosyvoice3/
hifigan.py:382-499 (CosyVoice2's `_stft` / `_istft`, cosyvoice2/hifigan.py:452-470, wrap the same arithmetic).

Differences from the S3Gen pair that the reference really has, kept here: the forward transform pads with ZEROS
(399-400, `mx.pad` constant), and the inverse clips the magnitude to [0, 1e2] (447).  Envelope sum w^2 with a 1e-8
floor (493-494), n_fft//2 stripped from both ends (497-498)."""
from __future__ import annotations

from ....dsp import istft_polar
from ....dsp import stft as _stft


def stft(x, n_fft: int, hop_len: int, window):
    spec = _stft(x, n_fft=n_fft, hop_length=hop_len, win_length=n_fft, window=window, center=True,
                 pad_mode="constant")  # (B, frames, F)
    spec = spec.swapaxes(1, 2)
    return spec.real, spec.imag


def istft(magnitude, phase, n_fft: int, hop_len: int, window):
    return istft_polar(magnitude, phase, n_fft, hop_len, window, center=True, normalized=True, div_clamp=True,
                       div_eps=1e-8, trim_tail=True, mag_clip_max=1e2, mag_clip_min_zero=True)
