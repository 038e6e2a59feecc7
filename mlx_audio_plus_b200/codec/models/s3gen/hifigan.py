"""Drop-in for the model-local STFT pair of the HiFT vocoder, mlx_audio/codec/models/s3gen/hifigan.py:408-549
(the same functions serve Chatterbox / Chatterbox-Turbo S3Gen; n_fft=16, hop=4, periodic Hann in HiFTGenerator).

    stft(x (B, T), n_fft, hop_length, window)              -> (real, imag), each (B, n_fft//2+1, frames)
    istft(magnitude, phase, n_fft, hop_length, window)     -> (B, (frames-1)*hop_length)

Reference behaviour kept: reflect padding of n_fft//2 without repeating the edge sample (421-427), frames =
(T_padded - n_fft)//hop + 1 (430), window applied as given (445-446); the inverse clips the magnitude to <= 1e2
(481), ignores Im(DC) / Im(Nyquist) (real part of the full ifft, 491-502), divides by max(sum w^2, 1e-8) (517-521)
and strips n_fft//2 samples from both ends (543-545).  Both directions process the whole batch in one launch; the
inverse forms clip(mag) * (cos p, sin p) inside the iSTFT kernel (dsp.istft_polar)."""
from __future__ import annotations

from ....dsp import hanning, istft_polar
from ....dsp import stft as _stft


def hann_window_periodic(size: int):
    """hifigan.py:13-19: 0.5 (1 - cos(2 pi n / size)) evaluated in float64 per tap — dsp.hanning(size, periodic=True)."""
    return hanning(size, True)


def stft(x, n_fft: int, hop_length: int, window):
    spec = _stft(x, n_fft=n_fft, hop_length=hop_length, win_length=n_fft, window=window, center=True,
                 pad_mode="reflect")  # (B, frames, F)
    spec = spec.swapaxes(1, 2)
    return spec.real, spec.imag


def istft(magnitude, phase, n_fft: int, hop_length: int, window):
    return istft_polar(magnitude, phase, n_fft, hop_length, window, center=True, normalized=True, div_clamp=True,
                       div_eps=1e-8, trim_tail=True, mag_clip_max=1e2)
