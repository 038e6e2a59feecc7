"""Drop-in for mlx_audio/codec/models/s3gen/mel.py:25-100: caller-side reflect padding of (n_fft - hop) / 2 samples, then
stft(center=False) with the symmetric "hann" window, magnitude, Slaney/slaney filterbank up to fmax, ln(max(., 1e-5)),
(B, num_mels, T') — the whole batch in one launch instead of the reference's per-item loop."""
from __future__ import annotations

from ...._arrays import emit
from ...._wrap import as_batch, reflect_pad_rows, run_frontend
from .... import _lib as L
from ....dsp import hanning, mel_filters


def mel_spectrogram(y, n_fft: int = 1920, num_mels: int = 80, sampling_rate: int = 24000, hop_size: int = 480,
                    win_size: int = 1920, fmin: int = 0, fmax: int = 8000, center: bool = False):
    ing, _ = as_batch(y)
    fb = mel_filters(sample_rate=sampling_rate, n_fft=n_fft, n_mels=num_mels, f_min=fmin, f_max=fmax, norm="slaney",
                     mel_scale="slaney")
    pad = (n_fft - hop_size) // 2
    if pad:
        ing = reflect_pad_rows(ing, pad)
    out = run_frontend(  # the reference passes center=False whatever its own `center` argument says (mel.py:69-76)
        ing, hanning(win_size), fb, n_fft=n_fft, hop=hop_size, center=False, spec_kind=L.SPEC_MAGNITUDE,
        log_kind=L.LOG_LN, guard_kind=L.GUARD_MAX, guard_eps=1e-5, out_layout=L.LAYOUT_MT)
    return emit(ing, out)  # (B, M, T')
