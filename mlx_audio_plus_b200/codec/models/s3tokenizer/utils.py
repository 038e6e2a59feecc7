"""Drop-in for the feature functions of mlx_audio/codec/models/s3tokenizer/utils.py:13-135."""
from __future__ import annotations

from ...._arrays import emit
from ...._wrap import as_batch, run_frontend
from .... import _lib as L
from ....dsp import hanning, mel_filters


def log_mel_spectrogram(audio, sample_rate: int = 16000, n_mels: int = 128, n_fft: int = 400,
                        hop_length: int = 160, padding: int = 0):
    """(L,) -> (n_mels, T'): periodic Hann via hanning(n_fft+1)[:-1], no frame drop (utils.py:13-65)."""
    ing, was_1d = as_batch(audio)
    fb = mel_filters(sample_rate=sample_rate, n_fft=n_fft, n_mels=n_mels, norm="slaney", mel_scale="slaney")
    out = run_frontend(
        ing, hanning(n_fft + 1)[:-1], fb, length=ing.data.shape[1] + max(int(padding), 0),
        n_fft=n_fft, hop=hop_length, center=True, pad_mode="reflect", spec_kind=L.SPEC_POWER,
        log_kind=L.LOG_LOG10, guard_kind=L.GUARD_MAX, guard_eps=1e-10, clamp_kind=L.CLAMP_CLIP_MAX,
        clamp_value=8.0, affine_add=4.0, affine_div=4.0, out_layout=L.LAYOUT_MT)
    return emit(ing, out[0] if was_1d else out)


def log_mel_spectrogram_compat(audio, n_mels: int = 128, padding: int = 0):
    """(L,) or (B, L) -> (n_mels, T') / (B, n_mels, T'): symmetric "hann", last frame dropped, ONE max over
    the whole batch (utils.py:68-135)."""
    ing, was_1d = as_batch(audio)
    fb = mel_filters(sample_rate=16000, n_fft=400, n_mels=n_mels, norm="slaney", mel_scale="slaney")
    out = run_frontend(
        ing, hanning(400), fb, length=ing.data.shape[1] + max(int(padding), 0),
        n_fft=400, hop=160, center=True, pad_mode="reflect", drop_last=True, spec_kind=L.SPEC_POWER,
        log_kind=L.LOG_LOG10, guard_kind=L.GUARD_MAX, guard_eps=1e-10, clamp_kind=L.CLAMP_BATCH_MAX,
        clamp_value=8.0, affine_add=4.0, affine_div=4.0, out_layout=L.LAYOUT_MT)
    return emit(ing, out[0] if was_1d else out)
