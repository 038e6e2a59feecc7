"""Drop-in for `kaldi_fbank` of mlx_audio/codec/models/s3gen/xvector.py:38-150 (the CAMPPlus speaker-encoder front-end):
snip-edges framing (400 / 160 at 16 kHz), per-frame DC removal and pre-emphasis 0.97, float32 Povey window, zero-extension to
512, power spectrum, dsp.mel_filters (HTK scale from 20 Hz, no normalisation — not the Kaldi banks of dsp.py),
ln(max(., float32 eps)) — one launch of the per-frame-preprocessing kernel that serves dsp.compute_fbank_kaldi."""
from __future__ import annotations

import numpy as np

from ...._arrays import emit, ingest
from .... import _lib as L
from ....dsp import _kaldi_window, _next_power_of_2, mel_filters
from ....frontend import FrontendPlan, _device_index, cached_plan

_FLT_EPSILON = 1.1920929e-07


def kaldi_fbank(audio, sample_rate: int = 16000, num_mel_bins: int = 80, frame_length: float = 25.0, frame_shift: float = 10.0):
    """(T,) -> (frames, num_mel_bins)"""
    ing = ingest(audio, "float32")
    if ing.data.ndim > 1:
        ing.data = ing.data.squeeze()  # xvector.py:72-73
    win_length = int(sample_rate * frame_length / 1000)
    hop_length = int(sample_rate * frame_shift / 1000)
    n_fft = _next_power_of_2(win_length)
    n = int(ing.data.shape[0])
    frames = max((n - win_length) // hop_length + 1, 1)  # :76-79; a short signal is zero-extended to ONE frame
    fb = mel_filters(sample_rate=sample_rate, n_fft=n_fft, n_mels=num_mel_bins, f_min=20.0, f_max=sample_rate / 2, norm=None,
                     mel_scale="htk")
    plan = cached_plan(FrontendPlan, _device_index(ing), _kaldi_window("povey", win_length), np.asarray(fb, dtype=np.float32),
                       n_fft=n_fft, hop=hop_length, center=False, spec_kind=L.SPEC_POWER, log_kind=L.LOG_LN,
                       guard_kind=L.GUARD_MAX, guard_eps=_FLT_EPSILON, frame_len=win_length, frame_dc=True, frame_preemph=0.97)
    ing.data = ing.data.reshape(1, -1)
    # virtual right padding: the zero extension of the last frame to n_fft (and of a short signal to one window)
    length = max(n, win_length) + (n_fft - win_length)
    out = plan.run(ing, length=length, frame_count=frames)
    return emit(ing, out[0])
