"""Drop-in for the two DSP calls of MossFormer2-SE's chunk loop, mlx_audio/sts/models/mossformer2_se/model.py:396-428:
`stft(audio_segment, fft_len, win_inc, win_len, window, center=False)` handed on as (freq, time) real / imaginary planes, and
`ISTFTCache.istft(real[None], imag[None], fft_len, win_inc, win_len, window, center=False, audio_length=chunk_length)`.
(The Kaldi fbank + deltas in front of them are dsp.compute_fbank_kaldi / compute_deltas_kaldi; the mask network between them
belongs to the model.)"""
from __future__ import annotations

import numpy as np

from ...._arrays import _is_torch
from ....dsp import ISTFTCache, hamming, stft

_cache = ISTFTCache()


def chunk_stft(audio_segment, fft_len: int = 1920, win_inc: int = 384, win_len: int = 1920, window=None):
    """-> (real, imag), each (fft_len // 2 + 1, T) (model.py:396-406)"""
    if window is None:
        window = hamming(win_len, periodic=False)  # model.py:145
    s = stft(audio_segment, fft_len, win_inc, win_len, window, center=False)
    if _is_torch(s):
        return s.real.T.contiguous(), s.imag.T.contiguous()
    s = np.asarray(s)
    return np.ascontiguousarray(s.real.T), np.ascontiguousarray(s.imag.T)


def chunk_istft(spectrum_real, spectrum_imag, fft_len: int = 1920, win_inc: int = 384, win_len: int = 1920, window=None,
                chunk_length: int = None, cache: ISTFTCache = None):
    """(F, T) masked planes -> (samples,) (model.py:415-428)"""
    if window is None:
        window = hamming(win_len, periodic=False)
    c = _cache if cache is None else cache
    return c.istft(spectrum_real[None], spectrum_imag[None], fft_len, win_inc, win_len, window, center=False,
                   audio_length=chunk_length)[0]
