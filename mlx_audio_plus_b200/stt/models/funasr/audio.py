"""Drop-in for mlx_audio/stt/models/funasr/audio.py: log_mel_spectrogram (hamming-400, 400/160, drop last frame, HTK scale
with Slaney norm, ln(max(., 1e-10))) on the fused log-mel kernel, apply_lfr (7-stack / 6-stride with edge replication) and
the precomputed CMVN of apply_cmvn as one gather kernel (csrc/post.cu)."""
from __future__ import annotations

from ...._arrays import emit
from ...._post import cmvn_utterance, lfr
from ...._wrap import as_batch, run_frontend
from .... import _lib as L
from ....dsp import hamming, mel_filters

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
N_MELS = 80
LFR_M = 7
LFR_N = 6


def log_mel_spectrogram(audio, n_mels: int = N_MELS, n_fft: int = N_FFT, hop_length: int = HOP_LENGTH,
                        sample_rate: int = SAMPLE_RATE):
    """funasr/audio.py:32-81 -> (n_frames, n_mels)"""
    if isinstance(audio, str):
        raise NotImplementedError("file decoding is outside the DSP hot path; pass a waveform (see stt/utils.py load_audio(pcm=...))")
    ing, was_1d = as_batch(audio)
    fb = mel_filters(sample_rate, n_fft, n_mels, norm="slaney", mel_scale="htk")
    out = run_frontend(ing, hamming(n_fft), fb, n_fft=n_fft, hop=hop_length, center=True, pad_mode="reflect", drop_last=True,
                       spec_kind=L.SPEC_POWER, log_kind=L.LOG_LN, guard_kind=L.GUARD_MAX, guard_eps=1e-10)
    return emit(ing, out[0] if was_1d else out)


def apply_lfr(features, lfr_m: int = LFR_M, lfr_n: int = LFR_N):
    """funasr/audio.py:84-139 -> (ceil(n_frames / lfr_n), n_mels * lfr_m)"""
    return lfr(features, lfr_m, lfr_n)


def apply_cmvn(features, cmvn_mean=None, cmvn_istd=None):
    """funasr/audio.py:142-169.  Precomputed statistics: (features + mean) * istd (an LFR of one frame per row with stride one
    is the identity gather, so the same kernel applies the affine map); without them: per-utterance normalisation."""
    if cmvn_mean is None or cmvn_istd is None:  # per-utterance: mean / std (ddof 0) over the frames, eps 1e-6 (160-164)
        return cmvn_utterance(features, 1e-6)
    return lfr(features, 1, 1, cmvn_mean, cmvn_istd)


def preprocess_audio(audio, n_mels: int = N_MELS, lfr_m: int = LFR_M, lfr_n: int = LFR_N, cmvn_mean=None, cmvn_istd=None,
                     apply_normalization: bool = True):
    """funasr/audio.py:172-215: log-mel -> LFR (-> CMVN, fused into the LFR gather when the statistics are given)"""
    feats = log_mel_spectrogram(audio, n_mels=n_mels)
    if apply_normalization and cmvn_mean is not None and cmvn_istd is not None:
        return lfr(feats, lfr_m, lfr_n, cmvn_mean, cmvn_istd)
    if apply_normalization:
        return apply_cmvn(apply_lfr(feats, lfr_m, lfr_n), cmvn_mean, cmvn_istd)
    return apply_lfr(feats, lfr_m, lfr_n)
