"""Drop-in for mlx_audio/tts/models/chatterbox/voice_encoder/melspec.py:13-77: symmetric "hann", reflect-centred STFT,
|X| ** mel_power, Slaney/slaney filterbank, optional 20 log10(max(., stft_magnitude_min)) and level normalisation,
(M, T') or (B, M, T').  The batch runs in one launch; dB and the normalisation fold into the kernel's affine epilogue:
(20 log10 m - min_db) / (15 - min_db) = (log10 m - min_db / 20) / ((15 - min_db) / 20)."""
from __future__ import annotations

import math

from ....._arrays import emit
from ....._wrap import as_batch, run_frontend
from ..... import _lib as L
from .....dsp import hanning, mel_filters
from .config import VoiceEncConfig


def melspectrogram(wav, hp: VoiceEncConfig = VoiceEncConfig(), pad: bool = True):
    ing, was_1d = as_batch(wav)
    if hp.mel_power == 2.0:
        spec_kind = L.SPEC_POWER
    elif hp.mel_power == 1.0:
        spec_kind = L.SPEC_MAGNITUDE
    else:
        raise NotImplementedError(f"mel_power={hp.mel_power}: the fused kernel computes |X| or |X|**2 (the shipped configs)")
    fb = mel_filters(sample_rate=hp.sample_rate, n_fft=hp.n_fft, n_mels=hp.num_mels, f_min=hp.fmin, f_max=hp.fmax,
                     norm="slaney", mel_scale="slaney")
    kw = {}
    if hp.mel_type == "db":
        kw.update(log_kind=L.LOG_LOG10, guard_kind=L.GUARD_MAX, guard_eps=float(hp.stft_magnitude_min))
        if hp.normalized_mels:
            min_db = 20 * math.log10(hp.stft_magnitude_min)
            kw.update(affine_add=-min_db / 20.0, affine_div=(15.0 - min_db) / 20.0)
        else:
            kw.update(affine_add=0.0, affine_div=1.0 / 20.0)
    elif hp.normalized_mels:  # amplitude mels through the same level map (melspec.py:71-75)
        min_db = 20 * math.log10(hp.stft_magnitude_min)
        kw.update(affine_add=-min_db, affine_div=15.0 - min_db)
    out = run_frontend(ing, hanning(hp.win_size), fb, n_fft=hp.n_fft, hop=hp.hop_size, center=True, pad_mode="reflect",
                       spec_kind=spec_kind, out_layout=L.LAYOUT_MT, **kw)
    return emit(ing, out[0] if was_1d else out)
