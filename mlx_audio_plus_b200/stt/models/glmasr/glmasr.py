"""Drop-in for the feature step of mlx_audio/stt/models/glmasr/glmasr.py:547-589 (`Model._preprocess_audio`): the Whisper
chain with 128 mel bins, returned with a leading batch axis, (1, T, n_mels); a 3-D input is taken as features already."""
from __future__ import annotations

from ..whisper.audio import log_mel_spectrogram

N_FFT = 400
HOP_LENGTH = 160


def preprocess_audio(audio, n_mels: int = 128):
    if isinstance(audio, str):
        raise NotImplementedError("file decoding (load_audio) is outside the DSP hot path; pass a waveform array")
    if getattr(audio, "ndim", 1) == 3:  # glmasr.py:569-570
        return audio
    return log_mel_spectrogram(audio, n_mels=n_mels)[None]
