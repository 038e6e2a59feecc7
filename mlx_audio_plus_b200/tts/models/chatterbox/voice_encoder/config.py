"""The mel front-end's view of the voice-encoder configuration (reference: tts/models/chatterbox/voice_encoder/config.py).

`melspectrogram` reads these eleven attributes and nothing else, so the reference's own `VoiceEncConfig` instance (which also
carries the LSTM's sizes) can be passed unchanged; this class exists for callers that only need the features."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class VoiceEncConfig:
    # framing / transform
    sample_rate: int = 16000
    n_fft: int = 400
    win_size: int = 400
    hop_size: int = 160
    # filterbank
    num_mels: int = 40
    fmin: int = 0
    fmax: int = 8000
    # compression
    mel_power: float = 2.0          # |X| ** mel_power before the filterbank (1.0 or 2.0 on the fused path)
    mel_type: str = "amp"           # "amp": linear mel energies; "db": 20 log10(max(mel, stft_magnitude_min))
    stft_magnitude_min: float = 1e-4
    normalized_mels: bool = False   # (mel - min_db) / (15 - min_db) with min_db = 20 log10(stft_magnitude_min)
