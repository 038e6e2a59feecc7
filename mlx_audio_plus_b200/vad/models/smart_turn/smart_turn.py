"""Drop-in for the feature path of mlx_audio/vad/models/smart_turn/smart_turn.py:158-229: keep the last
`max_audio_seconds` of the turn (left zero padding when shorter), normalise the waveform to zero mean / unit variance,
Whisper log-mel, keep / left-pad to the target frame count, return (n_mels, frames)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ...._arrays import _is_torch
from ....stt.models.whisper.audio import log_mel_spectrogram


@dataclass
class ProcessorConfig:  # vad/models/smart_turn/config.py:23-32
    sampling_rate: int = 16000
    max_audio_seconds: int = 8
    n_fft: int = 400
    hop_length: int = 160
    n_mels: int = 80
    normalize_audio: bool = True


def prepare_audio_array(audio, config: ProcessorConfig = ProcessorConfig()):
    """smart_turn.py:158-201 after decoding / resampling: 1-D waveform -> (max_samples,) float32, same family as the input."""
    if isinstance(audio, str):
        raise NotImplementedError("file decoding is outside the DSP hot path; pass a waveform array")
    max_samples = config.max_audio_seconds * config.sampling_rate
    if _is_torch(audio) and audio.is_cuda:
        import torch

        a = audio.detach().to(torch.float32)
        if a.ndim != 1:
            raise ValueError(f"Expected mono audio (1-D), got shape {tuple(a.shape)}")
        if a.shape[0] > max_samples:
            a = a[-max_samples:]
        elif a.shape[0] < max_samples:
            a = torch.nn.functional.pad(a, (max_samples - a.shape[0], 0))
        if config.normalize_audio and a.numel() > 0:  # (a - mean) / max(std, 1e-7), smart_turn.py:196-199: one launch
            from ...._post import rows_normalize

            a = rows_normalize(a.contiguous(), den_kind=1, eps=1e-7)
        return a.contiguous()
    a = np.asarray(audio.detach().numpy() if _is_torch(audio) else audio, dtype=np.float32)
    if a.ndim != 1:
        raise ValueError(f"Expected mono audio (1-D), got shape {a.shape}")
    if a.shape[0] > max_samples:
        a = a[-max_samples:]
    elif a.shape[0] < max_samples:
        a = np.pad(a, (max_samples - a.shape[0], 0), mode="constant")
    if config.normalize_audio and a.size > 0:
        a = (a - float(a.mean())) / max(float(a.std()), 1e-7)
    return a.astype(np.float32, copy=False)


def prepare_input_features(audio, config: ProcessorConfig = ProcessorConfig(), dtype="float32"):
    """smart_turn.py:203-229 -> (n_mels, max_audio_seconds * sampling_rate // hop_length)"""
    mel = log_mel_spectrogram(prepare_audio_array(audio, config), n_mels=config.n_mels)  # (T, n_mels)
    target = config.max_audio_seconds * config.sampling_rate // config.hop_length
    frames = mel.shape[0]
    if frames > target:
        mel = mel[-target:, :]
    elif frames < target:  # LEFT zero rows (smart_turn.py:224-225)
        if _is_torch(mel):
            import torch

            mel = torch.nn.functional.pad(mel, (0, 0, target - frames, 0))
        else:
            mel = np.pad(np.asarray(mel), [(target - frames, 0), (0, 0)])
    out = mel.T
    if _is_torch(out):
        import torch

        return out.to(getattr(torch, str(dtype).split(".")[-1]))
    return np.asarray(out).astype(str(dtype).split(".")[-1])
