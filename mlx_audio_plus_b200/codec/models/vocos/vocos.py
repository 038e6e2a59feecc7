"""Drop-in for the iSTFT head of mlx_audio/codec/models/vocos/vocos.py:119-140 (and the Soprano decoder,
tts/models/soprano/decoder.py:22-49).  The linear layer belongs to the model; this class takes its output.
"""
from __future__ import annotations

import numpy as np

from ....dsp import hanning, istft


class ISTFTHead:
    def __init__(self, dim: int, n_fft: int, hop_length: int, padding: str = "center"):
        self.n_fft = n_fft
        self.hop_length = hop_length

    def __call__(self, x):
        """x: (1, T, n_fft+2) output of the head's linear projection -> waveform ((T-1)*hop,)."""
        if type(x).__module__.split(".")[0] == "torch":
            import torch

            x = x.swapaxes(1, 2)
            mag, p = x.split(x.shape[1] // 2, dim=1)
            mag = torch.clamp(torch.exp(mag), max=1e2)
            S = torch.complex(mag * torch.cos(p), mag * torch.sin(p))
        else:
            x = np.swapaxes(np.asarray(x, dtype=np.float32), 1, 2)
            mag, p = np.split(x, 2, axis=1)
            mag = np.minimum(np.exp(mag), np.float32(1e2))
            S = (mag * (np.cos(p) + 1j * np.sin(p))).astype(np.complex64)
        return istft(S.squeeze(0), window=hanning(self.n_fft), hop_length=self.hop_length, win_length=self.n_fft)
