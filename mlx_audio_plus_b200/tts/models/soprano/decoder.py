"""Drop-in for the iSTFT head of mlx_audio/tts/models/soprano/decoder.py:14-49 (n_fft 2048 / hop 512): the Vocos head with
the batch axis put back on the waveform.  The linear layer belongs to the model; this class takes its output."""
from __future__ import annotations

from ....codec.models.vocos.vocos import ISTFTHead as _VocosHead


class ISTFTHead(_VocosHead):
    def __call__(self, x):
        """x: (1, L, n_fft + 2) -> (1, (L - 1) * hop)"""
        return super().__call__(x)[None, :]
