"""Drop-in for the iSTFT head of mlx_audio/codec/models/vocos/vocos.py:119-140 (and the Soprano decoder,
tts/models/soprano/decoder.py:22-49).  The linear layer belongs to the model; this class takes its output.

Two launches, no eager array arithmetic: `transpose_pad_kernel` turns the projection's (B, T, n_fft + 2) rows into
(B, n_fft + 2, T) — the frame axis contiguous, the layout the reference reaches with `swapaxes(1, 2)` — and the fused
polar iSTFT reads its log-magnitude half and its phase half in place: `clip(exp(mag), max=1e2)`, `cos` / `sin`, the
inverse FFT, the window, the overlap-add and the division by the window envelope all happen inside the kernel; neither
the magnitude nor the complex spectrum S ever exists in HBM.
"""
from __future__ import annotations

import numpy as np

from ...._post import transpose_pad
from ....dsp import hanning, istft_polar


class ISTFTHead:
    def __init__(self, dim: int, n_fft: int, hop_length: int, padding: str = "center"):
        self.n_fft = n_fft
        self.hop_length = hop_length

    def __call__(self, x):
        """x: (1, T, n_fft+2) output of the head's linear projection -> waveform ((T-1)*hop,).  (B, T, n_fft+2) batches give
        (B, (T-1)*hop): the reference's `S.squeeze(0)` only drops a batch axis of one."""
        is_torch = type(x).__module__.split(".")[0] == "torch"
        if is_torch:
            import torch

            xd = x if x.is_cuda else x.cuda()
            xd = xd.to(torch.float32)
        else:
            import torch

            xd = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).cuda()
        if xd.ndim != 3 or xd.shape[2] != self.n_fft + 2:
            raise ValueError(f"ISTFTHead expects (B, T, {self.n_fft + 2}), got {tuple(xd.shape)}")
        B, T, F2 = (int(v) for v in xd.shape)
        F = F2 // 2
        xt = transpose_pad(xd)  # (B, 2F, T): vocos.py:127 swapaxes, as one coalesced pass
        mag, p = xt[:, :F, :], xt[:, F:, :]  # vocos.py:128 split — views, never copied
        y = istft_polar(mag, p, self.n_fft, self.hop_length, hanning(self.n_fft), center=True, normalized=False,
                        div_clamp=False, trim_tail=True, mag_log=True, mag_clip_max=1e2, clip_stride=F2 * T)  # :129-139
        y = y[0] if B == 1 else y
        if is_torch:
            return y if x.is_cuda else y.cpu()
        from ...._arrays import DspArray

        return y.cpu().numpy().view(DspArray)
