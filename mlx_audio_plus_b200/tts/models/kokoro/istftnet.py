"""Drop-in for the STFT helper of mlx_audio/tts/models/kokoro/istftnet.py:399-528 (mlx_angle, mlx_unwrap,
MLXSTFT.transform / inverse; n_fft=20, hop=5 in Kokoro).  transform/inverse process the whole batch in one
launch each (the reference loops over batch items, istftnet.py:471-490, 500-519)."""
from __future__ import annotations

import math

import numpy as np

from ....dsp import _resolve_window, istft_polar, stft


def _is_torch(x):
    return type(x).__module__.split(".")[0] == "torch"


def mlx_angle(z, deg=False):
    if _is_torch(z):
        import torch

        a = torch.atan2(z.imag, z.real) if z.is_complex() else torch.atan2(torch.zeros_like(z), z)
    else:
        z = np.asarray(z)
        a = np.arctan2(z.imag, z.real).astype(np.float32) if np.iscomplexobj(z) else np.arctan2(np.zeros_like(z), z)
    return a * (180.0 / math.pi) if deg else a


def mlx_unwrap(p, discont=None, axis=-1, period=2 * math.pi):
    """Phase unwrap (reference istftnet.py:418-452): an float32 prefix sum of 2*pi corrections."""
    if discont is None:
        discont = period / 2
    discont = max(discont, period / 2)
    hi, lo = period / 2, -period / 2
    if _is_torch(p) and p.is_cuda:  # one kernel (csrc/post.cu unwrap_rows_kernel) instead of seven eager passes over (B, F, T)
        from ...._post import unwrap

        return unwrap(p, discont, period, axis)
    if _is_torch(p):
        import torch

        return torch.from_numpy(mlx_unwrap(p.detach().numpy(), discont, axis, period))
    p = np.asarray(p, dtype=np.float32)
    dd = np.diff(p, axis=axis).astype(np.float32)
    ddmod = (dd - np.float32(period) * np.floor((dd - np.float32(lo)) / np.float32(period))).astype(np.float32)
    ddmod = np.where((np.abs(dd - np.float32(hi)) < 1e-10) & (dd > 0), np.float32(hi), ddmod)
    corr = np.where(np.abs(dd) < discont, np.float32(0), (ddmod - dd).astype(np.float32))
    shape = list(corr.shape)
    shape[axis] = 1
    corr = np.concatenate([np.zeros(shape, np.float32), corr], axis=axis)
    return (p + np.cumsum(corr, axis=axis, dtype=np.float32)).astype(np.float32)


class MLXSTFT:
    def __init__(self, filter_length=800, hop_length=200, win_length=800, window="hann"):
        self.filter_length = filter_length
        self.hop_length = hop_length
        self.win_length = win_length
        self.window = window

    def transform(self, input_data):
        if input_data.ndim == 1:
            input_data = input_data[None, :]
        spec = stft(input_data, n_fft=self.filter_length, hop_length=self.hop_length, win_length=self.win_length,
                    window=self.window, center=True, pad_mode="reflect")  # (B, T, F)
        spec = spec.swapaxes(1, 2)
        return abs(spec), mlx_angle(spec)

    def inverse(self, magnitude, phase):
        phase_cont = mlx_unwrap(phase, axis=2)
        # mag * cos / sin of the unwrapped phase (istftnet.py:505-507) is formed inside the iSTFT kernel; string
        # windows are periodic in istft (dsp.py:172-176), sum-w envelope with the where-guard (dsp.py:199-209)
        w = _resolve_window(self.window, self.win_length, periodic_trick=True)
        audio = istft_polar(magnitude, phase_cont, self.filter_length, self.hop_length, w, center=True,
                            normalized=False, div_clamp=False, trim_tail=True)  # (B, L)
        return audio[:, None, :]

    def __call__(self, input_data):
        self.magnitude, self.phase = self.transform(input_data)
        reconstruction = self.inverse(self.magnitude, self.phase)
        return reconstruction[..., None, :] if _is_torch(reconstruction) else np.expand_dims(reconstruction, -2)
