"""TEST INFRASTRUCTURE ONLY (see oracle/dsp_oracle.py): CPU restatement of the step in front of the DSP path — what the
reference's load_audio does after the decoder.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may
import this.

Where the arithmetic lives: `scipy.signal.resample_poly` — a third-party dependency of the reference (pyproject.toml
`scipy`, no pin; scipy 1.18.1 in this image), called at mlx_audio/stt/utils.py:27.  `resample_audio` below calls scipy
itself (it IS the reference's arithmetic); `resample_poly_restated` restates scipy's published algorithm
(scipy/signal/_signaltools.py::resample_poly + firwin + upfirdn mode="edge") in NumPy and is pinned against scipy in
tests/test_resample_cpu.py — it documents the index arithmetic the CUDA kernel follows.
"""
from math import gcd

import numpy as np


def resample_audio(audio: np.ndarray, orig_sr: int, target_sr: int) -> np.ndarray:
    """mlx_audio/stt/utils.py:21-29"""
    from scipy import signal

    g = np.gcd(orig_sr, target_sr)
    up = target_sr // g
    down = orig_sr // g
    return signal.resample_poly(audio, up, down, padtype="edge")


def pcm_to_float(samples: np.ndarray, dtype="float64") -> np.ndarray:
    """mlx_audio/audio_io.py:253-262 + always_2d (264-266): int16 (n,) or (n, ch) -> float / 32768.0, 2-D"""
    s = np.asarray(samples)
    if s.dtype == np.int16:
        s = s.astype(dtype) / 32768.0
    if s.ndim == 1:
        s = s[:, np.newaxis]
    return s


def load_audio_from_pcm(samples: np.ndarray, sample_rate: int, sr: int = 16000) -> np.ndarray:
    """mlx_audio/stt/utils.py:52-57 with audio_io.read's decoder output given: float64 samples (always_2d) ->
    resample_audio if the rates differ -> mx.array(audio, dtype=float32).mean(axis=1)"""
    audio = pcm_to_float(samples)
    if sample_rate != sr:
        audio = resample_audio(audio, sample_rate, sr)
    a32 = np.asarray(audio, dtype=np.float32)
    return a32.sum(axis=1, dtype=np.float32) / np.float32(a32.shape[1]) if a32.shape[1] > 1 else a32[:, 0]


def resample_poly_restated(x: np.ndarray, up: int, down: int) -> np.ndarray:
    """scipy.signal.resample_poly(x, up, down, padtype="edge") for 1-D x, restated (float64): Kaiser(5.0) windowed sinc of
    2 * 10 * max(up, down) + 1 taps with unit DC gain, times `up`; n_pre_pad zeros in front so that output n sits at
    upfirdn index n + n_pre_remove; input extended by its edge samples."""
    g = gcd(int(up), int(down))
    up, down = int(up) // g, int(down) // g
    x = np.asarray(x, dtype=np.float64)
    if up == 1 and down == 1:
        return x.copy()
    max_rate = max(up, down)
    f_c = 1.0 / max_rate
    half_len = 10 * max_rate
    numtaps = 2 * half_len + 1
    alpha = 0.5 * (numtaps - 1)
    n = np.arange(numtaps, dtype=np.float64)
    h = f_c * np.sinc(f_c * (n - alpha))
    h *= np.i0(5.0 * np.sqrt(np.clip(1.0 - ((n - alpha) / alpha) ** 2, 0.0, None))) / np.i0(5.0)
    h /= h.sum()
    h *= up
    n_pre_pad = down - half_len % down
    n_pre_remove = (half_len + n_pre_pad) // down
    hp = np.concatenate([np.zeros(n_pre_pad), h])
    n_in = x.shape[0]
    n_out = n_in * up // down + bool(n_in * up % down)
    t = (np.arange(n_out, dtype=np.int64) + n_pre_remove) * down
    ph, i0 = t % up, t // up
    y = np.zeros(n_out, np.float64)
    for j in range(-(-len(hp) // up)):
        k = ph + j * up
        w = np.where(k < len(hp), hp[np.minimum(k, len(hp) - 1)], 0.0)
        y += w * x[np.clip(i0 - j, 0, n_in - 1)]
    return y
