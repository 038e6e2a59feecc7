"""The step in front of the DSP path — what the reference's ``load_audio`` does AFTER the decoder
(mlx_audio/stt/utils.py:21-57, mlx_audio/audio_io.py:258-262), on the GPU:

    int16 PCM / 32768  ->  scipy.signal.resample_poly(audio, up, down, padtype="edge") per channel  ->  mean over channels

in ONE kernel (csrc/resample.cu) that reads the interleaved PCM once and writes float32 mono once.  File decoding
(miniaudio / ffmpeg) stays out of scope: ``load_audio`` takes the decoder's output (``pcm=``, ``sample_rate=``).

The polyphase filter is designed here with NumPy, formula for formula what scipy.signal.resample_poly / firwin compute
(Kaiser beta 5.0 windowed sinc, 2 * 10 * max(up, down) + 1 taps, unit DC gain, times ``up``, zero-padded in front so the
output is centred) — tests/test_resample_cpu.py pins it against scipy itself.  The product does not import scipy.
"""
from __future__ import annotations

import ctypes as C
from functools import lru_cache
from math import gcd

import numpy as np

from .. import _lib as L
from .._arrays import _is_torch

SAMPLE_RATE = 16000  # whisper/audio.py:16


@lru_cache(maxsize=64)
def resample_poly_design(up: int, down: int):
    """(up, down, taps[J][up] float32, J, n_pre_remove) for scipy.signal.resample_poly(x, up, down) — the filter of
    scipy/signal/_signaltools.py::resample_poly with window=("kaiser", 5.0), split into its `up` polyphase branches:
    out[n] = sum_j taps[j][t % up] * x_edge[t // up - j],  t = (n + n_pre_remove) * down."""
    up, down = int(up), int(down)
    if up < 1 or down < 1:
        raise ValueError("up and down must be >= 1")  # scipy raises the same
    g = gcd(up, down)
    up //= g
    down //= g
    if up == 1 and down == 1:
        return 1, 1, np.ones((1, 1), np.float32), 1, 0
    max_rate = max(up, down)
    f_c = 1.0 / max_rate
    half_len = 10 * max_rate
    numtaps = 2 * half_len + 1
    alpha = 0.5 * (numtaps - 1)
    n = np.arange(numtaps, dtype=np.float64)
    h = f_c * np.sinc(f_c * (n - alpha))  # firwin: lowpass band (0, f_c), fs = 2
    h *= np.i0(5.0 * np.sqrt(np.clip(1.0 - ((n - alpha) / alpha) ** 2, 0.0, None))) / np.i0(5.0)  # kaiser(numtaps, 5.0)
    h /= h.sum()  # scale=True: unit gain at DC
    h *= up
    n_pre_pad = down - half_len % down
    n_pre_remove = (half_len + n_pre_pad) // down
    hp = np.concatenate([np.zeros(n_pre_pad), h])
    J = -(-len(hp) // up)
    taps = np.zeros(J * up, np.float64)
    taps[: len(hp)] = hp
    return up, down, np.ascontiguousarray(taps.reshape(J, up), dtype=np.float32), J, int(n_pre_remove)


class _Resampler:
    def __init__(self, up, down):
        self.up, self.down, taps, self.J, self.pre = resample_poly_design(up, down)
        self._h = C.c_void_p()
        L.check(L.lib.b2a_resampler_create(self.up, self.down, self.J, self.pre, taps.ctypes.data_as(C.c_void_p), C.byref(self._h)))

    def out_len(self, n_in):
        return -(-n_in * self.up // self.down)

    def __del__(self):
        try:
            if self._h:
                L.lib.b2a_resampler_destroy(self._h)
        except Exception:
            pass


_PLANS = {}


def _resampler(up, down, dev_index):
    key = (up, down, dev_index)
    if key not in _PLANS:
        _PLANS[key] = _Resampler(up, down)
    return _PLANS[key]


def _run(x, up, down, mono):
    """x: (n,), (n, ch) or (B, n, ch); numpy or torch; float or int16.  Returns float32 in the caller's family."""
    import torch

    if L.lib.b2a_device_count() < 1:
        raise L.B2AError("b200audio: no CUDA device — there is no CPU fallback")
    is_t = _is_torch(x)
    t = x if is_t else torch.from_numpy(np.ascontiguousarray(x))
    squeeze_ch = t.ndim == 1
    if squeeze_ch:
        t = t[:, None]
    batched = t.ndim == 3
    if not batched:
        t = t[None]
    if t.dtype == torch.int16:
        kind = L.PCM_I16
    else:
        kind = L.PCM_F32
        t = t.to(torch.float32)
    t = t.contiguous()
    if not t.is_cuda:
        t = t.cuda()
    B, n_in, ch = t.shape
    with torch.cuda.device(t.device):
        r = _resampler(up, down, t.device.index or 0)
        n_out = r.out_len(n_in)
        out = torch.empty((B, n_out) if mono else (B, n_out, ch), dtype=torch.float32, device=t.device)
        if n_in > 0 and n_out > 0:
            a = L.ResampleArgs()
            a.inp, a.out, a.n_in = t.data_ptr(), out.data_ptr(), n_in
            a.in_clip_stride = a.out_clip_stride = 0
            a.batch, a.channels, a.in_kind, a.mono = B, ch, kind, int(bool(mono))
            L.check(L.lib.b2a_resample(r._h, C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    if not batched:
        out = out[0]
    if squeeze_ch and not mono:
        out = out[..., 0]
    if is_t:
        return out if x.is_cuda else out.cpu()
    return out.cpu().numpy()


def resample_audio(audio, orig_sr: int, target_sr: int):
    """stt/utils.py:21-29 — ``signal.resample_poly(audio, up, down, padtype="edge")`` along axis 0, every channel.
    Returns float32 (the reference keeps the input dtype, float64 after audio_io.read; load_audio casts to float32 next)."""
    g = int(np.gcd(int(orig_sr), int(target_sr)))
    return _run(audio, int(target_sr) // g, int(orig_sr) // g, mono=False)


def read_wav_pcm16(file):
    """Host-side container parse for the one format that needs no codec: RIFF / WAVE with 16-bit integer PCM — what
    audio_io.read yields for it through miniaudio (interleaved int16, audio_io.py:250-262).  -> ((n,) or (n, channels) int16,
    sample_rate).  Anything else (mp3, flac, m4a, other sample widths) needs a decoder and stays upstream of the path."""
    import wave

    try:
        with wave.open(file, "rb") as w:
            ch, width, rate, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
            if width != 2 or w.getcomptype() != "NONE":
                raise NotImplementedError(f"b200audio: only 16-bit PCM WAVE is parsed here (sample width {width} bytes)")
            data = w.readframes(n)
    except wave.Error as e:
        raise NotImplementedError(f"b200audio: not a PCM WAVE file ({e}); decode it upstream and pass pcm= / sample_rate=")
    pcm = np.frombuffer(data, dtype="<i2").copy()  # writable: torch.from_numpy wraps it downstream
    return (pcm.reshape(-1, ch) if ch > 1 else pcm), int(rate)


def load_audio(file=None, sr: int = SAMPLE_RATE, from_stdin=False, dtype=None, *, pcm=None, sample_rate=None):
    """stt/utils.py:32-57 from the decoder's output on: ``pcm`` is what audio_io.read decodes — interleaved int16 of shape
    (n,) or (n, channels) (or float samples already divided by 32768) — at ``sample_rate``.  int16 / 32768 -> resample to
    ``sr`` if the rates differ -> mean over channels, float32 mono, in one kernel.  A 16-bit PCM WAVE path (or file object)
    is parsed on the host (read_wav_pcm16); every other container needs a decoder, which is upstream of the path."""
    if pcm is None and file is not None and not from_stdin:
        pcm, sample_rate = read_wav_pcm16(file)
    if pcm is None:
        raise NotImplementedError("b200audio: file decoding (miniaudio / ffmpeg) is upstream of the DSP path; pass pcm= and "
                                  "sample_rate= (audio_io.read's output)")
    if sample_rate is None:
        raise ValueError("load_audio(pcm=...) needs sample_rate=")
    g = int(np.gcd(int(sample_rate), int(sr)))
    return _run(pcm, int(sr) // g, int(sample_rate) // g, mono=True)
