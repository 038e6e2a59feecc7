"""Drop-in for the reference's shared DSP module `mlx_audio/dsp.py` (windows, stft, istft, ISTFTCache,
mel_filters) — same names, signatures, defaults, quirks and error behaviour, computed by the sm_100a
kernels behind include/b200audio.h.  Like the reference module this file imports nothing from the
model packages (mlx_audio/tests/test_dsp.py:7-24).

Array families: numpy in -> numpy out (a thin ndarray subclass carrying the mx.array methods the
reference's callers chain: .abs() .square() .log10() ...); torch CUDA in -> torch CUDA out (zero copy);
mlx in -> mlx out (buffer protocol).  Extension over the reference: `stft` also accepts (B, L) and `istft`
(B, F, T) and processes the batch in ONE launch (the reference loops in Python, dsp.py:131 is 1-D only).
"""
from __future__ import annotations

import ctypes as _C
from functools import lru_cache
from typing import Optional

import numpy as np

from . import _lib as _L
from ._arrays import DspArray, emit, host_window, ingest
from .frontend import FrontendPlan, IstftPlan, _device_index, cached_plan

__all__ = [
    "hanning",
    "hamming",
    "blackman",
    "bartlett",
    "STR_TO_WINDOW_FN",
    "stft",
    "istft",
    "ISTFTCache",
    "mel_filters",
    # Kaldi-compatible features (reference dsp.py:439-676)
    "compute_deltas_kaldi",
    "mel_scale_kaldi",
    "inverse_mel_scale_kaldi",
    "get_mel_banks_kaldi",
    "compute_fbank_kaldi",
]


def _window(kind: int, size, periodic) -> DspArray:
    out = np.empty(int(size), dtype=np.float32)
    _L.check(_L.lib.b2a_window(kind, int(size), int(bool(periodic)), out.ctypes.data_as(_C.c_void_p)))
    out.setflags(write=False)  # lru-cached shared object: callers must not mutate (SURVEY §8b ownership)
    return out.view(DspArray)


# dsp.py:33-79 — lru-cached like the reference, so arguments must be hashable
@lru_cache(maxsize=None)
def hanning(size, periodic=False):
    """Hanning (Hann) window; symmetric unless periodic=True (reference dsp.py:33-44)."""
    return _window(_L.WIN_HANN, size, periodic)


@lru_cache(maxsize=None)
def hamming(size, periodic=False):
    """Hamming window (reference dsp.py:47-58)."""
    return _window(_L.WIN_HAMMING, size, periodic)


@lru_cache(maxsize=None)
def blackman(size, periodic=False):
    """Blackman window (reference dsp.py:61-72)."""
    return _window(_L.WIN_BLACKMAN, size, periodic)


@lru_cache(maxsize=None)
def bartlett(size, periodic=False):
    """Bartlett (triangular) window (reference dsp.py:75-79)."""
    return _window(_L.WIN_BARTLETT, size, periodic)


STR_TO_WINDOW_FN = {  # reference dsp.py:82-88
    "hann": hanning,
    "hanning": hanning,
    "hamming": hamming,
    "blackman": blackman,
    "bartlett": bartlett,
}


def _resolve_window(window, size, periodic_trick: bool) -> np.ndarray:
    if isinstance(window, str):
        fn = STR_TO_WINDOW_FN.get(window.lower())
        if fn is None:
            raise ValueError(f"Unknown window function: {window}")  # dsp.py:109 / 175
        # stft: symmetric window_fn(win_length) (dsp.py:110); istft: window_fn(win_length+1)[:-1] (dsp.py:176)
        return np.asarray(fn(size + 1)[:-1] if periodic_trick else fn(size))
    return host_window(window)


def stft(x, n_fft=800, hop_length=None, win_length=None, window="hann", center=True, pad_mode="reflect"):
    """Short-time Fourier transform; reference dsp.py:92-141.  Returns complex64 (T, n_fft//2+1)."""
    if hop_length is None:
        hop_length = n_fft // 4
    if win_length is None:
        win_length = n_fft
    w = _resolve_window(window, win_length, periodic_trick=False)  # array windows ignore win_length
    if w.shape[0] > n_fft:  # `frames * w` cannot broadcast in the reference (dsp.py:141)
        raise ValueError(f"window of {w.shape[0]} taps cannot be broadcast against frames of n_fft={n_fft}")
    if center and pad_mode not in ("constant", "reflect"):
        raise ValueError(f"Invalid pad_mode {pad_mode}")  # dsp.py:126
    ing = ingest(x, "float32")
    squeeze = ing.data.ndim == 1
    if squeeze:
        ing.data = ing.data.reshape(1, -1)
    elif ing.data.ndim != 2:
        raise ValueError("stft expects a 1-D signal (or a (B, L) batch)")
    plan = cached_plan(FrontendPlan, _device_index(ing), w, n_fft=int(n_fft), hop=int(hop_length),
                       center=bool(center), pad_mode=pad_mode if center else "reflect")
    out = plan.run(ing)
    return emit(ing, out[0] if squeeze else out)


def istft(x, hop_length=None, win_length=None, window="hann", center=True, length=None, normalized=False):
    """Inverse STFT with windowed overlap-add; reference dsp.py:144-217.  x: complex (n_fft//2+1, T)."""
    ing = ingest(x, "complex64")
    squeeze = ing.data.ndim == 2
    if squeeze:
        ing.data = ing.data.reshape((1,) + tuple(ing.data.shape))
    elif ing.data.ndim != 3:
        raise ValueError("istft expects (F, T) (or a (B, F, T) batch)")
    F, T = int(ing.data.shape[1]), int(ing.data.shape[2])
    if win_length is None:
        win_length = (T - 1) * 2  # dsp.py:168 reads the frame axis — reproduced on purpose
    if hop_length is None:
        hop_length = win_length // 4
    w = _resolve_window(window, win_length, periodic_trick=True)
    n_time = 2 * (F - 1)  # irfft length (dsp.py:190)
    eff = max(w.shape[0], win_length)  # dsp.py:180-181 right-pads only up to win_length
    if eff != n_time or win_length != n_time:
        raise ValueError(
            f"istft: window/win_length ({w.shape[0]}/{win_length}) cannot be broadcast against irfft frames of "
            f"{n_time} samples"
        )
    plan = cached_plan(IstftPlan, _device_index(ing), w, n_fft=n_time, hop=int(hop_length), center=bool(center),
                       normalized=bool(normalized), div_clamp=False, trim_tail=True)
    out = plan.run(ing, length=length)
    return emit(ing, out[0] if squeeze else out)


def istft_polar(magnitude, phase, n_fft, hop_length, window, *, center=True, normalized=False, div_clamp=False,
                div_eps=0.0, trim_tail=True, length=None, mag_clip_max=0.0, mag_clip_min_zero=False, mag_log=False,
                clip_stride=0):
    """iSTFT of a spectrum given as (magnitude, phase) planes of shape (B, F, T): the fused form of the
    `mag*cos(phase) + 1j*mag*sin(phase)` -> istft sequences in kokoro/istftnet.py:500-519, s3gen/hifigan.py:480-549
    and cosyvoice3/hifigan.py:447-499 — clip, cos / sin, inverse FFT, window, overlap-add and envelope division run
    in ONE kernel; the complex spectrum never exists in HBM.  Not a reference name: the model-local drop-ins call it.
    `mag_log`: the magnitude plane holds log-magnitudes (`clip(exp(x), max=1e2)` of the Vocos head, vocos.py:129-130);
    `clip_stride`: elements between clips when the two planes are halves of one (B, 2F, T) device buffer."""
    m, ph = ingest(magnitude, "float32"), ingest(phase, "float32")
    if clip_stride and _is_strided_pair(magnitude, phase):  # views into one buffer: no copies, explicit clip stride
        m.data, ph.data = magnitude, phase
    if m.data.ndim != 3 or tuple(m.data.shape) != tuple(ph.data.shape):
        raise ValueError("istft_polar expects magnitude / phase of identical shape (batch, freq, time)")
    if m.on_device != ph.on_device:
        raise ValueError("magnitude and phase must live on the same device")
    w = host_window(window)
    if w.shape[0] > n_fft or m.data.shape[1] != n_fft // 2 + 1:
        raise ValueError("istft_polar: window / spectrum do not match n_fft")
    plan = cached_plan(IstftPlan, _device_index(m), w, n_fft=int(n_fft), hop=int(hop_length), center=bool(center),
                       normalized=bool(normalized), div_clamp=bool(div_clamp), trim_tail=bool(trim_tail),
                       div_eps=float(div_eps), polar=True, mag_clip_max=float(mag_clip_max),
                       mag_clip_min_zero=bool(mag_clip_min_zero), mag_log=bool(mag_log))
    return emit(m, plan.run(m, imag=ph, length=length, clip_stride=int(clip_stride)))


def _is_strided_pair(a, b):
    """Two float32 CUDA views of shape (B, F, T) whose (F, T) planes are dense: usable with an explicit clip stride."""
    if type(a).__module__.split(".")[0] != "torch" or not (a.is_cuda and b.is_cuda):
        return False
    import torch

    ok = lambda t: t.dtype == torch.float32 and t.ndim == 3 and t.stride(2) == 1 and t.stride(1) == t.shape[2]  # noqa: E731
    return ok(a) and ok(b) and a.stride(0) == b.stride(0)


@lru_cache(maxsize=None)
def mel_filters(
    sample_rate: int,
    n_fft: int,
    n_mels: int,
    f_min: float = 0,
    f_max: Optional[float] = None,
    norm: Optional[str] = None,
    mel_scale: str = "htk",
):
    """Triangular mel filterbank (n_mels, n_fft//2+1); reference dsp.py:223-296.  mel_scale other than
    "htk" (None included) means Slaney; norm other than exactly "slaney" means no area normalisation."""
    out = np.empty((int(n_mels), int(n_fft) // 2 + 1), dtype=np.float32)
    _L.check(_L.lib.b2a_mel_filters(int(sample_rate), int(n_fft), int(n_mels), float(f_min),
                                    float(f_max) if f_max else 0.0, int(norm == "slaney"),
                                    int(mel_scale == "htk"), out.ctypes.data_as(_C.c_void_p)))
    out.setflags(write=False)
    return out.view(DspArray)


class ISTFTCache:
    """Batched iSTFT with cached normalisation buffers; reference dsp.py:299-431.  Always window^2
    envelope clamped at 1e-10, strips only the leading half window.  The envelope is recomputed inside the
    fused kernel (it is position-only), so the caches below keep the reference's bookkeeping API
    (cache_info / clear_cache, sts/tests/test_mossformer2_se.py:98-117) and serve get_norm_buffer /
    get_positions callers, but the hot path never reads them."""

    def __init__(self):
        self.norm_buffer_cache = {}
        self.position_cache = {}

    def get_positions(self, num_frames: int, frame_length: int, hop_length: int):
        key = (num_frames, frame_length, hop_length)
        if key not in self.position_cache:
            pos = np.arange(num_frames, dtype=np.int32)[:, None] * hop_length + np.arange(frame_length, dtype=np.int32)[None, :]
            self.position_cache[key] = pos.reshape(-1).view(DspArray)
        return self.position_cache[key]

    def get_norm_buffer(self, n_fft: int, hop_length: int, win_length: int, window, num_frames: int):
        w = host_window(window)
        key = (n_fft, hop_length, win_length, hash(tuple(w.tolist())), num_frames)
        if key not in self.norm_buffer_cache:
            frame_length = w.shape[0]
            self.get_positions(num_frames, frame_length, hop_length)
            env = np.zeros((num_frames - 1) * hop_length + frame_length, dtype=np.float32)
            w2 = w * w
            for t in range(num_frames):  # frame-ordered accumulation, as a sequential scatter-add
                env[t * hop_length : t * hop_length + frame_length] += w2
            self.norm_buffer_cache[key] = np.maximum(env, np.float32(1e-10)).view(DspArray)
        return self.norm_buffer_cache[key]

    def istft(self, real_part, imag_part, n_fft: int, hop_length: int, win_length: int, window,
              center: bool = True, audio_length: int = None):
        re, im = ingest(real_part, "float32"), ingest(imag_part, "float32")
        if re.data.ndim != 3 or tuple(re.data.shape) != tuple(im.data.shape):
            raise ValueError("ISTFTCache.istft expects real/imag of identical shape (batch, freq, time)")
        if re.on_device != im.on_device:
            raise ValueError("real_part and imag_part must live on the same device")
        w = host_window(window)
        if w.shape[0] > n_fft or re.data.shape[1] != n_fft // 2 + 1:
            raise ValueError("ISTFTCache.istft: window / spectrum do not match n_fft")
        plan = cached_plan(IstftPlan, _device_index(re), w, n_fft=int(n_fft), hop=int(hop_length), center=bool(center),
                           normalized=True, div_clamp=True, trim_tail=False)
        out = plan.run(re, imag=im, length=audio_length)
        return emit(re, out)

    def clear_cache(self):
        self.norm_buffer_cache.clear()
        self.position_cache.clear()

    def cache_info(self):
        nb, pi = len(self.norm_buffer_cache), len(self.position_cache)
        return {"norm_buffers": nb, "position_indices": pi, "total_cached_items": nb + pi}


# ---- Kaldi-compatible features: reference dsp.py:439-676 -------------------------------------------------------
def compute_deltas_kaldi(specgram, win_length: int = 5, mode: str = "edge"):
    """Delta coefficients along the last axis, d_t = sum_n n (c_{t+n} - c_{t-n}) / (2 sum_n n^2); reference
    dsp.py:439-483 (a Python loop over time steps there, one launch here)."""
    if win_length < 3:
        raise ValueError(f"win_length should be >= 3, got {win_length}")  # dsp.py:456-457
    if win_length % 2 == 0:
        # the reference multiplies a win_length-wide slice by 2 * ((win_length - 1) // 2) + 1 weights (dsp.py:473-480): for an
        # even win_length the shapes do not broadcast and it raises; same exception type here instead of a silent n = (w-1)//2
        raise ValueError(f"win_length should be odd, got {win_length}: shapes ({win_length},) and ({win_length - 1},) cannot be broadcast")
    ing = ingest(specgram, "float32")
    shape = tuple(int(v) for v in ing.data.shape)
    cols = shape[-1]
    rows = int(np.prod(shape[:-1])) if len(shape) > 1 else 1
    edge = 1 if mode == "edge" else 0
    if ing.on_device:
        import torch

        x = ing.data.contiguous()
        out = torch.empty_like(x)
        with torch.cuda.device(ing.device):
            st = torch.cuda.current_stream(ing.device).cuda_stream
            _L.check(_L.lib.b2a_deltas(x.data_ptr(), out.data_ptr(), rows, cols, int(win_length), edge, _C.c_void_p(st)))
        return emit(ing, out)
    import torch

    dev = _device_index(None)
    with torch.cuda.device(dev):
        x = torch.from_numpy(np.ascontiguousarray(ing.data)).cuda(dev)
        out = torch.empty_like(x)
        st = torch.cuda.current_stream(dev).cuda_stream
        _L.check(_L.lib.b2a_deltas(x.data_ptr(), out.data_ptr(), rows, cols, int(win_length), edge, _C.c_void_p(st)))
        return emit(ing, out.cpu().numpy())


def mel_scale_kaldi(freq):
    """1127 ln(1 + f/700) in float32 (reference dsp.py:486-488)."""
    f = np.asarray(freq, dtype=np.float32)
    return (np.float32(1127.0) * np.log(np.float32(1.0) + f / np.float32(700.0))).astype(np.float32).view(DspArray)


def inverse_mel_scale_kaldi(mel_freq):
    """700 (exp(m/1127) - 1) in float32 (reference dsp.py:491-493)."""
    m = np.asarray(mel_freq, dtype=np.float32)
    return (np.float32(700.0) * (np.exp(m / np.float32(1127.0)) - np.float32(1.0))).astype(np.float32).view(DspArray)


def _next_power_of_2(x: int) -> int:
    return 1 if x == 0 else 2 ** (x - 1).bit_length()  # dsp.py:496-498


def get_mel_banks_kaldi(num_bins: int, window_length_padded: int, sample_freq: float, low_freq: float,
                        high_freq: float):
    """Kaldi mel filterbank (num_bins, n_fft/2) and the centre frequencies; reference dsp.py:526-574 (float32
    array arithmetic in the reference's order; the band edges are Python floats there)."""
    assert num_bins > 3, "Must have at least 3 mel bins"
    assert window_length_padded % 2 == 0
    f32 = np.float32
    num_fft_bins = window_length_padded // 2
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    assert (0.0 <= low_freq < nyquist) and (0.0 < high_freq <= nyquist)
    fft_bin_width = sample_freq / window_length_padded
    mel_low = float(mel_scale_kaldi(f32(low_freq)))
    mel_high = float(mel_scale_kaldi(f32(high_freq)))
    delta = (mel_high - mel_low) / (num_bins + 1)
    idx = np.arange(num_bins, dtype=np.int32).reshape(-1, 1)
    left = (f32(mel_low) + idx.astype(f32) * f32(delta)).astype(f32)
    center = (f32(mel_low) + (idx.astype(f32) + f32(1.0)) * f32(delta)).astype(f32)
    right = (f32(mel_low) + (idx.astype(f32) + f32(2.0)) * f32(delta)).astype(f32)
    center_freqs = np.asarray(inverse_mel_scale_kaldi(center))
    mel = np.asarray(mel_scale_kaldi(f32(fft_bin_width) * np.arange(num_fft_bins, dtype=np.int32).astype(f32))).reshape(1, -1)
    up = ((mel - left) / (center - left)).astype(f32)
    down = ((right - mel) / (right - center)).astype(f32)
    bins = np.maximum(np.zeros(1, f32), np.minimum(up, down)).astype(f32)
    return bins.view(DspArray), center_freqs.squeeze().view(DspArray)


def _kaldi_window(win_type: str, size: int) -> np.ndarray:
    # float32 array expressions of dsp.py:634-648 (NOT the float64 windows of hanning()/hamming())
    f32 = np.float32
    n = np.arange(size, dtype=np.int32).astype(f32)
    arg = (f32(2.0) * f32(np.pi) * n / f32(size - 1)).astype(f32)
    if win_type == "hamming":
        return (f32(0.54) - f32(0.46) * np.cos(arg)).astype(f32)
    if win_type == "hanning":
        return (f32(0.5) - f32(0.5) * np.cos(arg)).astype(f32)
    if win_type == "povey":
        hann = (f32(0.5) - f32(0.5) * np.cos(arg)).astype(f32)
        return np.power(hann, f32(0.85)).astype(f32)
    return np.ones(size, f32)


def compute_fbank_kaldi(waveform, sample_rate: int = 48000, win_len: int = 1920, win_inc: int = 384,
                        num_mels: int = 60, win_type: str = "hamming", preemphasis: float = 0.97, dither: float = 1.0,
                        snip_edges: bool = True, low_freq: float = 20.0, high_freq: float = 0.0, *, seed: int = 0):
    """Kaldi-compatible log mel filterbank features (time, num_mels); reference dsp.py:577-676.  Framing, dither,
    per-frame DC removal, per-frame pre-emphasis, window, zero-extension to the next power of two, real FFT, power,
    mel projection and ln(max(., 1e-8)) run in ONE kernel.  `dither` != 0 draws N(0,1) per frame element from a
    Philox stream keyed by `seed` (the reference draws from MLX's global generator: equal in distribution only)."""
    ing = ingest(waveform, "float32")
    if ing.data.ndim == 2:
        ing.data = ing.data[0]  # dsp.py:607-608
    frame_length_ms = win_len / sample_rate * 1000
    frame_shift_ms = win_inc / sample_rate * 1000
    shift = int(sample_rate * frame_shift_ms * 0.001)
    size = int(sample_rate * frame_length_ms * 0.001)
    n_fft = _next_power_of_2(size)
    num_samples = int(ing.data.shape[0])
    x = ing.data
    if snip_edges:  # dsp.py:507-510
        m = 0 if num_samples < size else 1 + (num_samples - size) // shift
    else:  # dsp.py:511-521: reflect on the left without the edge sample, on the right INCLUDING it
        m = (num_samples + shift // 2) // shift
        pad = size // 2 - shift // 2
        if ing.on_device:
            import torch

            flip = lambda v: torch.flip(v, dims=(0,))
            cat = torch.cat
        else:
            flip = lambda v: v[::-1]
            cat = np.concatenate
        if pad > 0:
            left = flip(x[1 : pad + 1])
            # waveform[-1 : -pad - 1 : -1]: a signal shorter than `pad` contributes ALL of its samples, reversed
            right = flip(x[max(num_samples - pad, 0) :]) if pad > 1 else flip(x[1:])
            x = cat([left, x, right])
        else:
            x = cat([x[-pad:], flip(x)])
    if m <= 0:
        return emit(ing, np.zeros((0, num_mels), np.float32)) if not ing.on_device else _empty_like_device(ing, num_mels)
    bins, _ = get_mel_banks_kaldi(num_mels, n_fft, float(sample_rate), low_freq, high_freq)
    fb = np.pad(np.asarray(bins), [(0, 0), (0, 1)])  # dsp.py:668
    plan = cached_plan(FrontendPlan, _device_index(ing), _kaldi_window(win_type, size), fb, n_fft=n_fft, hop=shift,
                       center=False, spec_kind=_L.SPEC_POWER, log_kind=_L.LOG_LN, guard_kind=_L.GUARD_MAX,
                       guard_eps=1e-8, frame_len=size, frame_dc=True, frame_preemph=float(preemphasis),
                       dither=float(dither))
    ing.data = x.reshape(1, -1) if not ing.on_device else x.reshape(1, -1).contiguous()
    length = int(ing.data.shape[1]) + (n_fft - size)  # the zero extension of the LAST frame is virtual padding
    out = plan.run(ing, length=length, frame_count=m, seed=seed)
    return emit(ing, out[0])


def _empty_like_device(ing, num_mels):
    import torch

    return torch.zeros((0, num_mels), dtype=torch.float32, device=ing.device)
