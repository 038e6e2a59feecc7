"""ORACLE — TEST INFRASTRUCTURE ONLY (see dsp_oracle.py header).

NumPy float32 restatement of the per-model feature front-ends / iSTFT heads that sit on
the hot path.  Each function cites the reference lines it
follows.  All of them are compositions of dsp_oracle.{stft, istft, mel_filters}.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import dsp_oracle as D

F32 = np.float32


def _rpad(x, n, value=0.0):
    x = np.asarray(x, dtype=F32)
    return np.concatenate([x, np.full(n, value, dtype=F32)]) if n > 0 else x


# -- Whisper: stt/models/whisper/audio.py:44-85 -------------------------------------------
def whisper_log_mel(audio, n_mels=80, padding=0):
    x = _rpad(audio, padding)  # :73-74
    spec = D.stft(x, window=D.hanning(400), n_fft=400, hop_length=160)  # :75-76
    power = np.square(np.abs(spec[:-1, :])).astype(F32)  # :77 drop last frame; abs() then square()
    fb = D.mel_filters(16000, 400, n_mels, norm="slaney", mel_scale=None)  # :79
    mel = power @ fb.T  # :80
    y = np.log10(np.maximum(mel, F32(1e-10)))  # :82
    y = np.maximum(y, y.max() - F32(8.0))  # :83 one max over the whole input
    return ((y + F32(4.0)) / F32(4.0)).astype(F32)  # :84  -> (T, n_mels)


def whisper_pad_or_trim(a, length=480000, axis=-1):  # whisper/audio.py:27-41
    a = np.asarray(a)
    if a.shape[axis] > length:
        a = np.take(a, np.arange(length), axis=axis)
    if a.shape[axis] < length:
        pw = [(0, 0)] * a.ndim
        pw[axis] = (0, length - a.shape[axis])
        a = np.pad(a, pw)
    return a


# -- Parakeet / NeMo: stt/models/parakeet/audio.py:16-78 ----------------------------------
@dataclass
class PreprocessArgs:  # parakeet/audio.py:16-36
    sample_rate: int
    normalize: str
    window_size: float
    window_stride: float
    window: str
    features: int
    n_fft: int
    dither: float
    pad_to: int = 0
    pad_value: float = 0
    preemph: float = 0.97

    @property
    def win_length(self):
        return int(self.window_size * self.sample_rate)

    @property
    def hop_length(self):
        return int(self.window_stride * self.sample_rate)


def parakeet_log_mel(x, args: PreprocessArgs):
    x = np.asarray(x, dtype=F32)
    if args.pad_to > 0 and x.shape[-1] < args.pad_to:  # :42-45
        x = _rpad(x, args.pad_to - x.shape[-1], args.pad_value)
    fn = D.STR_TO_WINDOW_FN.get(args.window, None)  # :47-48 (no .lower() here)
    w = fn(args.win_length) if fn else D.hanning(args.win_length)
    if args.preemph > 0:  # :53-55
        x = np.concatenate([x[:1], x[1:] - F32(args.preemph) * x[:-1]]).astype(F32)
    spec = D.stft(x, args.n_fft, args.hop_length, args.win_length, w)  # :57
    power = np.square(np.abs(spec)).astype(F32)  # :58
    fb = D.mel_filters(args.sample_rate, args.n_fft, args.features, norm=args.normalize,
                       mel_scale=None)  # :59-61 — norm="per_feature" => NO slaney normalisation
    m = fb @ power.T  # :62  (M, T)
    m = np.log(m + F32(1e-5))  # :64
    if args.normalize == "per_feature":  # :66-69, std with ddof=0
        mean = m.mean(axis=1, keepdims=True, dtype=F32)
        std = m.std(axis=1, keepdims=True, dtype=F32)
    else:  # :70-73
        mean = m.mean(dtype=F32)
        std = m.std(dtype=F32)
    out = (m - mean) / (std + F32(1e-5))
    return out.T[None].astype(F32)  # :75-78 -> (1, T, M)


# -- Voxtral-Realtime: stt/models/voxtral_realtime/audio.py:19-96 ------------------------
def voxtral_rt_mel_filters(num_mel_bins=128, window_size=400, sample_rate=16000):  # :19-38
    fb = D.mel_filters(sample_rate, window_size, num_mel_bins, 0, 8000, "slaney", "slaney")
    return np.array(fb).T  # (F, M)


def voxtral_rt_mel(audio, mel_filters_fm, window_size=400, hop_length=160, global_log_mel_max=1.5):
    n = np.arange(window_size, dtype=F32)  # :60-61 periodic Hann evaluated in float32
    w = (F32(0.5) * (F32(1.0) - np.cos(F32(2.0 * math.pi) * n / F32(window_size)))).astype(F32)
    p = window_size // 2
    xp = np.pad(np.asarray(audio, dtype=F32), (p, p), mode="reflect")  # :64-67
    T = 1 + (xp.shape[0] - window_size) // hop_length  # :70-71
    idx = np.arange(window_size)[None, :] + (np.arange(T) * hop_length)[:, None]  # :74-77
    spec = np.fft.rfft(xp[idx] * w[None, :], n=window_size, axis=-1).astype(np.complex64)  # :80
    mags = (np.abs(spec) ** 2).astype(F32)[:-1, :].T  # :83-84 (F, T-1)
    mel = np.asarray(mel_filters_fm, F32).T @ mags  # :87
    y = np.log10(np.maximum(mel, F32(1e-10)))  # :90
    y = np.maximum(y, F32(global_log_mel_max - 8.0))  # :91-92 fixed floor
    return ((y + F32(4.0)) / F32(4.0)).astype(F32)  # :93 -> (M, T-1)


# -- Vocos: codec/models/vocos/mel.py:8-33 and vocos.py:126-140 --------------------------
def vocos_log_mel(audio, sample_rate=24000, n_mels=100, n_fft=1024, hop_length=256, padding=0):
    x = _rpad(audio, padding)
    # NB hop_length is NOT forwarded: win_length=hop_length is ignored for an array window and
    # hop defaults to n_fft//4 (mel.py:22)
    spec = D.stft(x, window=D.hanning(n_fft), n_fft=n_fft, win_length=hop_length)
    mag = np.abs(spec[:-1, :]).astype(F32)  # :23 magnitude, last frame dropped
    fb = D.mel_filters(sample_rate=sample_rate, n_fft=n_fft, n_mels=n_mels, norm=None, mel_scale="htk")
    mel = mag @ fb.T
    return np.log(np.maximum(mel, F32(1e-5)))[None].astype(F32)  # :31-33 -> (1, T, M)


def vocos_istft_head(x_lin, n_fft, hop_length):
    """x_lin: output of the head's linear layer, shape (1, T, n_fft+2) (vocos.py:127).
    Returns the waveform; follows vocos.py:127-140."""
    x = np.swapaxes(np.asarray(x_lin, F32), 1, 2)  # (1, n_fft+2, T)
    mag, p = np.split(x, 2, axis=1)
    mag = np.minimum(np.exp(mag), F32(1e2))  # :129-130
    S = (mag * (np.cos(p) + 1j * np.sin(p))).astype(np.complex64)  # :131-133
    return D.istft(S[0], window=D.hanning(n_fft), hop_length=hop_length, win_length=n_fft)  # :134-139


# -- Kokoro iSTFTNet: tts/models/kokoro/istftnet.py:399-528 ------------------------------
def kokoro_unwrap(p, axis=-1, period=2 * math.pi):  # mlx_unwrap :418-452 (discont=None)
    p = np.asarray(p, F32)
    discont = period / 2
    dd = np.diff(p, axis=axis).astype(F32)
    hi = period / 2
    lo = -hi
    ddmod = (dd - F32(period) * np.floor((dd - F32(lo)) / F32(period))).astype(F32)
    ddmod = np.where((np.abs(dd - F32(hi)) < 1e-10) & (dd > 0), F32(hi), ddmod)
    corr = (ddmod - dd).astype(F32)
    corr = np.where(np.abs(dd) < discont, F32(0), corr)
    shape = list(corr.shape)
    shape[axis] = 1
    corr = np.concatenate([np.zeros(shape, F32), corr], axis=axis)
    return (p + np.cumsum(corr, axis=axis, dtype=F32)).astype(F32)


def kokoro_transform(x, n_fft=20, hop=5, win=20, window="hann"):  # MLXSTFT.transform :464-495
    x = np.asarray(x, F32)
    if x.ndim == 1:
        x = x[None, :]
    mags, phases = [], []
    for row in x:
        s = D.stft(row, n_fft=n_fft, hop_length=hop, win_length=win, window=window, center=True,
                   pad_mode="reflect").T  # (F, T)
        mags.append(np.abs(s).astype(F32))
        phases.append(np.arctan2(s.imag, s.real).astype(F32))  # mlx_angle :399-415
    return np.stack(mags), np.stack(phases)


def kokoro_inverse(magnitude, phase, hop=5, win=20, window="hann"):  # MLXSTFT.inverse :497-523
    outs = []
    for m, ph in zip(np.asarray(magnitude, F32), np.asarray(phase, F32)):
        pc = kokoro_unwrap(ph, axis=1)
        spec = (m * np.cos(pc) + 1j * (m * np.sin(pc))).astype(np.complex64)
        outs.append(D.istft(spec, hop_length=hop, win_length=win, window=window, center=True, length=None))
    return np.stack(outs)[:, None, :]


# -- Qwen3-TTS mel (the golden-tested wrapper): tts/models/qwen3_tts/qwen3_tts.py:33-90 ---
def qwen3_tts_mel(audio, n_fft=1024, num_mels=128, sample_rate=24000, hop_size=256, win_size=1024,
                  fmin=0.0, fmax=12000.0):
    a = np.asarray(audio, F32)
    if a.ndim == 1:
        a = a[None, :]
    fb = D.mel_filters(sample_rate, n_fft, num_mels, fmin, fmax, "slaney", "slaney")
    pad = (n_fft - hop_size) // 2
    out = []
    for s in a:
        s = np.concatenate([s[1 : pad + 1][::-1], s, s[-(pad + 1) : -1][::-1]])  # :71-73
        spec = D.stft(s, n_fft=n_fft, hop_length=hop_size, win_length=win_size, window="hann",
                      center=False, pad_mode="reflect")
        mag = np.sqrt(np.abs(spec) ** 2 + F32(1e-9)).astype(F32)  # :85
        mel = mag @ fb.T
        out.append(np.log(np.clip(mel, F32(1e-5), None)).astype(F32))  # :89
    return np.stack(out)  # (B, T, M)


# -- S3Tokenizer: codec/models/s3tokenizer/utils.py:13-135 -------------------------------
def s3tokenizer_log_mel(audio, sample_rate=16000, n_mels=128, n_fft=400, hop_length=160, padding=0):
    x = _rpad(audio, padding)
    w = D.hanning(n_fft + 1)[:-1]  # :46 periodic via N+1 trick
    spec = D.stft(x, window=w, n_fft=n_fft, hop_length=hop_length, win_length=n_fft).swapaxes(0, 1)
    power = (np.abs(spec) ** 2).astype(F32)  # (F, T) — no frame drop
    fb = D.mel_filters(sample_rate=sample_rate, n_fft=n_fft, n_mels=n_mels, norm="slaney", mel_scale="slaney")
    mel = fb @ power
    y = np.log10(np.maximum(mel, F32(1e-10)))
    y = np.maximum(y, y.max() - F32(8.0))
    return ((y + F32(4.0)) / F32(4.0)).astype(F32)  # (M, T)


def s3tokenizer_log_mel_compat(audio, n_mels=128, padding=0):  # :68-135
    a = np.asarray(audio, F32)
    was_1d = a.ndim == 1
    if was_1d:
        a = a[None]
    if padding > 0:
        a = np.pad(a, [(0, 0), (0, padding)])
    spec = np.stack([D.stft(r, window="hann", n_fft=400, hop_length=160, win_length=400) for r in a])
    power = (np.abs(spec[:, :-1, :]) ** 2).astype(F32)
    fb = D.mel_filters(sample_rate=16000, n_fft=400, n_mels=n_mels, norm="slaney", mel_scale="slaney")
    mel = np.transpose(power @ fb.T, [0, 2, 1])  # (B, M, T)
    y = np.log10(np.maximum(mel, F32(1e-10)))
    y = np.maximum(y, y.max() - F32(8.0))  # ONE max over the whole batch
    y = ((y + F32(4.0)) / F32(4.0)).astype(F32)
    return y[0] if was_1d else y


# -- Sortformer / NeMo (constant pad, centred window, Bessel std): vad/models/sortformer/sortformer.py:36-120
def sortformer_mel(waveform, sample_rate=16000, n_fft=512, hop_length=160, win_length=400, n_mels=80,
                   preemphasis_coeff=0.97, normalize="per_feature", pad_to=16):
    wv = np.asarray(waveform, F32)
    if wv.ndim == 1:
        wv = wv[None, :]
    wv = np.concatenate([wv[..., :1], wv[..., 1:] - F32(preemphasis_coeff) * wv[..., :-1]], axis=-1).astype(F32)
    fb = D.mel_filters(sample_rate, n_fft, n_mels, 0, None, "slaney", "slaney")
    w = D.hanning(win_length)
    if win_length < n_fft:  # caller-side CENTRE padding of the window (:78-83)
        left = (n_fft - win_length) // 2
        w = np.concatenate([np.zeros(left, F32), w, np.zeros(n_fft - win_length - left, F32)])
    feats = []
    for row in wv:
        spec = D.stft(row, n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=w,
                      center=True, pad_mode="constant")
        power = (np.abs(spec) ** 2).astype(F32)
        feats.append(np.log(power @ fb.T + F32(2.0**-24)).T.astype(F32))  # _LOG_GUARD = 2**-24
    f = np.stack(feats)  # (B, M, T)
    if normalize == "per_feature":
        mean = f.mean(axis=2, keepdims=True, dtype=F32)
        var = ((f - mean) ** 2).sum(axis=2, keepdims=True, dtype=F32) / F32(f.shape[2] - 1)
        f = (f - mean) / (np.sqrt(var) + F32(1e-5))
    if pad_to > 0 and f.shape[2] % pad_to:
        f = np.pad(f, [(0, 0), (0, 0), (0, pad_to - f.shape[2] % pad_to)])
    return f.astype(F32)


# -- HiFT vocoder STFT pair (model-local, batched): codec/models/s3gen/hifigan.py:408-549 (reflect pad, clip <= 1e2)
#    and tts/models/cosyvoice3/hifigan.py:382-499 (zero pad, clip to [0, 1e2]); n_fft=16, hop=4 in HiFTGenerator
def hift_stft(x, n_fft, hop, window, pad_mode="reflect"):
    x = np.asarray(x, F32)
    window = np.asarray(window, F32)
    p = n_fft // 2
    if pad_mode == "reflect":  # s3gen/hifigan.py:421-427
        xp = np.concatenate([x[:, 1 : p + 1][:, ::-1], x, x[:, -(p + 1) : -1][:, ::-1]], axis=1)
    else:  # cosyvoice3/hifigan.py:399-400
        xp = np.pad(x, [(0, 0), (p, p)])
    n_frames = (xp.shape[1] - n_fft) // hop + 1
    idx = np.arange(n_frames)[:, None] * hop + np.arange(n_fft)[None, :]
    frames = (xp[:, idx] * window).astype(F32)  # (B, frames, n_fft)
    spec = np.fft.rfft(frames, axis=-1).astype(np.complex64)  # fft[:n_fft//2+1] == rfft (s3gen :449-453)
    spec = np.swapaxes(spec, 1, 2)
    return np.ascontiguousarray(spec.real), np.ascontiguousarray(spec.imag)


def hift_istft(magnitude, phase, n_fft, hop, window, clip_min_zero=False):
    m = np.asarray(magnitude, F32)
    ph = np.asarray(phase, F32)
    window = np.asarray(window, F32)
    m = np.clip(m, 0.0 if clip_min_zero else None, F32(1e2))  # s3gen :481 / cosyvoice3 :447
    re, im = (m * np.cos(ph)).astype(F32), (m * np.sin(ph)).astype(F32)
    spec = np.swapaxes(re + 1j * im, 1, 2).astype(np.complex64)  # (B, frames, F)
    # real(ifft(Hermitian-extended spectrum)) == irfft: Im(DC), Im(Nyquist) drop out (s3gen :491-502)
    frames = (np.fft.irfft(spec, n=n_fft, axis=-1).astype(F32) * window).astype(F32)
    B, n_frames, _ = frames.shape
    out_len = n_fft + (n_frames - 1) * hop
    idx = (np.arange(n_frames)[:, None] * hop + np.arange(n_fft)[None, :]).reshape(-1)
    wsum = np.zeros(out_len, F32)
    np.add.at(wsum, idx, np.tile((window * window).astype(F32), n_frames))
    wsum = np.maximum(wsum, F32(1e-8))  # s3gen :521 / cosyvoice3 :493
    out = np.zeros((B, out_len), F32)
    for b in range(B):
        np.add.at(out[b], idx, frames[b].reshape(-1))
    out = (out / wsum[None, :]).astype(F32)
    p = n_fft // 2
    return out[:, p:-p]


# -- Kaldi-compatible features: dsp.py:439-676 (compute_deltas_kaldi, mel banks, compute_fbank_kaldi with dither=0) --
def kaldi_deltas(specgram, win_length=5, mode="edge"):  # dsp.py:439-483
    x = np.asarray(specgram, F32)
    shape = x.shape
    x = x.reshape(-1, shape[-1])
    n = (win_length - 1) // 2
    denom = F32(float(n * (n + 1) * (2 * n + 1)) / 3.0)
    padded = np.pad(x, [(0, 0), (n, n)], mode="edge" if mode == "edge" else "constant")
    k = np.arange(-n, n + 1).astype(F32)
    out = np.zeros_like(x)
    for i in range(x.shape[1]):
        out[:, i] = (padded[:, i : i + win_length] * k).sum(axis=1, dtype=F32) / denom
    return out.reshape(shape)


def kaldi_mel_scale(f):  # dsp.py:486-488
    return (F32(1127.0) * np.log(F32(1.0) + np.asarray(f, F32) / F32(700.0))).astype(F32)


def kaldi_inverse_mel_scale(m):  # dsp.py:491-493
    return (F32(700.0) * (np.exp(np.asarray(m, F32) / F32(1127.0)) - F32(1.0))).astype(F32)


def kaldi_mel_banks(num_bins, n_fft, sample_freq, low_freq, high_freq):  # dsp.py:526-574
    nyq = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyq
    width = sample_freq / n_fft
    lo, hi = float(kaldi_mel_scale(low_freq)), float(kaldi_mel_scale(high_freq))
    d = (hi - lo) / (num_bins + 1)
    i = np.arange(num_bins).reshape(-1, 1).astype(F32)
    left, center, right = (F32(lo) + i * F32(d)), (F32(lo) + (i + F32(1)) * F32(d)), (F32(lo) + (i + F32(2)) * F32(d))
    mel = kaldi_mel_scale(F32(width) * np.arange(n_fft // 2).astype(F32)).reshape(1, -1)
    up, down = (mel - left) / (center - left), (right - mel) / (right - center)
    return np.maximum(F32(0), np.minimum(up, down)).astype(F32), kaldi_inverse_mel_scale(center).squeeze()


def kaldi_fbank(waveform, sample_rate=48000, win_len=1920, win_inc=384, num_mels=60, win_type="hamming",
                preemphasis=0.97, snip_edges=True, low_freq=20.0, high_freq=0.0):  # dsp.py:577-676, dither = 0
    x = np.asarray(waveform, F32)
    if x.ndim == 2:
        x = x[0]
    shift = int(sample_rate * (win_inc / sample_rate * 1000) * 0.001)
    size = int(sample_rate * (win_len / sample_rate * 1000) * 0.001)
    n_fft = 1 if size == 0 else 2 ** (size - 1).bit_length()
    n = x.shape[0]
    if snip_edges:  # dsp.py:507-510
        if n < size:
            return np.zeros((0, num_mels), F32)
        m = 1 + (n - size) // shift
    else:  # dsp.py:511-521
        m = (n + shift // 2) // shift
        pad = size // 2 - shift // 2
        if pad > 0:
            right = x[-1 : -pad - 1 : -1] if pad > 1 else x[-1:0:-1]
            x = np.concatenate([x[1 : pad + 1][::-1], x, right])
        else:
            x = np.concatenate([x[-pad:], x[::-1]])
    fr = np.lib.stride_tricks.as_strided(x, shape=(m, size), strides=(4 * shift, 4)).astype(F32)
    fr = fr - fr.mean(axis=1, keepdims=True, dtype=F32)  # dsp.py:624-626
    if preemphasis != 0.0:  # dsp.py:628-632
        fr = np.concatenate([fr[:, :1], fr[:, 1:] - F32(preemphasis) * fr[:, :-1]], axis=1).astype(F32)
    k = np.arange(size).astype(F32)
    arg = F32(2) * F32(np.pi) * k / F32(size - 1)
    if win_type == "hamming":
        w = F32(0.54) - F32(0.46) * np.cos(arg)
    elif win_type == "hanning":
        w = F32(0.5) - F32(0.5) * np.cos(arg)
    elif win_type == "povey":
        w = np.power(F32(0.5) - F32(0.5) * np.cos(arg), F32(0.85))
    else:
        w = np.ones(size, F32)
    fr = (fr * w.astype(F32)).astype(F32)
    spec = np.abs(np.fft.rfft(fr, n=n_fft, axis=1).astype(np.complex64)) ** F32(2.0)  # dsp.py:659-660
    bins, _ = kaldi_mel_banks(num_mels, n_fft, float(sample_rate), low_freq, high_freq)
    fb = np.pad(bins, [(0, 0), (0, 1)])
    return np.log(np.maximum(spec.astype(F32) @ fb.T, F32(1e-8))).astype(F32)


# ---- the steps right after the path -----------------------------------------------------------
def funasr_log_mel(audio, n_mels=80, n_fft=400, hop_length=160, sample_rate=16000):  # funasr/audio.py:32-81
    freqs = D.stft(np.asarray(audio, np.float32), n_fft, hop_length, window=D.hamming(n_fft))
    mags = np.square(np.abs(freqs[:-1, :])).astype(np.float32)
    fb = D.mel_filters(sample_rate, n_fft, n_mels, norm="slaney", mel_scale="htk")
    return np.log(np.maximum((mags @ fb.T).astype(np.float32), np.float32(1e-10))).astype(np.float32)


def funasr_apply_lfr(features, lfr_m=7, lfr_n=6):  # funasr/audio.py:84-139
    f = np.asarray(features, np.float32)
    T, n_mels = f.shape
    t_lfr = int(math.ceil(T / lfr_n))
    left = (lfr_m - 1) // 2
    if left > 0:
        f = np.concatenate([np.broadcast_to(f[0:1], (left, n_mels)), f], axis=0)
    need = (t_lfr - 1) * lfr_n + lfr_m
    if need > f.shape[0]:
        f = np.concatenate([f, np.broadcast_to(f[-1:], (need - f.shape[0], n_mels))], axis=0)
    idx = (np.arange(t_lfr) * lfr_n)[:, None] + np.arange(lfr_m)[None, :]
    return f[idx].reshape(t_lfr, -1)


def funasr_apply_cmvn(features, cmvn_mean=None, cmvn_istd=None):  # funasr/audio.py:142-169
    if cmvn_mean is None or cmvn_istd is None:  # per-utterance (160-164): mx.mean / mx.std (ddof 0) over axis 0, std + 1e-6
        f = np.asarray(features, np.float32)
        mean = f.mean(axis=0, keepdims=True, dtype=np.float32)
        std = f.std(axis=0, keepdims=True, dtype=np.float32) + np.float32(1e-6)
        return ((f - mean) / std).astype(np.float32)
    return ((np.asarray(features, np.float32) + np.asarray(cmvn_mean, np.float32)) * np.asarray(cmvn_istd, np.float32)).astype(np.float32)


def whisper_mel_segment(mel, seek, segment_size, n_frames=3000, dtype=np.float16):  # whisper/whisper.py:990-996
    seg = np.asarray(mel)[seek: seek + segment_size]
    return whisper_pad_or_trim(seg, n_frames, axis=-2).astype(dtype)


# -- the remaining parameter variants of SURVEY §8a row a12 ---------------------------------------------------------------
def glmasr_preprocess_audio(audio, n_mels=128):  # stt/models/glmasr/glmasr.py:547-589: Whisper-128 chain, (1, T, M)
    a = np.asarray(audio, F32)
    if a.ndim == 3:  # already features (:569-570)
        return a
    return whisper_log_mel(a, n_mels)[None]


def smart_turn_prepare_audio(audio, max_audio_seconds=8, sampling_rate=16000, normalize_audio=True):  # smart_turn.py:158-201
    a = np.asarray(audio, F32)
    max_samples = max_audio_seconds * sampling_rate
    if a.shape[0] > max_samples:
        a = a[-max_samples:]  # keeps the END of the turn
    elif a.shape[0] < max_samples:
        a = np.pad(a, (max_samples - a.shape[0], 0), mode="constant")  # LEFT zero padding
    if normalize_audio and a.size > 0:
        mean, std = float(a.mean()), float(a.std())
        a = (a - mean) / max(std, 1e-7)
    return a.astype(F32, copy=False)


def smart_turn_features(audio, n_mels=80, max_audio_seconds=8, sampling_rate=16000, hop_length=160, normalize_audio=True):
    """smart_turn.py:203-229 -> (n_mels, target_frames)"""
    mel = whisper_log_mel(smart_turn_prepare_audio(audio, max_audio_seconds, sampling_rate, normalize_audio), n_mels)
    target = max_audio_seconds * sampling_rate // hop_length
    if mel.shape[0] > target:
        mel = mel[-target:]
    elif mel.shape[0] < target:
        mel = np.pad(mel, [(target - mel.shape[0], 0), (0, 0)])
    return mel.T.astype(F32)


def s3gen_mel(y, n_fft=1920, num_mels=80, sampling_rate=24000, hop_size=480, win_size=1920, fmin=0, fmax=8000):
    """codec/models/s3gen/mel.py:25-100: manual reflect pad (n_fft-hop)/2, center=False, magnitude, ln(max 1e-5), (B, M, T)"""
    a = np.asarray(y, F32)
    if a.ndim == 1:
        a = a[None]
    pad = (n_fft - hop_size) // 2
    fb = D.mel_filters(sampling_rate, n_fft, num_mels, fmin, fmax, "slaney", "slaney")
    out = []
    for s in a:
        if pad:
            s = np.concatenate([s[1 : pad + 1][::-1], s, s[-(pad + 1) : -1][::-1]])
        spec = D.stft(s, window="hann", n_fft=n_fft, hop_length=hop_size, win_length=win_size, center=False)
        out.append((np.abs(spec).astype(F32) @ fb.T).T)
    return np.log(np.maximum(np.stack(out), F32(1e-5))).astype(F32)


def indextts_log_mel(audio, sample_rate=24000, n_mels=100, n_fft=1024, hop_length=256, padding=0):
    """tts/models/indextts/mel.py:6-37: symmetric "hann", hop forwarded, NO frame drop, magnitude, HTK, ln(max 1e-5), (1, T, M)"""
    spec = D.stft(_rpad(audio, padding), window="hann", n_fft=n_fft, hop_length=hop_length, win_length=n_fft)
    fb = D.mel_filters(sample_rate=sample_rate, n_fft=n_fft, n_mels=n_mels, norm=None, mel_scale="htk")
    return np.log(np.maximum(np.abs(spec).astype(F32) @ fb.T, F32(1e-5)))[None].astype(F32)


def spark_mel(audio, sample_rate=16000, n_mels=128, n_fft=1024, f_min=10, f_max=None, hop_length=320, win_length=640, padding=0):
    """tts/models/spark/bicodec.py:20-49: periodic Hann-640 right-padded to 1024 by dsp.stft, magnitude, no log, (1, T, M)"""
    w = D.hanning(win_length + 1)[:-1]
    spec = D.stft(_rpad(audio, padding), window=w, win_length=win_length, hop_length=hop_length, n_fft=n_fft)
    fb = D.mel_filters(sample_rate=sample_rate, n_fft=n_fft, n_mels=n_mels, f_min=f_min, f_max=f_max, norm="slaney",
                       mel_scale="slaney")
    return (np.abs(spec).astype(F32) @ fb.T)[None].astype(F32)


@dataclass
class VoiceEncConfig:  # tts/models/chatterbox/voice_encoder/config.py (the fields melspectrogram reads)
    num_mels: int = 40
    sample_rate: int = 16000
    n_fft: int = 400
    hop_size: int = 160
    win_size: int = 400
    fmax: int = 8000
    fmin: int = 0
    mel_power: float = 2.0
    mel_type: str = "amp"
    normalized_mels: bool = False
    stft_magnitude_min: float = 1e-4


def chatterbox_ve_melspectrogram(wav, hp: VoiceEncConfig):  # tts/models/chatterbox/voice_encoder/melspec.py:13-77
    a = np.asarray(wav, F32)
    was_1d = a.ndim == 1
    if was_1d:
        a = a[None]
    spec = np.stack([D.stft(r, window="hann", n_fft=hp.n_fft, hop_length=hp.hop_size, win_length=hp.win_size) for r in a])
    mag = np.abs(spec).astype(F32)
    if hp.mel_power != 1.0:
        mag = (mag ** F32(hp.mel_power)).astype(F32)
    fb = D.mel_filters(hp.sample_rate, hp.n_fft, hp.num_mels, hp.fmin, hp.fmax, "slaney", "slaney")
    mel = np.transpose(mag @ fb.T, [0, 2, 1])
    if hp.mel_type == "db":
        mel = F32(20) * np.log10(np.maximum(mel, F32(hp.stft_magnitude_min)))
    if hp.normalized_mels:
        min_level_db = 20 * math.log10(hp.stft_magnitude_min)
        mel = (mel - F32(min_level_db)) / F32(-min_level_db + 15)
    mel = mel.astype(F32)
    return mel[0] if was_1d else mel


def lfm2_preprocess(audio, sample_rate=16000, features=128, n_fft=512, window_size=0.025, window_stride=0.01, window="hann",
                    preemph=0.97, log=True, normalize="per_feature"):
    """sts/models/lfm_audio/processor.py:61-140 with dither = 0: pre-emphasis, CONSTANT centre padding, power, Slaney/slaney,
    ln(x + 5.96e-8), mean / Bessel std over the first len // hop frames applied to ALL frames; (B, T, M) or (T, M)"""
    a = np.asarray(audio, F32)
    single = a.ndim == 1
    if single:
        a = a[None]
    hop, win = int(sample_rate * window_stride), int(sample_rate * window_size)
    fb = D.mel_filters(sample_rate, n_fft, features, 0.0, sample_rate // 2, "slaney", "slaney")
    out = []
    for w in a:
        if preemph > 0:
            w = np.concatenate([w[:1], w[1:] - F32(preemph) * w[:-1]]).astype(F32)
        spec = D.stft(w, n_fft=n_fft, hop_length=hop, win_length=win, window=window, center=True, pad_mode="constant")
        mel = (np.abs(spec) ** 2).astype(F32) @ fb.T
        if log:
            mel = np.log(mel + F32(5.96e-8))
        if normalize == "per_feature":
            n = min(len(w) // hop, mel.shape[0])
            v = mel[:n]
            mean = v.mean(axis=0, keepdims=True, dtype=F32)
            var = ((v - mean) ** 2).sum(axis=0, keepdims=True, dtype=F32) / F32(n - 1)
            mel = (mel - mean) / (np.sqrt(var) + F32(1e-5))
        out.append(mel.astype(F32))
    f = np.stack(out)
    return f[0] if single else f


def lfm2_detokenizer_istft(mag, phase, n_fft=1280, hop_length=320, window=None):
    """sts/models/lfm_audio/detokenizer.py:468-507: per item istft(center=False, normalized=True), then the "same" trim of
    (n_fft - hop) / 2 samples at both ends; mag, phase: (B, T, F) -> (B, T * hop)"""
    mag, phase = np.asarray(mag, F32), np.asarray(phase, F32)
    S = (mag * np.cos(phase) + 1j * (mag * np.sin(phase))).astype(np.complex64)
    pad = (n_fft - hop_length) // 2
    out = []
    for s in S:
        y = D.istft(s.T, hop_length=hop_length, win_length=n_fft, window=window, center=False, normalized=True)
        out.append(y[pad:-pad] if pad > 0 else y)
    return np.stack(out).astype(F32)


def soprano_istft_head(x_lin, n_fft, hop_length):  # tts/models/soprano/decoder.py:22-49: the Vocos head, output (1, samples)
    return vocos_istft_head(x_lin, n_fft, hop_length)[None, :]


def mossformer2_chunk_stft(audio_segment, fft_len=1920, win_inc=384, win_len=1920, window=None):
    """sts/models/mossformer2_se/model.py:396-406: positional dsp.stft(center=False), returned as (F, T) real / imag planes"""
    s = D.stft(np.asarray(audio_segment, F32), fft_len, win_inc, win_len, window, center=False)
    return np.ascontiguousarray(s.real.T), np.ascontiguousarray(s.imag.T)


def mossformer2_chunk_istft(real, imag, fft_len=1920, win_inc=384, win_len=1920, window=None, chunk_length=None):
    """model.py:415-428: ISTFTCache.istft on a batch of one, center=False, audio_length=chunk_length -> (samples,)"""
    return D.ISTFTCache().istft(np.asarray(real, F32)[None], np.asarray(imag, F32)[None], fft_len, win_inc, win_len,
                                window, center=False, audio_length=chunk_length)[0]


def s3gen_xvector_fbank(audio, sample_rate=16000, num_mel_bins=80, frame_length=25.0, frame_shift=10.0):
    """codec/models/s3gen/xvector.py:38-150 (CAMPPlus front-end): snip-edges framing, per-frame DC removal and pre-emphasis
    0.97, float32 Povey window, zero-extension to the next power of two, power, dsp.mel_filters HTK from 20 Hz WITHOUT
    normalisation, ln(max(., float32 eps)); a signal shorter than one window gives one zero-extended frame."""
    x = np.asarray(audio, F32).squeeze()
    size, shift = int(sample_rate * frame_length / 1000), int(sample_rate * frame_shift / 1000)
    n_fft = 1 if size <= 1 else 1 << (size - 1).bit_length()
    m = (x.shape[0] - size) // shift + 1
    if m < 1:  # :77-78 then mx.take clamps nothing: the reference's gather of a short signal is restated as zero extension
        m = 1
        x = np.concatenate([x, np.zeros(size - x.shape[0], F32)])
    fr = np.lib.stride_tricks.as_strided(x, shape=(m, size), strides=(4 * shift, 4)).astype(F32)
    fr = fr - fr.mean(axis=1, keepdims=True, dtype=F32)
    fr = np.concatenate([fr[:, :1], fr[:, 1:] - F32(0.97) * fr[:, :-1]], axis=1).astype(F32)
    k = np.arange(size).astype(F32)
    w = np.power(F32(0.5) - F32(0.5) * np.cos(F32(2) * F32(np.pi) * k / F32(size - 1)), F32(0.85)).astype(F32)
    spec = np.abs(np.fft.rfft((fr * w).astype(F32), n=n_fft, axis=1).astype(np.complex64)) ** F32(2.0)
    fb = D.mel_filters(sample_rate=sample_rate, n_fft=n_fft, n_mels=num_mel_bins, f_min=20.0, f_max=sample_rate / 2, norm=None,
                       mel_scale="htk")
    return np.log(np.maximum(spec.astype(F32) @ fb.T, F32(1.1920929e-07))).astype(F32)


# -- Hugging Face WhisperFeatureExtractor, the third-party front-end of Qwen3-ASR / Qwen3-ForcedAligner --------------------------
#    (mlx_audio/stt/models/qwen3_asr/qwen3_asr.py:800-846; transformers/models/whisper/feature_extraction_whisper.py)
def hf_mel_filter_bank(num_frequency_bins, num_mel_filters, min_frequency, max_frequency, sampling_rate):
    """transformers.audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney"), float64, (bins, mels)"""
    def hz_to_mel(f):
        f = np.asarray(f, np.float64)
        return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * (27.0 / np.log(6.4)), 3.0 * f / 200.0)

    def mel_to_hz(m):
        m = np.asarray(m, np.float64)
        return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), 200.0 * m / 3.0)

    ff = mel_to_hz(np.linspace(hz_to_mel(min_frequency), hz_to_mel(max_frequency), num_mel_filters + 2))
    fft = np.linspace(0, sampling_rate // 2, num_frequency_bins)
    d = np.diff(ff)
    sl = ff[None, :] - fft[:, None]
    fb = np.maximum(0.0, np.minimum(-sl[:, :-2] / d[:-1], sl[:, 2:] / d[1:]))
    return fb * (2.0 / (ff[2 : num_mel_filters + 2] - ff[:num_mel_filters]))[None, :]


def hf_whisper_features(clips, feature_size=128, padding="longest", max_length=480000, truncation=False, do_normalize=False,
                        n_fft=400, hop=160, sampling_rate=16000):
    """clips: list of 1-D waveforms -> (features (B, M, T), attention_mask (B, T) int32).  Follows __call__ (pad /
    truncate / optional zero-mean-unit-variance), _torch_extract_fbank_features (periodic Hann, reflect-centred STFT, power,
    last frame dropped, log10(clamp 1e-10), PER-CLIP max - 8, (x + 4) / 4) and the mask rescale (::hop, minus one when ragged)."""
    clips = [np.asarray(c, F32) for c in clips]
    if truncation:
        clips = [c[:max_length] for c in clips]
    target = max(len(c) for c in clips) if padding == "longest" else max_length
    x = np.zeros((len(clips), target), F32)
    mask = np.zeros((len(clips), target), np.int32)
    for i, c in enumerate(clips):
        x[i, : len(c)], mask[i, : len(c)] = c, 1
    if do_normalize:
        for i, c in enumerate(clips):
            n = len(c)
            v = (x[i] - x[i, :n].mean()) / np.sqrt(x[i, :n].var() + 1e-7)
            v[n:] = 0.0
            x[i] = v
    fb = hf_mel_filter_bank(1 + n_fft // 2, feature_size, 0.0, 8000.0, sampling_rate).astype(F32)
    w = D.hanning(n_fft, True)  # torch.hann_window(n_fft): periodic
    out = []
    for row in x:
        p = (np.abs(D.stft(row, window=w, n_fft=n_fft, hop_length=hop)[:-1]) ** 2).astype(F32)
        y = np.log10(np.maximum(fb.T @ p.T, F32(1e-10)))
        y = np.maximum(y, y.max() - F32(8.0))
        out.append(((y + F32(4.0)) / F32(4.0)).astype(F32))
    m = mask[:, ::hop]
    if target % hop:
        m = m[:, :-1]
    return np.stack(out), m
