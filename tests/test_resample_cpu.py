"""CPU: the resampler's host logic — filter design and index arithmetic — pinned against scipy.signal.resample_poly,
the dependency the reference calls (mlx_audio/stt/utils.py:27)."""
import os
import sys
from math import gcd

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pre_oracle as P  # noqa: E402

RATES = [(44100, 16000), (48000, 16000), (8000, 16000), (22050, 16000), (24000, 16000), (16000, 24000), (44100, 48000),
         (32000, 16000), (11025, 16000)]


@pytest.mark.parametrize("orig,target", RATES)
@pytest.mark.parametrize("n", [1, 7, 1000, 12345])
def test_restatement_matches_scipy(orig, target, n):
    from scipy import signal

    x = np.random.default_rng(n).standard_normal(n)
    g = gcd(orig, target)
    ref = signal.resample_poly(x, target // g, orig // g, padtype="edge")
    y = P.resample_poly_restated(x, target // g, orig // g)
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("orig,target", RATES)
def test_product_filter_design_matches_scipy(orig, target):
    """the product's NumPy filter design (no scipy import) == scipy.signal.firwin(...) * up, polyphase split"""
    from scipy import signal

    from mlx_audio_plus_b200.stt.utils import resample_poly_design

    g = gcd(orig, target)
    up, down, taps, J, pre = resample_poly_design(target // g, orig // g)
    assert (up, down) == (target // g, orig // g) and taps.shape == (J, up) and taps.dtype == np.float32
    max_rate = max(up, down)
    h = signal.firwin(2 * 10 * max_rate + 1, 1.0 / max_rate, window=("kaiser", 5.0)) * up
    n_pre_pad = down - (10 * max_rate) % down
    assert pre == (10 * max_rate + n_pre_pad) // down
    hp = np.concatenate([np.zeros(n_pre_pad), h])
    flat = taps.reshape(-1)
    assert np.abs(flat[: len(hp)] - hp.astype(np.float32)).max() <= 1e-9  # float64 design rounded once
    assert not flat[len(hp):].any()


def test_load_audio_oracle_follows_reference_steps():
    rng = np.random.default_rng(0)
    pcm = rng.integers(-30000, 30000, size=(4410, 2), dtype=np.int16)
    y = P.load_audio_from_pcm(pcm, 44100, 16000)
    assert y.dtype == np.float32 and y.shape == (1600,)
    same = P.load_audio_from_pcm(pcm, 16000, 16000)  # no resampling: conversion + mean only
    np.testing.assert_array_equal(same, ((pcm[:, 0].astype(np.float64) / 32768).astype(np.float32)
                                         + (pcm[:, 1].astype(np.float64) / 32768).astype(np.float32)) / np.float32(2))


def _write_wav(path, pcm, rate, width=2):
    import wave

    with wave.open(str(path), "wb") as w:
        w.setnchannels(1 if pcm.ndim == 1 else pcm.shape[1])
        w.setsampwidth(width)
        w.setframerate(rate)
        w.writeframes(pcm.astype("<i2").tobytes() if width == 2 else bytes(pcm.size * width))


def test_load_audio_needs_decoded_pcm_unless_pcm16_wave(tmp_path):
    """Only the codec-free container is parsed on the host (16-bit PCM WAVE -> the interleaved int16 audio_io.read yields,
    audio_io.py:250-262); anything else must arrive decoded (pcm= / sample_rate=)."""
    from mlx_audio_plus_b200.stt.utils import load_audio, read_wav_pcm16

    rng = np.random.default_rng(3)
    mono = rng.integers(-32768, 32767, 1234, dtype=np.int16)
    stereo = rng.integers(-32768, 32767, (777, 2), dtype=np.int16)
    _write_wav(tmp_path / "m.wav", mono, 16000)
    _write_wav(tmp_path / "s.wav", stereo, 44100)
    pcm, sr = read_wav_pcm16(str(tmp_path / "m.wav"))
    assert sr == 16000 and pcm.dtype == np.int16 and np.array_equal(pcm, mono)
    pcm, sr = read_wav_pcm16(str(tmp_path / "s.wav"))
    assert sr == 44100 and pcm.shape == (777, 2) and np.array_equal(pcm, stereo)
    (tmp_path / "clip.mp3").write_bytes(b"ID3\x03\x00" + bytes(64))
    with pytest.raises(NotImplementedError):
        load_audio(str(tmp_path / "clip.mp3"))
    _write_wav(tmp_path / "w24.wav", mono, 16000, width=3)
    with pytest.raises(NotImplementedError):
        load_audio(str(tmp_path / "w24.wav"))
    with pytest.raises(FileNotFoundError):
        load_audio(str(tmp_path / "missing.wav"))
    with pytest.raises(NotImplementedError):
        load_audio(None)
