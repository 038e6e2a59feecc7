"""GPU: parity of the CUDA path (through the C ABI) against the oracle and the committed fixtures.

Tolerances (BASELINE.json north_star):
  framing / padding / indexing ........ bit-exact
  STFT complex ........................ rel-L2 <= 2e-6 and max-abs <= 1e-4 * max|X|
  log-mel features .................... max-abs <= 1e-4 (log units, after clamp/affine)
  Parakeet / Sortformer normalised .... max-abs <= 5e-4 (division by std amplifies)
  iSTFT waveform ...................... max-abs <= 1e-5 * peak
"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import dsp_oracle as O  # noqa: E402
from oracle import wrappers_oracle as W  # noqa: E402
from oracle.make_golden import synth  # noqa: E402

from test_oracle_golden import ISTFT_CASES, STFT_CASES, mkwin  # noqa: E402


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def host(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def assert_stft_close(y, ref):
    y = host(y)
    assert y.shape == ref.shape and y.dtype == np.complex64
    rel = np.linalg.norm(y - ref) / max(np.linalg.norm(ref), 1e-30)
    assert rel <= 2e-6, f"rel-L2 {rel}"
    assert np.abs(y - ref).max() <= 1e-4 * np.abs(ref).max()


def assert_wave_close(y, ref, tol=1e-5):
    y = host(y)
    assert y.shape == ref.shape
    if ref.size == 0:
        return
    fin = np.isfinite(ref)
    assert (np.isfinite(y) == fin).all()
    assert np.abs(y[fin] - ref[fin]).max() <= tol * np.abs(ref[fin]).max()


# ---- framing: bit-exact -------------------------------------------------------------------------------
@pytest.mark.parametrize("L,n_fft,hop,center,mode", [
    (4000, 400, 160, True, "reflect"), (4001, 400, 160, True, "constant"), (4159, 512, 160, True, "reflect"),
    (999, 20, 5, True, "reflect"), (5000, 1024, 256, False, "reflect"), (201, 400, 160, True, "reflect"),
    (200, 400, 160, True, "reflect"), (48000, 1920, 384, False, "reflect"), (333, 16, 4, True, "reflect")])
def test_framing_bit_exact(L, n_fft, hop, center, mode):
    from mlx_audio_plus_b200.frontend import FrontendPlan

    ramp = np.arange(L, dtype=np.float32)  # exact in fp32 (< 2^24)
    plan = FrontendPlan(n_fft=n_fft, hop=hop, window=np.ones(n_fft, np.float32), center=center, pad_mode=mode)
    fr = host(plan.dump_frames(dev(ramp[None])))[0]
    np.testing.assert_array_equal(fr, O.frames_of(ramp, n_fft, hop, center, mode))
    w = np.asarray(O.hanning(n_fft))
    plan2 = FrontendPlan(n_fft=n_fft, hop=hop, window=w, center=center, pad_mode=mode)
    frw = host(plan2.dump_frames(dev(ramp[None]), apply_window=True))[0]
    np.testing.assert_array_equal(frw, O.frames_of(ramp, n_fft, hop, center, mode) * w)


def test_preemphasis_and_virtual_padding_bit_exact():
    from mlx_audio_plus_b200.frontend import FrontendPlan

    x = synth(5, 3000)
    plan = FrontendPlan(n_fft=512, hop=160, window=np.ones(512, np.float32), preemph=0.97)
    fr = host(plan.dump_frames(dev(x[None]), length=4000))[0]
    xp = np.concatenate([x, np.zeros(1000, np.float32)])
    y = np.concatenate([xp[:1], xp[1:] - np.float32(0.97) * xp[:-1]]).astype(np.float32)
    ref = O.frames_of(y, 512, 160, True, "reflect")
    np.testing.assert_array_equal(fr, ref)  # mul and sub are rounded separately on the device too


# ---- stft ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(STFT_CASES))
@pytest.mark.parametrize("where", ["cuda", "numpy"])
def test_stft_parity(golden, name, where):
    from mlx_audio_plus_b200.dsp import stft

    g = golden("stft")
    n_fft, hop, win, wspec, center, pad_mode = STFT_CASES[name]
    x = g[f"stft|{name}|x"]
    ref = g[f"stft|{name}|y"]
    np.testing.assert_array_equal(ref, O.stft(x, n_fft, hop, win, mkwin(wspec), center, pad_mode))
    y = stft(dev(x) if where == "cuda" else x, n_fft, hop, win, mkwin(wspec), center, pad_mode)
    if where == "cuda":
        assert y.is_cuda and y.dtype == torch.complex64
    else:
        assert isinstance(y, np.ndarray) and hasattr(y, "abs")
    assert_stft_close(y, ref)


def test_stft_batch_matches_loop():
    from mlx_audio_plus_b200.dsp import stft

    xb = np.stack([synth(30 + i, 5000) * (0.5 + i / 7) for i in range(5)])
    yb = host(stft(dev(xb), 400, 160, window=O.hanning(400)))
    for i in range(5):
        assert_stft_close(yb[i], O.stft(xb[i], 400, 160, window=O.hanning(400)))


def test_stft_errors_on_gpu():
    from mlx_audio_plus_b200.dsp import stft

    with pytest.raises(ValueError, match="too short"):
        stft(dev(np.zeros(100, np.float32)), 400, center=False)
    with pytest.raises(ValueError):
        stft(dev(np.zeros(1000, np.float32)), 400, window=np.ones(500, np.float32))


@pytest.mark.parametrize("n_fft,hop,L", [(400, 160, 1600), (400, 160, 1601), (400, 160, 1759), (512, 160, 257),
                                         (1024, 256, 1024), (20, 5, 11), (16, 4, 9), (800, 200, 401), (14, 3, 100),
                                         (22, 11, 300), (360, 90, 2000), (2048, 512, 5000), (1280, 320, 3000)])
def test_stft_edge_lengths_and_sizes(n_fft, hop, L):
    from mlx_audio_plus_b200.dsp import stft

    x = synth(77, L)
    assert_stft_close(stft(dev(x), n_fft, hop), O.stft(x, n_fft, hop))


@pytest.mark.parametrize("n_fft,hop,win,pre", [(400, 160, 400, 0.0), (512, 160, 400, 0.0), (1024, 256, 1024, 0.0),
                                               (512, 160, 512, 0.97), (400, 160, 400, 0.97), (800, 200, 800, 0.0),
                                               (800, 200, 600, 0.97), (1024, 320, 640, 0.0)])
def test_fast_stft_kernel_parity(n_fft, hop, win, pre):
    """dsp.stft for the common sizes runs on fast_stft_* (register FFT, in-place spectrum tile, coalesced complex rows):
    same tolerance as the generic kernel, several clips, ragged last tile, reflect edges, optional pre-emphasis."""
    from mlx_audio_plus_b200._arrays import Ingested
    from mlx_audio_plus_b200.frontend import FrontendPlan

    w = np.asarray(O.hanning(win))
    plan = FrontendPlan(n_fft=n_fft, hop=hop, window=w, preemph=pre)
    assert plan.kernel_name == f"fast_stft_{n_fft}x{hop}"
    xb = np.stack([synth(50 + i, 40017) * (0.5 + i / 5) for i in range(3)])
    xd = dev(xb)
    y = host(plan.run(Ingested("torch", True, xd, None, xd.device)))
    for i in range(3):
        xi = xb[i]
        if pre:
            xi = np.concatenate([xi[:1], xi[1:] - np.float32(pre) * xi[:-1]]).astype(np.float32)
        assert_stft_close(y[i], O.stft(xi, n_fft, hop, win, w))


# ---- istft --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(ISTFT_CASES))
@pytest.mark.parametrize("where", ["cuda", "numpy"])
def test_istft_parity(golden, name, where):
    from mlx_audio_plus_b200.dsp import istft

    g = golden("istft")
    hop, win, wspec, center, length, normalized = ISTFT_CASES[name]
    x, ref = g[f"istft|{name}|x"], g[f"istft|{name}|y"]
    y = istft(dev(x) if where == "cuda" else x, hop, win, mkwin(wspec), center, length, normalized)
    assert_wave_close(y, ref)


@pytest.mark.parametrize("name,n_fft,hop,wkind,center,alen", [
    ("c64", 64, 16, "hamming", True, 700), ("c1920", 1920, 384, "hamming", True, None),
    ("c20nc", 20, 5, "hanning", False, None)])
def test_istft_cache_parity(golden, name, n_fft, hop, wkind, center, alen):
    from mlx_audio_plus_b200.dsp import ISTFTCache

    g = golden("istft")
    c = ISTFTCache()
    y = c.istft(dev(g[f"icache|{name}|re"]), dev(g[f"icache|{name}|im"]), n_fft, hop, n_fft,
                getattr(O, wkind)(n_fft, False), center, alen)
    assert_wave_close(y, g[f"icache|{name}|y"])


def test_istft_batch_and_roundtrip_property():
    """istft(stft(x)) with window^2 normalisation and matching windows is the identity away from the edges
    (size-independent property; the reference's default sum-w normalisation is NOT an identity, App. B.10)."""
    from mlx_audio_plus_b200.dsp import istft, stft

    x = np.stack([synth(40 + i, 24000, 24000) for i in range(3)])
    w = np.asarray(O.hanning(1025)[:-1])
    S = stft(dev(x), 1024, 256, window=w)  # (B, T, F)
    y = host(istft(S.swapaxes(1, 2).contiguous(), 256, 1024, window=w, normalized=True))
    n = y.shape[1]
    assert np.abs(y[:, 1024:-1024] - x[:, 1024 : n - 1024]).max() <= 2e-5


def _rand_spec(seed, F, T, batch=None):
    rng = np.random.default_rng(seed)
    shape = (F, T) if batch is None else (batch, F, T)
    mag = np.minimum(np.exp(rng.normal(0, 0.5, shape)), 1e2)
    ph = rng.uniform(-np.pi, np.pi, shape)
    return (mag * np.exp(1j * ph)).astype(np.complex64)  # Im(DC), Im(Nyquist) non-zero on purpose (App. B.12)


@pytest.mark.parametrize("T,window,center,length,normalized", [
    (468, ("hanning", 1024, False, None), True, None, False),     # Vocos head: symmetric array window, sum-w
    (37, "hann", True, None, True),                                # periodic string window, sum-w^2, T % 4 != 0
    (29, "hamming", False, None, False),                           # no centre trim; exactly one tile + 1 frame
    (5, ("hanning", 1024, False, None), True, 777, False),         # `length` keeps the centre pad, odd length
    (1, "hann", True, None, False),                                # a single frame: empty after the centre trim
    (1, "hann", False, None, False),                               # a single frame, untrimmed
    (3, "blackman", True, None, True),
    (120, ("hanning", 640, True, None), True, None, False),        # short window, right zero-extended
])
def test_fast_istft_1024_parity(T, window, center, length, normalized):
    """The fused 1024/256 inverse kernel (fast_inv.cu) against the oracle on every option of dsp.istft."""
    from mlx_audio_plus_b200.dsp import istft

    x = _rand_spec(900 + T, 513, T)
    w = mkwin(window)
    ref = O.istft(x, 256, 1024, w, center, length, normalized)
    y = istft(dev(x), 256, 1024, w, center, length, normalized)
    assert_wave_close(y, ref)


def test_fast_istft_1024_batch_planar_and_kernel_name():
    """ISTFTCache form (separate real / imag planes, sum-w^2, clamp guard, front-only trim) through the fused
    kernel, batched; also pins that the fused kernel is the one that ran."""
    from mlx_audio_plus_b200.dsp import ISTFTCache
    from mlx_audio_plus_b200 import frontend

    B, T = 5, 61
    x = _rand_spec(4242, 513, T, batch=B)
    w = np.asarray(O.hanning(1024, False))
    oc = O.ISTFTCache()
    c = ISTFTCache()
    for alen in (None, 9000):
        ref = oc.istft(x.real, x.imag, 1024, 256, 1024, w, True, alen)
        y = c.istft(dev(np.ascontiguousarray(x.real)), dev(np.ascontiguousarray(x.imag)), 1024, 256, 1024, w, True, alen)
        assert_wave_close(y, ref)
    names = {pl.kernel_name for pl in list(frontend._CACHE.values())
             if isinstance(pl, frontend.IstftPlan) and pl.n_fft == 1024 and pl.hop == 256}
    assert names == {"fast_istft_1024x256"}, names


# ---- HiFT model-local stft / istft ------------------------------------------------------
@pytest.mark.parametrize("n_fft,hop", [(16, 4), (20, 5)])
@pytest.mark.parametrize("name,mod", [("hift_s3gen", "mlx_audio_plus_b200.codec.models.s3gen.hifigan"),
                                      ("hift_cosy3", "mlx_audio_plus_b200.tts.models.cosyvoice3.hifigan")])
@pytest.mark.parametrize("where", ["cuda", "numpy"])
def test_hift_pair_parity(golden, n_fft, hop, name, mod, where):
    """Batched stft -> (real, imag) and the fused polar istft (clip, cos / sin, irfft, window, overlap-add,
    max(sum w^2, 1e-8)) against fixtures produced by the reference's own functions."""
    import importlib

    m = importlib.import_module(mod)
    g = golden("hift")
    put = dev if where == "cuda" else (lambda a: a)
    w, x = g[f"n{n_fft}|w"], g[f"n{n_fft}|x"]
    re, im = m.stft(put(x), n_fft, hop, w)
    ref_re, ref_im = g[f"n{n_fft}|{name}|re"], g[f"n{n_fft}|{name}|im"]
    assert tuple(re.shape) == ref_re.shape
    assert_stft_close((host(re) + 1j * host(im)).astype(np.complex64), (ref_re + 1j * ref_im).astype(np.complex64))
    y = m.istft(put(g[f"n{n_fft}|mag"]), put(g[f"n{n_fft}|phase"]), n_fft, hop, w)
    assert_wave_close(y, g[f"n{n_fft}|{name}|y"])


def test_polar_istft_1024_matches_complex_path():
    """The polar input form on the fused 1024/256 kernel == the complex form fed with clip(mag)*(cos, sin)."""
    from mlx_audio_plus_b200.dsp import istft, istft_polar

    rng = np.random.default_rng(5)
    mag = np.exp(rng.normal(0, 2.0, (2, 513, 40))).astype(np.float32)
    ph = rng.uniform(-20, 20, mag.shape).astype(np.float32)
    w = np.asarray(O.hanning(1024))
    mc = np.minimum(mag, np.float32(1e2))
    spec = (mc * np.cos(ph) + 1j * (mc * np.sin(ph))).astype(np.complex64)
    ref = np.stack([O.istft(spec[i], 256, 1024, w, True, None, False) for i in range(2)])
    y = istft_polar(dev(mag), dev(ph), 1024, 256, w, mag_clip_max=1e2)
    assert_wave_close(y, ref)


@pytest.mark.parametrize("n_fft,hop", [(20, 5), (16, 4)])
@pytest.mark.parametrize("phase_kind", ["unit", "one_outlier", "wide"])
def test_small_polar_istft_phase_ranges(n_fft, hop, phase_kind):
    """istft_small from magnitude / phase planes: phases inside [-1, 1] (what Kokoro's generator produces: sin(.) of a network
    output, istftnet.py) take the reduction-free sin / cos polynomials, a warp holding ONE larger phase falls back to the
    Cody-Waite path for all of its frames, and wide phases always do — all three against the float64-exact complex form."""
    from mlx_audio_plus_b200.dsp import istft_polar

    rng = np.random.default_rng(11)
    F, T = n_fft // 2 + 1, 1000
    mag = np.exp(rng.normal(0, 1.0, (3, F, T))).astype(np.float32)
    if phase_kind == "wide":
        ph = rng.uniform(-30, 30, mag.shape).astype(np.float32)
    else:
        ph = np.sin(rng.normal(0, 1.0, mag.shape)).astype(np.float32)
        ph[0, :, :40] = np.array([1.0, -1.0] * 20, np.float32)  # the end points of the polynomial range
        if phase_kind == "one_outlier":
            ph[1, 3, 500] = np.float32(1.0000001)
            ph[2, 0, 17] = np.float32(-7.5)
    w = np.asarray(O.hanning(n_fft + 1)[:-1])
    spec = (mag.astype(np.float64) * np.exp(1j * ph.astype(np.float64))).astype(np.complex64)
    ref = np.stack([O.istft(spec[i], hop, n_fft, w, True, None, False) for i in range(3)])
    y = istft_polar(dev(mag), dev(ph), n_fft, hop, w)
    assert_wave_close(y, ref)


@pytest.mark.parametrize("n_fft,hop,pre,offset", [(1280, 320, 0.0, 0), (1280, 320, 0.0, 1), (1920, 480, 0.97, 0), (1920, 384, 0.0, 3),
                                                  (2048, 512, 0.97, 2), (360, 90, 0.0, 1)])
def test_generic_forward_interior_copy_matches_oracle(n_fft, hop, pre, offset):
    """frontend_generic_kernel stages INTERIOR tiles with a linear copy (16-byte loads when the clip's first sample is 16-byte
    aligned, 4-byte loads otherwise, a two-load form with pre-emphasis) and only edge tiles through the per-sample padding logic:
    several clips whose storage starts `offset` floats into an allocation, against the oracle's stft."""
    from mlx_audio_plus_b200._arrays import Ingested
    from mlx_audio_plus_b200.frontend import FrontendPlan

    w = np.asarray(O.hanning(n_fft))
    plan = FrontendPlan(n_fft=n_fft, hop=hop, window=w, preemph=pre)
    assert not plan.kernel_name.startswith("fast_"), plan.kernel_name
    B, L = 3, 40 * hop + n_fft + 37
    xb = np.stack([synth(90 + i, L) * (0.5 + i / 3) for i in range(B)])
    base = torch.zeros(B * L + 8, dtype=torch.float32, device="cuda")
    base[offset:offset + B * L] = dev(xb).reshape(-1)
    xd = base[offset:offset + B * L].view(B, L)
    assert xd.is_contiguous() and xd.data_ptr() % 16 == (4 * offset) % 16
    y = host(plan.run(Ingested("torch", True, xd, None, xd.device)))
    for i in range(B):
        xi = xb[i]
        if pre:
            xi = np.concatenate([xi[:1], xi[1:] - np.float32(pre) * xi[:-1]]).astype(np.float32)
        assert_stft_close(y[i], O.stft(xi, n_fft, hop, n_fft, w))


# ---- Kaldi-compatible features (dsp.py:439-676) -------------------------------------------------
@pytest.mark.parametrize("where", ["cuda", "numpy"])
def test_kaldi_fbank_and_deltas_parity(golden, where):
    """compute_fbank_kaldi (one fused launch: framing, per-frame DC removal and pre-emphasis, float32 Kaldi window,
    zero-extension to 2^k, FFT, power, Kaldi mel banks, ln(max(., 1e-8))) and compute_deltas_kaldi against fixtures
    produced by the reference's own code.  2e-4 absolute on the log energies (values ~ 5..20)."""
    from mlx_audio_plus_b200 import dsp
    from test_oracle_golden import KALDI_CASES

    g = golden("kaldi")
    put = dev if where == "cuda" else (lambda a: a)
    for name, (xk, kw) in KALDI_CASES.items():
        y = host(dsp.compute_fbank_kaldi(put(g[xk]), dither=0.0, **kw))
        ref = g[f"fbank|{name}"]
        assert y.shape == ref.shape, name
        assert np.abs(y - ref).max() <= 2e-4, (name, np.abs(y - ref).max())
    assert tuple(dsp.compute_fbank_kaldi(put(g["x16"][:300]), sample_rate=16000, win_len=400, win_inc=160, num_mels=23,
                                         dither=0.0).shape) == (0, 23)
    f = g["fbank|moss"].T.copy()
    for key, x, kw in (("deltas|edge5", f, dict(win_length=5)), ("deltas|const9", f, dict(win_length=9, mode="constant")),
                       ("deltas|3d", f[:24].reshape(2, 12, -1), dict(win_length=3))):
        d = host(dsp.compute_deltas_kaldi(put(x), **kw))
        assert d.shape == g[key].shape and np.abs(d - g[key]).max() <= 1e-5 * np.abs(g[key]).max()


@pytest.mark.parametrize("where", ["cuda", "numpy"])
def test_kaldi_snip_edges_false_short_input(where):
    """snip_edges=False with a signal SHORTER than the reflect pad (dsp.py:511-521): `waveform[-1 : -pad - 1 : -1]` keeps the
    whole reversed signal (a negative start would keep only its tail).  180 samples, pad 192: the padded signal (539 samples)
    still holds the one 512-sample frame, so the reference's strided view stays inside its buffer."""
    from mlx_audio_plus_b200 import dsp

    x = (synth(403, 180, 16000) * 8000.0).astype(np.float32)
    kw = dict(sample_rate=16000, win_len=512, win_inc=128, num_mels=40, win_type="rectangular", preemphasis=0.5, snip_edges=False)
    ref = W.kaldi_fbank(x, **kw)
    y = host(dsp.compute_fbank_kaldi(dev(x) if where == "cuda" else x, dither=0.0, **kw))
    assert y.shape == ref.shape == (1, 40)
    assert np.abs(y - ref).max() <= 2e-4


def test_kaldi_dither_is_seeded_noise_of_the_right_size():
    """dither != 0 cannot match MLX's generator; it must be reproducible per seed, differ across seeds, and perturb
    the frames like N(0, dither^2): on silence the mean frame energy after the Hamming window is dither^2 * sum(w^2)
    (minus the removed DC), which the flat part of the mel output reflects."""
    from mlx_audio_plus_b200 import dsp

    x = dev(np.zeros(48000, np.float32))
    kw = dict(sample_rate=16000, win_len=400, win_inc=160, num_mels=40, win_type="hamming", preemphasis=0.0)
    a = host(dsp.compute_fbank_kaldi(x, dither=1.0, seed=11, **kw))
    b = host(dsp.compute_fbank_kaldi(x, dither=1.0, seed=11, **kw))
    c = host(dsp.compute_fbank_kaldi(x, dither=1.0, seed=12, **kw))
    z = host(dsp.compute_fbank_kaldi(x, dither=0.0, **kw))
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.allclose(z, np.log(1e-8))  # silence without dither: the floor everywhere
    # white noise of variance 1 through window w and a triangular mel filter with weights t_k: E[mel] = sum(w^2) * sum_k t_k
    w = dsp._kaldi_window("hamming", 400)
    fb = np.pad(np.asarray(dsp.get_mel_banks_kaldi(40, 512, 16000.0, 20.0, 0.0)[0]), [(0, 0), (0, 1)])
    expect = np.log((w.astype(np.float64) ** 2).sum() * fb.sum(axis=1))
    got = np.log(np.exp(a.astype(np.float64)).mean(axis=0))
    assert np.abs(got[5:] - expect[5:]).max() <= 0.15, np.abs(got - expect).max()  # low bins lose the removed DC


# ---- model front-ends ---------------------------------------------------------------------------------
def test_whisper_parity(golden):
    from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram

    g = golden("models")
    x = g["whisper|x"]
    for n_mels, key, pad in ((80, "whisper|80", 0), (128, "whisper|128", 0), (80, "whisper|80pad", 8000)):
        y = log_mel_spectrogram(dev(x), n_mels=n_mels, padding=pad)
        assert tuple(y.shape) == g[key].shape
        assert np.abs(host(y) - g[key]).max() <= 1e-4
    ys = host(log_mel_spectrogram(dev(g["whisper|sil|x"]), n_mels=80))  # clamp active
    assert np.abs(ys - g["whisper|sil|80"]).max() <= 1e-4
    yn = log_mel_spectrogram(x, n_mels=80)  # numpy in -> numpy out through the host entry point
    assert isinstance(yn, np.ndarray) and np.abs(yn - g["whisper|80"]).max() <= 1e-4
    y30 = log_mel_spectrogram(dev(np.zeros(16000, np.float32)), n_mels=80, padding=480000)
    assert tuple(y30.shape) == (3100, 80)  # stt/tests/test_models.py contract: (N_FRAMES+100, n_mels)


def test_whisper_all_silent_tiles_are_written_once_by_the_fixup():
    """Digital silence covering whole 32-frame tiles: the fused kernel does not store such a tile (every value is the
    epilogue's constant for zero power) and the clamp fix-up writes max(c, floor) there.  Against the oracle, and the
    output buffer is poisoned first so an unwritten element cannot pass.  Cases: silence in the middle, at both ends, a
    clip that is silent throughout (floor = c - 2: the constant itself survives), and a batch mixing them."""
    from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram

    n = 16000 * 12
    clips = []
    for i, (a, b) in enumerate(((40000, 120000), (0, 70000), (100000, n), (0, n), (5117, 5117 + 160 * 40))):
        x = synth(700 + i, n) * (0.3 + 0.2 * i)
        x[a:b] = 0.0
        clips.append(x)
    xb = np.stack(clips)
    torch.empty((5, 1200, 128), device="cuda").fill_(float("nan"))  # poison what the allocator hands out next
    torch.cuda.synchronize()
    yb = host(log_mel_spectrogram(dev(xb), n_mels=128))
    assert np.isfinite(yb).all()
    for i in range(5):
        ref = W.whisper_log_mel(xb[i], 128)
        assert np.abs(yb[i] - ref).max() <= 1e-4, i
        torch.empty((1200, 128), device="cuda").fill_(float("nan"))
        one = host(log_mel_spectrogram(dev(xb[i]), n_mels=128))
        np.testing.assert_array_equal(one, yb[i])
    assert np.all(yb[3] == yb[3][0, 0]) and abs(float(yb[3][0, 0]) - (-10.0 + 4.0) / 4.0) <= 1e-6  # silent clip: log10(1e-10)
    y80 = host(log_mel_spectrogram(dev(xb), n_mels=80))  # the other generated instances of the family
    for i in range(5):
        assert np.abs(y80[i] - W.whisper_log_mel(xb[i], 80)).max() <= 1e-4


def test_whisper_batch_per_clip_max():
    from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram

    xb = np.stack([synth(50 + i, 48000) * (0.5 + (i % 7) / 7) for i in range(9)])
    xb[3, 20000:] = 0
    yb = host(log_mel_spectrogram(dev(xb), n_mels=128))
    for i in range(9):
        assert np.abs(yb[i] - W.whisper_log_mel(xb[i], 128)).max() <= 1e-4


def test_parakeet_parity(golden):
    from mlx_audio_plus_b200.stt.models.parakeet.audio import PreprocessArgs, log_mel_spectrogram

    g = golden("models")
    pa = PreprocessArgs(16000, "per_feature", 0.025, 0.01, "hann", 80, 512, 1e-5)
    y = host(log_mel_spectrogram(dev(g["parakeet|x"]), pa))
    assert y.shape == g["parakeet|pf"].shape and np.abs(y - g["parakeet|pf"]).max() <= 5e-4
    pa2 = PreprocessArgs(16000, "all_features", 0.025, 0.01, "hamming", 64, 512, 0.0, pad_to=30000, pad_value=0.0, preemph=0.0)
    y2 = host(log_mel_spectrogram(dev(g["parakeet|x"]), pa2))
    assert np.abs(y2 - g["parakeet|global"]).max() <= 5e-4


def _parakeet_float64_truth(x, n_fft=512, hop=160, win=400, n_mels=80, preemph=0.97):
    """parakeet/audio.py:39-78 evaluated in float64 end to end (the filterbank keeps its float32 VALUES: it is data)."""
    from oracle import dsp_oracle as D

    x = np.asarray(x, np.float64)
    y = np.concatenate([x[:1], x[1:] - preemph * x[:-1]])
    fr = y[D.frame_indices(len(y), n_fft, hop, True, "reflect")]
    w = D._fit_window(D.hanning(win), n_fft).astype(np.float64)
    p = np.abs(np.fft.rfft(fr * w, axis=1)) ** 2
    fb = D.mel_filters(16000, n_fft, n_mels, norm="per_feature", mel_scale=None).astype(np.float64)
    m = np.log(p @ fb.T + 1e-5)
    return ((m - m.mean(0)) / (m.std(0) + 1e-5))[None]


def _parakeet_single_precision_chain(x, n_fft=512, hop=160, win=400, n_mels=80):
    """The same chain in float32 with a SINGLE-PRECISION FFT (scipy's pocketfft on float32 input — the arithmetic MLX's CPU
    backend performs; NumPy 2.x's np.fft.rfft would transform in float64 and round once)."""
    import scipy.fft
    from oracle import dsp_oracle as D

    f32 = np.float32
    x = np.asarray(x, f32)
    y = np.concatenate([x[:1], x[1:] - f32(0.97) * x[:-1]]).astype(f32)
    fr = y[D.frame_indices(len(y), n_fft, hop, True, "reflect")]
    w = D._fit_window(D.hanning(win), n_fft).astype(f32)
    s = scipy.fft.rfft((fr * w).astype(f32), axis=1)
    assert s.dtype == np.complex64
    p = (np.abs(s).astype(f32) ** 2).astype(f32)
    fb = D.mel_filters(16000, n_fft, n_mels, norm="per_feature", mel_scale=None)
    m = np.log((p @ fb.T).astype(f32) + f32(1e-5)).astype(f32)
    return ((m - m.mean(0, dtype=f32)) / (m.std(0, dtype=f32) + f32(1e-5)))[None].astype(f32)


@pytest.mark.parametrize("seed,n", [(0, 24000), (3, 160000), (9, 480000)])
def test_parakeet_error_against_float64_truth(golden, seed, n):
    """err(kernel, float64 truth) <= 2 x err(float32 reference arithmetic, float64 truth) — the kernel is a
    float32 implementation of the chain and may not be materially worse than the reference's own float32 arithmetic.
    "float32 reference arithmetic" = the chain with a single-precision FFT, as MLX computes it; the NumPy oracle (whose FFT
    is float64 rounded once, ~3x more accurate) is checked too, at the 5e-4 parity tolerance."""
    from mlx_audio_plus_b200.stt.models.parakeet.audio import PreprocessArgs, log_mel_spectrogram

    x = golden("models")["parakeet|x"] if seed == 0 else synth(seed, n) * 0.7
    pa = PreprocessArgs(16000, "per_feature", 0.025, 0.01, "hann", 80, 512, 1e-5)
    truth = _parakeet_float64_truth(x)
    ref32 = _parakeet_single_precision_chain(x).astype(np.float64)
    oracle = W.parakeet_log_mel(x, W.PreprocessArgs(16000, "per_feature", 0.025, 0.01, "hann", 80, 512, 1e-5)).astype(np.float64)
    y = host(log_mel_spectrogram(dev(x), pa)).astype(np.float64)
    assert y.shape == truth.shape == ref32.shape == oracle.shape
    e_kernel, e_ref32 = np.abs(y - truth), np.abs(ref32 - truth)
    assert e_kernel.max() <= 2.0 * e_ref32.max() + 1e-6, (e_kernel.max(), e_ref32.max())
    assert np.sqrt((e_kernel ** 2).mean()) <= 2.0 * np.sqrt((e_ref32 ** 2).mean()) + 1e-7
    assert np.abs(y - oracle).max() <= 5e-4


def test_parakeet_returns_the_input_dtype(golden):
    """parakeet/audio.py:40,78: the features come back in the dtype of the waveform (bfloat16 from parakeet.py:184,227)."""
    from mlx_audio_plus_b200.stt.models.parakeet.audio import PreprocessArgs, log_mel_spectrogram

    pa = PreprocessArgs(16000, "per_feature", 0.025, 0.01, "hann", 80, 512, 1e-5)
    x = dev(golden("models")["parakeet|x"])
    xb = x.to(torch.bfloat16)
    yb = log_mel_spectrogram(xb, pa)
    assert yb.dtype == torch.bfloat16 and tuple(yb.shape) == (1, 151, 80)
    assert torch.equal(yb, log_mel_spectrogram(xb.float(), pa).to(torch.bfloat16))  # float32 arithmetic, one final cast
    assert log_mel_spectrogram(x, pa).dtype == torch.float32


def test_other_frontends_parity(golden):
    from mlx_audio_plus_b200.codec.models.s3tokenizer.utils import log_mel_spectrogram as s3_mel
    from mlx_audio_plus_b200.codec.models.s3tokenizer.utils import log_mel_spectrogram_compat as s3_compat
    from mlx_audio_plus_b200.codec.models.vocos.mel import log_mel_spectrogram as vocos_mel
    from mlx_audio_plus_b200.codec.models.vocos.vocos import ISTFTHead
    from mlx_audio_plus_b200.stt.models.voxtral_realtime.audio import compute_mel_filters, compute_mel_spectrogram
    from mlx_audio_plus_b200.tts.models.qwen3_tts.qwen3_tts import mel_spectrogram
    from mlx_audio_plus_b200.vad.models.sortformer.sortformer import extract_mel_features

    g = golden("models")
    assert np.abs(host(compute_mel_spectrogram(dev(g["voxtral|x"]), compute_mel_filters())) - g["voxtral|y"]).max() <= 1e-4
    assert np.abs(host(vocos_mel(dev(g["vocos|x"]))) - g["vocos|mel"]).max() <= 1e-4
    assert_wave_close(ISTFTHead(8, 1024, 256)(dev(g["vocos|head_in"])), g["vocos|head_out"])
    assert tuple(vocos_mel(dev(np.zeros(120000, np.float32))).shape) == (1, 468, 100)
    assert tuple(ISTFTHead(8, 1024, 256)(dev(np.zeros((1, 468, 1026), np.float32))).shape) == (119552,)
    np.random.seed(42)
    xq = np.random.randn(12000).astype(np.float32)
    mq = host(mel_spectrogram(dev(xq)))
    assert np.abs(mq - g["qwen3|y"]).max() <= 1e-4
    np.testing.assert_allclose(mq[0][0, [0, 1, 2, 63, 126, 127]],
                               [-0.21803714, 0.06630915, -0.31858957, -0.02480409, -0.4512914, -0.5911693], rtol=1e-4, atol=1e-4)
    assert np.abs(host(s3_mel(dev(g["s3|x"]))) - g["s3|y"]).max() <= 1e-4
    assert np.abs(host(s3_compat(dev(g["s3|xb"]), 80)) - g["s3|compat"]).max() <= 1e-4
    assert np.abs(host(extract_mel_features(dev(g["sortformer|x"]))) - g["sortformer|y"]).max() <= 5e-4


def test_kokoro_parity(golden):
    from mlx_audio_plus_b200.tts.models.kokoro.istftnet import MLXSTFT

    g = golden("models")
    K = MLXSTFT(filter_length=20, hop_length=5, win_length=20)
    mag, ph = K.transform(dev(g["kokoro|x"]))
    assert np.abs(host(mag) - g["kokoro|mag"]).max() <= 1e-4 * np.abs(g["kokoro|mag"]).max()
    # phase of near-zero bins is ill-conditioned: compare where the magnitude is significant
    sig = g["kokoro|mag"] > 1e-3 * g["kokoro|mag"].max()
    d = np.abs(host(ph) - g["kokoro|phase"])[sig]
    assert np.minimum(d, 2 * math.pi - d).max() <= 1e-3
    y = K.inverse(dev(g["kokoro|inv_mag"]), dev(g["kokoro|inv_phase"]))
    assert_wave_close(y, g["kokoro|inv_y"])
    yw = K.inverse(dev(g["kokoro|inv_mag"][:1]), dev(g["kokoro|inv_phase_wrapped"]))
    assert_wave_close(yw, g["kokoro|inv_y_wrapped"], tol=1e-4)  # fp32 cumsum in unwrap is order dependent


# ---- BASELINE.json full sizes: size-independent properties + spot checks against the oracle ---------------------
def _bench_like_batch(B, n, sr, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    t = torch.arange(n, device="cuda", dtype=torch.float64) / sr
    tone = (0.2 * (torch.sin(2 * np.pi * 440 * t) + torch.sin(2 * np.pi * 3000 * t))).float()
    x = torch.empty((B, n), dtype=torch.float32, device="cuda")
    for c0 in range(0, B, 256):
        c1 = min(B, c0 + 256)
        scale = (0.5 + (torch.arange(c0, c1, device="cuda") % 7).float() / 7)[:, None]
        x[c0:c1] = (0.1 * torch.randn((c1 - c0, n), generator=g, device="cuda") + tone[None]) * scale
    return x


def test_c2_full_batch_4096x30s_properties():
    """BASELINE configs[1] at full size (4096 x 30 s, 128 mels): (a) sampled clips equal the oracle within 1e-4,
    (b) the batched launch is BIT-identical to running a clip alone (clips are independent: per-clip max, no
    cross-clip state), (c) the per-clip maximum maps to exactly (max+4)/4 and nothing lies below max - 2.0
    (the max-8 clamp after the /4 affine map)."""
    from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram

    free, _ = torch.cuda.mem_get_info()
    B = 4096 if free > 40e9 else 512
    x = _bench_like_batch(B, 480000, 16000, 1235)
    y = log_mel_spectrogram(x, n_mels=128)
    assert tuple(y.shape) == (B, 3000, 128)
    picks = [0, 1, 7, B // 2 + 3, B - 1]
    for i in picks:
        ref = W.whisper_log_mel(x[i].cpu().numpy(), 128)
        assert np.abs(y[i].cpu().numpy() - ref).max() <= 1e-4
        alone = log_mel_spectrogram(x[i : i + 1].clone(), n_mels=128)[0]
        assert torch.equal(alone, y[i])
    mx = y.amax(dim=(1, 2))
    mn = y.amin(dim=(1, 2))
    assert bool(((mx - mn) <= 2.0 + 1e-6).all())
    assert bool(torch.isfinite(y).all())


def test_c3_one_hour_file_full_size_vs_oracle():
    """BASELINE configs[2] at full size (one 1-hour file, Parakeet front-end: pre-emphasis, n_fft 512 / win 400 /
    hop 160, 80 mels, ln(x + 1e-5), per-feature normalisation over all 360 001 frames): the whole output against the
    oracle (<= 5e-4: the division by the per-mel std amplifies), plus the size-independent property that
    every mel row has mean 0 and std 1."""
    from mlx_audio_plus_b200.stt.models.parakeet.audio import PreprocessArgs, log_mel_spectrogram

    args = PreprocessArgs(sample_rate=16000, normalize="per_feature", window_size=0.025, window_stride=0.01,
                          window="hann", features=80, n_fft=512, dither=0.0)
    x = _bench_like_batch(1, 57_600_000, 16000, 1236)[0]
    y = log_mel_spectrogram(x, args)
    assert tuple(y.shape) == (1, 360001, 80)
    yd = y[0].double()
    assert float(yd.mean(0).abs().max()) <= 1e-4
    assert float((yd.std(0, unbiased=False) - 1).abs().max()) <= 1e-3
    ref = W.parakeet_log_mel(x.cpu().numpy(), W.PreprocessArgs(**{k: getattr(args, k) for k in (
        "sample_rate", "normalize", "window_size", "window_stride", "window", "features", "n_fft", "dither")}))
    assert ref.shape == (1, 360001, 80)
    assert np.abs(y.cpu().numpy() - ref).max() <= 5e-4


def test_c4_full_batch_kokoro_istft_1024_items():
    """BASELINE configs[3] at full size (B = 1024 x (11, 24 001) magnitude / phase, n_fft 20, hop 5): sampled items against the
    oracle's MLXSTFT.inverse (<= 1e-5 of the peak), the batched launch bit-identical to single-item launches, and exact
    linearity in the magnitude over the whole batch."""
    from mlx_audio_plus_b200.tts.models.kokoro.istftnet import MLXSTFT

    B, T = 1024, 24001
    g = torch.Generator(device="cuda")
    g.manual_seed(9)
    mag = torch.exp(0.5 * torch.randn((B, 11, T), generator=g, device="cuda")).clamp(max=1e2)
    ph = torch.sin(torch.randn((B, 11, T), generator=g, device="cuda"))  # Kokoro: phase = sin(.) (istftnet.py:805)
    st = MLXSTFT(filter_length=20, hop_length=5, win_length=20)
    y = st.inverse(mag, ph)
    assert tuple(y.shape) == (B, 1, 120000) and bool(torch.isfinite(y).all())
    for i in (0, 5, B // 2, B - 1):
        ref = W.kokoro_inverse(mag[i : i + 1].cpu().numpy(), ph[i : i + 1].cpu().numpy())
        assert np.abs(y[i].cpu().numpy() - ref[0]).max() <= 1e-5 * np.abs(ref).max()
        assert torch.equal(st.inverse(mag[i : i + 1].clone(), ph[i : i + 1].clone())[0], y[i])
    # size-independent property: the inverse is linear in the magnitude, and scaling by a power of two is exact in every
    # float32 operation of the chain — inverse(2 * mag, phase) == 2 * inverse(mag, phase) bit for bit, over the whole batch
    assert torch.equal(st.inverse(mag * 2.0, ph), y * 2.0)


def test_c5_full_batch_vocos_forward_8192_and_inverse_1024():
    """BASELINE configs[4] at full size: forward mel on B = 8192 x 5 s (sampled clips vs the oracle <= 1e-4, batched ==
    alone bit for bit) and the iSTFT head's inverse on B = 1024 x (513, 468) (sampled items vs the oracle <= 1e-5 of the
    peak, batched == alone, 119 552 samples per item as codec/tests/test_vocos.py:61-73 pins)."""
    from mlx_audio_plus_b200.codec.models.vocos.mel import log_mel_spectrogram
    from mlx_audio_plus_b200.dsp import hanning, istft

    free, _ = torch.cuda.mem_get_info()
    B = 8192 if free > 20e9 else 1024
    x = _bench_like_batch(B, 120000, 24000, 1238)
    y = log_mel_spectrogram(x)
    assert tuple(y.shape) == (B, 468, 100) and bool(torch.isfinite(y).all())
    for i in (0, 3, B // 2 + 1, B - 1):
        ref = W.vocos_log_mel(x[i].cpu().numpy())[0]
        assert np.abs(y[i].cpu().numpy() - ref).max() <= 1e-4
        assert torch.equal(log_mel_spectrogram(x[i : i + 1].clone())[0], y[i])
    del x, y
    Bi = 1024
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    mag = torch.exp(0.5 * torch.randn((Bi, 513, 468), generator=g, device="cuda")).clamp(max=1e2)
    ph = torch.randn((Bi, 513, 468), generator=g, device="cuda")
    spec = torch.complex(mag * torch.cos(ph), mag * torch.sin(ph)).contiguous()
    del mag, ph
    w = hanning(1024)
    out = istft(spec, window=w, hop_length=256, win_length=1024)
    assert tuple(out.shape) == (Bi, 119552)
    for i in (0, 17, Bi - 1):
        ref = O.istft(spec[i].cpu().numpy(), window=O.hanning(1024), hop_length=256, win_length=1024)
        assert np.abs(out[i].cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()
        assert torch.equal(istft(spec[i].clone(), window=w, hop_length=256, win_length=1024), out[i])


# ---- the step in front of the path: PCM -> resample -> mono (stt/utils.py:21-57) -------------------------------
@pytest.mark.parametrize("orig,target,ch,kind", [(44100, 16000, 2, "i16"), (48000, 16000, 1, "i16"), (8000, 16000, 1, "f32"),
                                                 (22050, 16000, 2, "f32"), (24000, 16000, 3, "i16"), (16000, 16000, 2, "i16"),
                                                 (16000, 24000, 1, "f32"), (44100, 48000, 2, "i16")])
@pytest.mark.parametrize("n", [9, 30011])
def test_load_audio_resample_parity(orig, target, ch, kind, n):
    """one kernel: int16 / 32768 -> resample_poly(padtype="edge") per channel -> float32 mean over channels, against
    scipy in float64 (the reference's own arithmetic); tolerance 1e-5 of the peak (fp32 FIR of <= 61 taps)."""
    from mlx_audio_plus_b200.stt.utils import load_audio, resample_audio
    from oracle import pre_oracle as P

    rng = np.random.default_rng(n + ch)
    t = np.arange(n) / orig
    wave = 0.4 * np.sin(2 * np.pi * 440 * t)[:, None] + 0.1 * rng.standard_normal((n, ch))
    pcm = np.clip(np.round(wave * 32768), -32768, 32767).astype(np.int16)
    src = pcm if kind == "i16" else (pcm.astype(np.float64) / 32768.0).astype(np.float32)
    ref = P.load_audio_from_pcm(src if kind == "i16" else src.astype(np.float64), orig, target)
    for put in (lambda a: a, dev):
        y = host(load_audio(pcm=put(src), sample_rate=orig, sr=target))
        assert y.shape == ref.shape and y.dtype == np.float32
        tol = 0.0 if orig == target else 1e-5 * max(np.abs(ref).max(), 1e-3)
        assert np.abs(y - ref).max() <= tol, np.abs(y - ref).max()
    if orig != target:  # per-channel resampling (stt/utils.py:21-29)
        r2 = P.resample_audio(P.pcm_to_float(src if kind == "i16" else src.astype(np.float64)), orig, target)
        y2 = host(resample_audio(dev(src), orig, target))
        assert y2.shape == r2.shape
        assert np.abs(y2 - r2).max() <= 1e-5 * max(np.abs(r2).max(), 1e-3)


def test_load_audio_and_whisper_from_a_pcm16_wave_file(tmp_path):
    """stt/utils.py:32-57 / whisper/audio.py:68-69 with a path: host parse of the container, then the same kernels."""
    import wave

    from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram
    from mlx_audio_plus_b200.stt.utils import load_audio
    from oracle import pre_oracle as P

    rng = np.random.default_rng(8)
    pcm = (synth(40, 44100 * 2, 44100)[:, None] * np.array([9000.0, 5000.0])[None] + rng.normal(0, 50, (88200, 2))).astype(np.int16)
    path = str(tmp_path / "clip.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(44100)
        w.writeframes(pcm.astype("<i2").tobytes())
    a = load_audio(path)
    b = load_audio(pcm=pcm, sample_rate=44100)
    assert a.shape == (32000,) and np.array_equal(host(a), host(b))
    ref = P.load_audio_from_pcm(pcm, 44100, 16000)
    assert np.abs(host(a) - ref).max() <= 1e-5 * np.abs(ref).max()
    m = host(log_mel_spectrogram(path, n_mels=80))
    assert m.shape == (200, 80) and np.abs(m - W.whisper_log_mel(host(a), 80)).max() <= 1e-4


def test_resample_batch_matches_loop():
    from mlx_audio_plus_b200.stt.utils import _run

    rng = np.random.default_rng(5)
    pcm = rng.integers(-20000, 20000, size=(3, 5000, 2), dtype=np.int16)
    yb = host(_run(dev(pcm), 160, 441, True))
    for i in range(3):
        np.testing.assert_array_equal(yb[i], host(_run(dev(pcm[i]), 160, 441, True)))


# ---- the steps right after the path ----------------------------------------------------------
def test_funasr_frontend_lfr_cmvn_parity(golden):
    from mlx_audio_plus_b200.stt.models.funasr import audio as FA

    g = golden("post")
    for put in (lambda a: a, dev):
        lm = FA.log_mel_spectrogram(put(g["funasr|x"]))
        assert np.abs(host(lm) - g["funasr|logmel"]).max() <= 1e-4 * np.log(10) * 4  # the Whisper bound in ln units
        np.testing.assert_array_equal(host(FA.apply_lfr(put(g["funasr|logmel"]))), g["funasr|lfr"])  # a gather: bit-exact
        np.testing.assert_array_equal(host(FA.apply_lfr(put(g["funasr|logmel"]), 5, 3)), g["funasr|lfr_5_3"])
        np.testing.assert_array_equal(host(FA.apply_lfr(put(g["funasr|logmel"][:4]))), g["funasr|lfr_short"])
        np.testing.assert_array_equal(host(FA.apply_cmvn(put(g["funasr|lfr"]), g["funasr|cmvn_mean"], g["funasr|cmvn_istd"])),
                                      g["funasr|lfr_cmvn"])
        u = host(FA.apply_cmvn(put(g["funasr|lfr"])))  # per-utterance: float64 statistics on the device
        assert u.shape == g["funasr|lfr_cmvn_utt"].shape and np.abs(u - g["funasr|lfr_cmvn_utt"]).max() <= 5e-5
    ub = host(FA.apply_cmvn(dev(np.stack([g["funasr|lfr"], 2 * g["funasr|lfr"] + 1]))))
    assert np.abs(ub[1] - ub[0]).max() <= 5e-5  # (x - mean) / std is invariant under x -> 2x + 1
    for cols in (80, 81, 7):  # narrow rows (whole-warp blocks) and the scalar path (width not a multiple of 4)
        xr = np.random.default_rng(cols).standard_normal((3, 1000, cols)).astype(np.float32) * 3 + 1
        ur = host(FA.apply_cmvn(dev(xr)))
        ref = (xr - xr.mean(axis=1, keepdims=True, dtype=np.float64)) / (xr.std(axis=1, keepdims=True, dtype=np.float64) + 1e-6)
        assert ur.shape == xr.shape and np.abs(ur - ref).max() <= 5e-6 * np.abs(ref).max()
    fused = host(FA.preprocess_audio(dev(g["funasr|x"]), cmvn_mean=g["funasr|cmvn_mean"], cmvn_istd=g["funasr|cmvn_istd"]))
    assert fused.shape == g["funasr|lfr_cmvn"].shape
    assert np.abs(fused - g["funasr|lfr_cmvn"]).max() <= 1e-3 * 1.5  # log-mel tolerance times the largest istd (1.5)
    xb = np.stack([g["funasr|logmel"][:100], g["funasr|logmel"][50:150]])
    yb = host(FA.apply_lfr(dev(xb)))
    for i in range(2):
        np.testing.assert_array_equal(yb[i], W.funasr_apply_lfr(xb[i]))


def test_whisper_mel_segment_parity(golden):
    from mlx_audio_plus_b200.stt.models.whisper.audio import mel_segment

    g = golden("post")
    mel = g["whisper|mel"]
    for k in [k for k in g.files if k.startswith("whisper|seg|")]:
        _, _, seek, size = k.split("|")
        for put in (lambda a: a, dev):
            y = host(mel_segment(put(mel), int(seek), int(size), 500))
            assert y.dtype == np.float16
            np.testing.assert_array_equal(y, g[k])  # float32 -> float16 round-to-nearest-even, zero rows: bit-exact
    y32 = host(mel_segment(dev(mel), 10, 10_000, 3000, "float32"))  # segment longer than what is left: trimmed to T - seek
    np.testing.assert_array_equal(y32, W.whisper_mel_segment(mel, 10, 10_000, 3000, np.float32))
    yb = mel_segment(dev(mel), 0, 77, 96, "bfloat16")
    ref = torch.from_numpy(W.whisper_mel_segment(mel, 0, 77, 96, np.float32)).to(torch.bfloat16)
    assert yb.dtype == torch.bfloat16 and torch.equal(yb.cpu(), ref)
    b = host(mel_segment(dev(np.stack([mel, 2 * mel])), 5, 50, 64))
    np.testing.assert_array_equal(b[1], W.whisper_mel_segment(2 * mel, 5, 50, 64))


def test_whisper_16bit_epilogue_is_the_cast_of_the_float32_result(golden):
    """out_dtype float16 / bfloat16: phase B stores the encoder's dtype directly; the clamp fix-up works on the 16-bit rows.
    cast(max(y, floor)) == max(cast(y), cast(floor)), so the result is BIT-identical to casting the float32 features
    (whisper/whisper.py:994-996 `.astype(self.dtype)`), also where the max-8 clamp is active."""
    from mlx_audio_plus_b200.frontend import FrontendPlan
    from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram

    g = golden("models")
    for key in ("whisper|x", "whisper|sil|x"):
        x = g[key]
        for n_mels in (80, 128):
            y32 = log_mel_spectrogram(dev(x), n_mels=n_mels)
            y16 = log_mel_spectrogram(dev(x), n_mels=n_mels, dtype="float16")
            yb = log_mel_spectrogram(dev(x), n_mels=n_mels, dtype=torch.bfloat16)
            assert y16.dtype == torch.float16 and yb.dtype == torch.bfloat16 and y16.shape == y32.shape
            assert torch.equal(y16, y32.to(torch.float16))
            assert torch.equal(yb, y32.to(torch.bfloat16))
    xb = np.stack([g["whisper|x"], g["whisper|sil|x"]])
    y = log_mel_spectrogram(dev(xb), n_mels=128, padding=4000, dtype="float16")
    assert torch.equal(y, log_mel_spectrogram(dev(xb), n_mels=128, padding=4000).to(torch.float16))
    yh = log_mel_spectrogram(g["whisper|sil|x"], n_mels=80, dtype="float16")  # host path: half the D2H bytes
    assert yh.dtype == np.float16
    np.testing.assert_array_equal(np.asarray(yh), host(log_mel_spectrogram(dev(g["whisper|sil|x"]), n_mels=80)).astype(np.float16))
    with pytest.raises(NotImplementedError):  # Parakeet: cross-frame normalisation needs the float32 second sweep
        FrontendPlan(n_fft=512, hop=160, window=np.asarray(O.hanning(400)), spec_kind=1, log_kind=2, guard_kind=2,
                     guard_eps=1e-5, filterbank=np.asarray(O.mel_filters(16000, 512, 80, norm=None, mel_scale=None)),
                     norm_kind=1, norm_eps=1e-5, out_dtype="float16")
    with pytest.raises(NotImplementedError):  # complex spectrum
        FrontendPlan(n_fft=400, hop=160, window=np.asarray(O.hanning(400)), out_dtype="bfloat16")


# ---- seeded random sweeps (property-style coverage of the parameter space against the oracle) -------------
def _sweep_cases(seed, n):
    rng = np.random.default_rng(seed)
    sizes = [(400, 160), (512, 160), (1024, 256), (800, 200), (20, 5), (16, 4), (64, 16), (360, 90), (1920, 480), (48, 12)]
    out = []
    for i in range(n):
        n_fft, hop = sizes[int(rng.integers(len(sizes)))]
        if rng.random() < 0.25:
            hop = int(rng.integers(1, n_fft // 2 + 1))
        win = n_fft if rng.random() < 0.6 else int(rng.integers(max(2, n_fft // 3), n_fft + 1))
        center = bool(rng.random() < 0.75)
        mode = "reflect" if rng.random() < 0.7 else "constant"
        wname = ["hann", "hamming", "blackman", "bartlett"][int(rng.integers(4))]
        p = n_fft // 2
        edge = [p + 1, p + 2, n_fft, n_fft + 1, 3 * hop + n_fft - 1, 10 * hop, 10 * hop + 1, 11 * hop - 1]
        L = int(edge[int(rng.integers(len(edge)))]) if rng.random() < 0.5 else int(rng.integers(n_fft, 6 * n_fft + 40 * hop))
        if not center:
            L = max(L, n_fft)
        out.append((n_fft, hop, win, wname, center, mode, L, int(rng.integers(1, 4))))
    return out


@pytest.mark.parametrize("case", _sweep_cases(2024, 48), ids=lambda c: "-".join(map(str, c)))
def test_stft_random_sweep(case):
    from mlx_audio_plus_b200.dsp import stft

    n_fft, hop, win, wname, center, mode, L, B = case
    xb = np.stack([synth(1000 + i + L, L) * (0.3 + 0.4 * i) for i in range(B)])
    y = host(stft(dev(xb if B > 1 else xb[0]), n_fft, hop, win, wname, center, mode))
    y = y if B > 1 else y[None]
    for i in range(B):
        assert_stft_close(y[i], O.stft(xb[i], n_fft, hop, win, wname, center, mode))


@pytest.mark.parametrize("case", _sweep_cases(77, 32), ids=lambda c: "-".join(map(str, c)))
def test_istft_random_sweep(case):
    from mlx_audio_plus_b200.dsp import istft

    n_fft, hop, _, wname, center, _, L, B = case
    if n_fft % 2 or hop > n_fft:
        pytest.skip("istft needs an even n_fft and hop <= n_fft")
    rng = np.random.default_rng(L + n_fft)
    T = max(2, L // hop % 97 + 2)
    F = n_fft // 2 + 1
    spec = (rng.standard_normal((B, F, T)) + 1j * rng.standard_normal((B, F, T))).astype(np.complex64)
    normalized = bool(rng.random() < 0.5)
    length = None if rng.random() < 0.6 else int(rng.integers(1, (T - 1) * hop + n_fft))
    y = host(istft(dev(spec if B > 1 else spec[0]), hop, n_fft, wname, center, length, normalized))
    y = y if B > 1 else y[None]
    for i in range(B):
        assert_wave_close(y[i], O.istft(spec[i], hop, n_fft, wname, center, length, normalized))


@pytest.mark.parametrize("L,padding,n_mels,B", [(201, 0, 80, 1), (400, 0, 128, 1), (5279, 0, 80, 2), (5280, 160, 128, 1), (16001, 0, 128, 3),
                                                (31999, 4000, 80, 2), (48000, 48000, 128, 1), (160 * 37 + 1, 7, 80, 1)])
def test_whisper_length_sweep(L, padding, n_mels, B):
    """edge tiles of the fused kernel (clip start / end inside one tile, ragged last tile, virtual right padding, odd lengths)
    against the oracle; batch == per-clip calls"""
    from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram

    xb = np.stack([synth(500 + i + L, L) * (0.2 + 0.5 * i) for i in range(B)])
    y = host(log_mel_spectrogram(dev(xb if B > 1 else xb[0]), n_mels=n_mels, padding=padding))
    y = y if B > 1 else y[None]
    for i in range(B):
        ref = W.whisper_log_mel(xb[i], n_mels, padding)
        assert y[i].shape == ref.shape
        assert np.abs(y[i] - ref).max() <= 1e-4


@pytest.mark.parametrize("L", [400, 4001, 16000, 160 * 64 + 3])
def test_parakeet_and_vocos_length_sweep(L):
    from mlx_audio_plus_b200.codec.models.vocos.mel import log_mel_spectrogram as vocos_mel
    from mlx_audio_plus_b200.stt.models.parakeet.audio import PreprocessArgs, log_mel_spectrogram as pk_mel

    x = synth(900 + L, L)
    pa = PreprocessArgs(sample_rate=16000, normalize="per_feature", window_size=0.025, window_stride=0.01, window="hann",
                        features=80, n_fft=512, dither=0.0)
    opa = W.PreprocessArgs(sample_rate=16000, normalize="per_feature", window_size=0.025, window_stride=0.01, window="hann",
                           features=80, n_fft=512, dither=0.0)
    y = host(pk_mel(dev(x), pa))
    ref = W.parakeet_log_mel(x, opa)
    assert y.shape == ref.shape and np.abs(y - ref).max() <= 5e-4 * max(1.0, np.abs(ref).max() / 5)
    if L >= 1024:
        x24 = synth(901 + L, L, sr=24000)
        yv = host(vocos_mel(dev(x24)))
        rv = W.vocos_log_mel(x24)
        assert yv.shape == rv.shape and np.abs(yv - rv).max() <= 1e-4 * np.log(10) * 4


def test_whisper_padding_rows_are_filled_not_transformed():
    """log_mel_spectrogram(audio, padding=N_SAMPLES) (how whisper.py always calls it): frames that see only the virtual zero
    padding are written as constant rows; the result is bit-identical to transforming them, also in float16, and the
    max-8 clamp still applies to them."""
    import os

    from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram

    for L, padding, scale in ((16000 * 3, 480000, 1.0), (16000 * 2 + 123, 16000 * 4, 1e-4), (5000, 12000, 1.0)):
        xb = np.stack([synth(700 + L, L) * scale, synth(701 + L, L) * scale * 0.1])
        for dt in ("float32", "float16"):
            y = log_mel_spectrogram(dev(xb), n_mels=128, padding=padding, dtype=dt)
            os.environ["B2A_NO_PAD_SKIP"] = "1"
            try:
                y_full = log_mel_spectrogram(dev(xb), n_mels=128, padding=padding, dtype=dt)
            finally:
                del os.environ["B2A_NO_PAD_SKIP"]
            assert torch.equal(y, y_full)
        ref = W.whisper_log_mel(xb[0], 128, padding)
        assert np.abs(host(log_mel_spectrogram(dev(xb[0]), n_mels=128, padding=padding)) - ref).max() <= 1e-4


@pytest.mark.parametrize("n_fft,hop,win,n_mels,sr,kind", [(800, 200, 800, 80, 16000, "power"), (1024, 320, 640, 128, 16000, "magnitude")])
def test_fast_logmel_800_and_1024x320(n_fft, hop, win, n_mels, sr, kind):
    """the two extra sizes of the fast family (dsp.stft's defaults 800/200; Spark's 1024/320 with a 640-tap window,
    bicodec.py:20-49) with a run-time filterbank: log-mel against the oracle's stft + mel_filters"""
    from mlx_audio_plus_b200 import _lib as L
    from mlx_audio_plus_b200._arrays import Ingested
    from mlx_audio_plus_b200.frontend import FrontendPlan

    w = np.asarray(O.hanning(win + 1)[:-1]) if win != n_fft else np.asarray(O.hanning(win))
    fb = np.asarray(O.mel_filters(sr, n_fft, n_mels, norm="slaney", mel_scale=None))
    plan = FrontendPlan(n_fft=n_fft, hop=hop, window=w, filterbank=fb, spec_kind=L.SPEC_POWER if kind == "power" else L.SPEC_MAGNITUDE,
                        log_kind=L.LOG_LN, guard_kind=L.GUARD_MAX, guard_eps=1e-5)
    assert plan.kernel_name == f"fast_logmel_{n_fft}x{hop}"
    xb = np.stack([synth(60 + i, 30011) * (0.5 + i) for i in range(2)])
    xd = dev(xb)
    y = host(plan.run(Ingested("torch", True, xd, None, xd.device)))
    for i in range(2):
        S = np.abs(O.stft(xb[i], n_fft, hop, win, w))
        S = np.square(S) if kind == "power" else S
        ref = np.log(np.maximum((S.astype(np.float32) @ fb.T).astype(np.float32), np.float32(1e-5)))
        assert y[i].shape == ref.shape and np.abs(y[i] - ref).max() <= 1e-4 * np.log(10) * 4


# ---- the remaining thin dsp callers ---------------------------------------------------
def test_variant_wrappers_parity(golden):
    from mlx_audio_plus_b200.codec.models.s3gen.mel import mel_spectrogram as s3gen_mel
    from mlx_audio_plus_b200.stt.models.glmasr.glmasr import preprocess_audio
    from mlx_audio_plus_b200.tts.models.chatterbox.voice_encoder.config import VoiceEncConfig
    from mlx_audio_plus_b200.tts.models.chatterbox.voice_encoder.melspec import melspectrogram
    from mlx_audio_plus_b200.tts.models.indextts.mel import log_mel_spectrogram as indextts_mel
    from mlx_audio_plus_b200.tts.models.soprano.decoder import ISTFTHead as SopranoHead
    from mlx_audio_plus_b200.tts.models.spark.bicodec import mel_spectrogram as spark_mel
    from mlx_audio_plus_b200.vad.models.smart_turn.smart_turn import ProcessorConfig, prepare_input_features

    g = golden("variants")

    def close(y, ref, atol):
        y = host(y)
        assert y.shape == ref.shape, (y.shape, ref.shape)
        assert np.abs(y - ref).max() <= atol, np.abs(y - ref).max()

    for put in (dev, np.asarray):  # torch CUDA in / out, and the NumPy (host-buffer) entry
        close(s3gen_mel(put(g["s3gen|x"])), g["s3gen|y"], 1e-4)
        close(indextts_mel(put(g["indextts|x"])), g["indextts|y"], 1e-4)
        close(spark_mel(put(g["spark|x"])), g["spark|y"], 2e-5 * np.abs(g["spark|y"]).max())
        close(melspectrogram(put(g["ve|x"]), VoiceEncConfig()), g["ve|amp"], 2e-5 * np.abs(g["ve|amp"]).max())
        close(preprocess_audio(put(g["glmasr|x"])), g["glmasr|y"], 1e-4)
    close(s3gen_mel(dev(g["s3gen|x"][1])), g["s3gen|y1d"], 1e-4)
    close(indextts_mel(dev(g["indextts|x"]), padding=500), g["indextts|ypad"], 1e-4)
    close(melspectrogram(dev(g["ve|x"][0])), g["ve|amp1d"], 2e-5 * np.abs(g["ve|amp"]).max())
    close(melspectrogram(dev(g["ve|x"]), VoiceEncConfig(mel_power=1.0, mel_type="db", normalized_mels=True)), g["ve|db_norm"], 1e-4)
    close(melspectrogram(dev(g["ve|x"]), VoiceEncConfig(mel_type="db")), g["ve|db"], 20 * 1e-4)  # 20 log10: 20 x the log-mel bound
    with pytest.raises(NotImplementedError):
        melspectrogram(dev(g["ve|x"]), VoiceEncConfig(mel_power=1.5))
    assert_wave_close(SopranoHead(8, 2048, 512)(dev(g["soprano|head_in"])), g["soprano|head_out"])
    y3 = dev(g["glmasr|y"])
    assert preprocess_audio(y3) is y3  # 3-D input is taken as features (glmasr.py:569-570)
    pc = ProcessorConfig(max_audio_seconds=2)
    for n in ("short", "long"):
        close(prepare_input_features(dev(g[f"smart|{n}|x"]), pc), g[f"smart|{n}|y"], 2e-4)  # torch mean / std of the waveform
        close(prepare_input_features(g[f"smart|{n}|x"], pc), g[f"smart|{n}|y"], 1e-4)


def test_lfm2_and_mossformer_parity(golden):
    from mlx_audio_plus_b200.dsp import hamming
    from mlx_audio_plus_b200.sts.models.lfm_audio.detokenizer import istft_same
    from mlx_audio_plus_b200.sts.models.lfm_audio.processor import AudioPreprocessor, PreprocessorConfig
    from mlx_audio_plus_b200.sts.models.mossformer2_se.model import chunk_istft, chunk_stft

    g = golden("variants")
    pre = AudioPreprocessor(PreprocessorConfig(dither=0.0))
    for put in (dev, np.asarray):
        y = host(pre(put(g["lfm2|x"])))
        assert y.shape == g["lfm2|y"].shape and np.abs(y - g["lfm2|y"]).max() <= 5e-4  # normalised: the Parakeet bound
    y1 = host(pre(dev(g["lfm2|x"][1])))
    assert y1.shape == g["lfm2|y1d"].shape and np.abs(y1 - g["lfm2|y1d"]).max() <= 5e-4
    # the statistic covers the first len // hop frames only: it must differ from the all-frames normalisation
    ref_all = W.sortformer_mel  # noqa: F841  (same chain with all-frame statistics; see test_other_frontends_parity)
    raw = host(AudioPreprocessor(PreprocessorConfig(dither=0.0, normalize="none"))(dev(g["lfm2|x"])))
    n = g["lfm2|x"].shape[1] // 160
    mean = raw[:, :n].mean(axis=1, keepdims=True, dtype=np.float64)
    std = raw[:, :n].astype(np.float64).std(axis=1, keepdims=True, ddof=1) + 1e-5
    assert np.abs((raw - mean) / std - host(pre(dev(g["lfm2|x"])))).max() <= 5e-4
    yd = host(pre.__class__(PreprocessorConfig())(dev(g["lfm2|x"])))  # default dither 1e-5: same features to ~1e-2
    assert yd.shape == g["lfm2|y"].shape and np.isfinite(yd).all()
    for put in (dev, np.asarray):
        assert_wave_close(istft_same(put(g["lfm2|mag"]), put(g["lfm2|phase"]), g["lfm2|w"], 1280, 320), g["lfm2|wave"])
    w = hamming(1920, periodic=False)
    re, im = chunk_stft(dev(g["moss|x"]), window=w)
    assert tuple(re.shape) == (961, 21)
    assert_stft_close(torch.complex(re, im), (g["moss|re"] + 1j * g["moss|im"]).astype(np.complex64))
    assert_wave_close(chunk_istft(dev(g["moss|re"]), dev(g["moss|im"]), window=w, chunk_length=9600), g["moss|y"])
    re_h, im_h = chunk_stft(g["moss|x"])  # NumPy in, default window
    assert_wave_close(chunk_istft(re_h, im_h, chunk_length=9600), g["moss|y"])  # mask of ones: the reference's own round trip


def test_chatterbox_turbo_and_cosyvoice2_hift_parity(golden):
    from mlx_audio_plus_b200.tts.models.chatterbox_turbo.models.s3gen import hifigan as CT
    from mlx_audio_plus_b200.tts.models.cosyvoice2 import hifigan as C2

    g = golden("variants")
    w = CT.hann_window_periodic(16)
    np.testing.assert_array_equal(np.asarray(w), O.hanning(16, True))
    for put in (dev, np.asarray):
        re, im = CT.stft(put(g["cturbo|x"]), 16, 4, w)
        assert np.abs(host(re) - g["cturbo|re"]).max() <= 1e-5 * np.abs(g["cturbo|re"]).max()
        assert np.abs(host(im) - g["cturbo|im"]).max() <= 1e-5 * np.abs(g["cturbo|re"]).max()
        re, im = CT.stft(put(g["cturbo|x"][:, :9]), 16, 4, w)  # shorter than n_fft: one zero-extended frame
        assert tuple(re.shape) == (2, 9, 1) and np.abs(host(re) - g["cturbo|short|re"]).max() <= 1e-5
        assert np.abs(host(im) - g["cturbo|short|im"]).max() <= 1e-5
        assert_wave_close(CT.istft(put(g["cturbo|mag"]), put(g["cturbo|phase"]), 16, 4, w), g["cturbo|y"])
    h = golden("hift")  # CosyVoice2 forwards to the S3Gen pair
    re, im = C2.stft(dev(h["n16|x"]), 16, 4, C2.hann_window_periodic(16))
    assert np.abs(host(re) - h["n16|hift_s3gen|re"]).max() <= 1e-5 * np.abs(h["n16|hift_s3gen|re"]).max()
    assert_wave_close(C2.istft(dev(h["n16|mag"]), dev(h["n16|phase"]), 16, 4, C2.hann_window_periodic(16)), h["n16|hift_s3gen|y"])


def _xvector_fbank_f64(x, num_mel_bins=80):
    """float64 evaluation of the xvector.py:38-150 chain (float32 window and filterbank values, float64 arithmetic)"""
    x = np.asarray(x, np.float64).squeeze()
    m = (x.shape[0] - 400) // 160 + 1
    fr = np.lib.stride_tricks.as_strided(x, shape=(m, 400), strides=(8 * 160, 8)).copy()
    fr -= fr.mean(axis=1, keepdims=True)
    fr = np.concatenate([fr[:, :1], fr[:, 1:] - np.float64(np.float32(0.97)) * fr[:, :-1]], axis=1)
    k = np.arange(400).astype(np.float32)
    w = np.power(np.float32(0.5) - np.float32(0.5) * np.cos(np.float32(2) * np.float32(np.pi) * k / np.float32(399)), np.float32(0.85))
    p = np.abs(np.fft.rfft(fr * w.astype(np.float64), n=512, axis=1)) ** 2
    fb = O.mel_filters(16000, 512, num_mel_bins, 20.0, 8000.0, None, "htk").astype(np.float64)
    return np.log(np.maximum(p @ fb.T, 1.1920929e-07))


def test_xvector_fbank_parity(golden):
    """CAMPPlus front-end.  Pre-emphasis and DC removal leave the lowest mel bins ~1e5 below the strongest ones, so their
    float32 FFT round-off is ~1e-4 relative whatever the implementation (the reference fixture itself is 6e-5 off the float64
    evaluation there).  Bounds: 1e-4 in log units on the bins within e^-9 of the strongest one; on ALL bins 2e-3 in log units
    and, in the linear domain, 1e-5 of the frame's peak mel energy against both the fixture and the float64 evaluation."""
    from mlx_audio_plus_b200.codec.models.s3gen.xvector import kaldi_fbank

    g = golden("variants")
    truth = _xvector_fbank_f64(g["xvector|x"])

    def lin_err(y, ref):
        return (np.abs(np.exp(y.astype(np.float64)) - np.exp(ref.astype(np.float64))) / np.exp(ref).max(axis=1, keepdims=True)).max()

    strong = g["xvector|y"] > g["xvector|y"].max() - 9.0  # within e^-9 of the strongest bin
    for put in (dev, np.asarray):
        y = host(kaldi_fbank(put(g["xvector|x"])))
        assert y.shape == g["xvector|y"].shape
        assert np.abs(y - g["xvector|y"])[strong].max() <= 1e-4
        assert np.abs(y - g["xvector|y"]).max() <= 2e-3
        assert lin_err(y, g["xvector|y"]) <= 1e-5 and lin_err(y, truth) <= 1e-5, (lin_err(y, g["xvector|y"]), lin_err(y, truth))
    y = host(kaldi_fbank(dev(g["xvector|x"][None, :4000]), num_mel_bins=40))
    t40 = _xvector_fbank_f64(g["xvector|x"][:4000], 40)
    assert y.shape == g["xvector|y40"].shape and np.abs(y - g["xvector|y40"]).max() <= 2e-3 and lin_err(y, t40) <= 1e-5
    short = g["xvector|x"][:250]  # shorter than one window: one zero-extended frame (xvector.py:77-78, 103-112)
    ys = host(kaldi_fbank(dev(short)))
    assert ys.shape == (1, 80) and lin_err(ys, W.s3gen_xvector_fbank(short)) <= 1e-5


def test_hf_whisper_feature_extractor_parity():
    """Qwen3-ASR / Qwen3-ForcedAligner features (qwen3_asr.py:800-846): the drop-in against the real transformers
    WhisperFeatureExtractor's output (tests/golden/hf_whisper_fe.npz), with the reference's own call arguments."""
    import os

    from mlx_audio_plus_b200.stt.models.qwen3_asr.feature_extractor import WhisperFeatureExtractor

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hf_whisper_fe.npz"))
    fe = WhisperFeatureExtractor(feature_size=128)
    np.testing.assert_array_equal(fe.mel_filters, g["fb128"])
    kw = dict(sampling_rate=16000, return_attention_mask=True, truncation=False, padding=True, return_tensors="np")
    o = fe(g["one|x"], **kw)
    assert o["input_features"].dtype == np.float32 and o["input_features"].shape == (1, 128, 200)
    assert np.abs(o["input_features"] - g["one|features"]).max() <= 1e-4
    np.testing.assert_array_equal(o["attention_mask"], g["one|mask"])
    o = fe([g["batch|a"], g["batch|b"], g["batch|c"]], **kw)
    assert np.abs(o["input_features"] - g["batch|features"]).max() <= 1e-4
    np.testing.assert_array_equal(o["attention_mask"], g["batch|mask"])
    o = fe(g["one|x"], sampling_rate=16000, return_tensors="np", max_length=48000)
    assert set(o) == {"input_features"} and np.abs(o["input_features"] - g["default|features"]).max() <= 1e-4
    o = fe(np.concatenate([g["one|x"], g["one|x"]]), sampling_rate=16000, return_tensors="np", max_length=48000,
           return_attention_mask=True, do_normalize=True)
    assert np.abs(o["input_features"] - g["trunc_norm|features"]).max() <= 1e-4
    np.testing.assert_array_equal(o["attention_mask"], g["trunc_norm|mask"])
    t = fe(dev(g["one|x"]), sampling_rate=16000, padding=True, truncation=False, return_tensors="cuda")["input_features"]
    assert t.is_cuda and np.abs(host(t) - g["one|features"]).max() <= 1e-4
    with pytest.raises(ValueError):
        fe(g["one|x"], sampling_rate=8000)
    # rectangular 2-D batches (NumPy and torch CUDA) take the copy-free path: same rows as clip-by-clip calls
    xb = np.stack([g["batch|a"], 0.5 * g["batch|a"][::-1].copy()])
    ob = fe(xb, **kw)
    for i in range(2):
        oi = fe(xb[i], **kw)
        np.testing.assert_array_equal(ob["input_features"][i], oi["input_features"][0])
    od = fe(dev(xb), sampling_rate=16000, max_length=48000, return_attention_mask=True, return_tensors="cuda")  # padded to 3 s
    assert tuple(od["input_features"].shape) == (2, 128, 300) and int(od["attention_mask"].sum()) == 2 * 150
    np.testing.assert_array_equal(host(od["input_features"])[0], fe(xb[0], sampling_rate=16000, max_length=48000)["input_features"][0])


def test_transpose_pad_kernel_bit_exact():
    """csrc/post.cu transpose_pad_kernel: (B, T, M) -> (B, M, T_padded), zeros behind column T (Sortformer pad_to)."""
    from mlx_audio_plus_b200._post import transpose_pad

    rng = np.random.default_rng(5)
    for B, T, M, Tp in ((2, 57, 80, 64), (1, 1, 3, 1), (3, 301, 128, 304), (2, 64, 33, 64)):
        x = rng.standard_normal((B, T, M)).astype(np.float32)
        ref = np.zeros((B, M, Tp), np.float32)
        ref[:, :, :T] = np.swapaxes(x, 1, 2)
        np.testing.assert_array_equal(host(transpose_pad(dev(x), Tp)), ref)
        np.testing.assert_array_equal(np.asarray(transpose_pad(x, Tp)), ref)
