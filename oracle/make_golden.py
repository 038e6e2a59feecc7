#!/usr/bin/env python
"""ORACLE / TEST INFRASTRUCTURE — generates tests/golden/refshim_*.npz.

Executes the UNMODIFIED reference sources under /root/reference (read-only) with
`mlx.core` resolved to oracle/mlx_shim (a NumPy float32 stand-in; MLX itself is not
installable in this image) and stores inputs + outputs as small fixtures.  The fixtures
travel with the repo; /root/reference does not exist on the GPU box, so this script is
only ever run in the build container:

    python oracle/make_golden.py            # writes tests/golden/refshim_*.npz

Reference files executed (whole file or AST-extracted definitions, never copied):
  mlx_audio/dsp.py                                   (whole module)
  mlx_audio/stt/models/whisper/audio.py              (whole module)
  mlx_audio/stt/models/parakeet/audio.py             (whole module)
  mlx_audio/stt/models/voxtral_realtime/audio.py     (whole module)
  mlx_audio/codec/models/vocos/mel.py                (whole module)
  mlx_audio/codec/models/s3tokenizer/utils.py        (whole module)
  mlx_audio/codec/models/vocos/vocos.py              (class ISTFTHead)
  mlx_audio/tts/models/kokoro/istftnet.py            (mlx_angle, mlx_unwrap, MLXSTFT)
  mlx_audio/tts/models/qwen3_tts/qwen3_tts.py        (mel_spectrogram)
  mlx_audio/vad/models/sortformer/sortformer.py      (preemphasis_filter, extract_mel_features)
  SURVEY §8a row a12 (refshim_variants.npz):
  mlx_audio/codec/models/s3gen/mel.py                (_reflect_pad_2d, mel_spectrogram)
  mlx_audio/tts/models/indextts/mel.py               (whole module)
  mlx_audio/tts/models/spark/bicodec.py              (mel_spectrogram)
  mlx_audio/tts/models/chatterbox/voice_encoder/melspec.py, config.py (melspectrogram, VoiceEncConfig)
  mlx_audio/tts/models/soprano/decoder.py            (class ISTFTHead)
  mlx_audio/stt/models/glmasr/glmasr.py              (method Model._preprocess_audio)
  mlx_audio/vad/models/smart_turn/smart_turn.py      (methods _prepare_audio_array, prepare_input_features)
  mlx_audio/sts/models/lfm_audio/processor.py        (class AudioPreprocessor)
  mlx_audio/sts/models/lfm_audio/detokenizer.py      (method _istft)
  mlx_audio/sts/models/mossformer2_se/model.py:396-428 (its two dsp calls, made with the same positional arguments)
"""
from __future__ import annotations

import ast
import importlib.util
import math
import os
import sys
import types

import numpy as np

REF = os.environ.get("B2A_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _load_reference():
    sys.path.insert(0, os.path.join(HERE, "mlx_shim"))
    import mlx.core as mx  # the shim
    import mlx.nn as nn

    def pkg(name):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        return m

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    for p in ("mlx_audio", "mlx_audio.stt", "mlx_audio.stt.models", "mlx_audio.codec",
              "mlx_audio.codec.models"):
        pkg(p)
    dsp = load("mlx_audio.dsp", "mlx_audio/dsp.py")
    # mlx_audio/utils.py:29-38 re-exports these names from dsp; the real utils.py drags in model
    # loading (huggingface_hub, mlx.nn ...), so a module carrying only the re-export is used.
    utils = types.ModuleType("mlx_audio.utils")
    for n in ("STR_TO_WINDOW_FN", "bartlett", "blackman", "hamming", "hanning", "istft",
              "mel_filters", "stft"):
        setattr(utils, n, getattr(dsp, n))
    sys.modules["mlx_audio.utils"] = utils
    stt_utils = types.ModuleType("mlx_audio.stt.utils")
    stt_utils.load_audio = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no file IO in goldens"))
    sys.modules["mlx_audio.stt.utils"] = stt_utils

    mods = {
        "dsp": dsp,
        "whisper": load("mlx_audio.stt.models.whisper_audio", "mlx_audio/stt/models/whisper/audio.py"),
        "parakeet": load("mlx_audio.stt.models.parakeet_audio", "mlx_audio/stt/models/parakeet/audio.py"),
        "voxtral": load("mlx_audio.stt.models.voxtral_rt_audio", "mlx_audio/stt/models/voxtral_realtime/audio.py"),
        "vocos_mel": load("mlx_audio.codec.models.vocos_mel", "mlx_audio/codec/models/vocos/mel.py"),
        "s3tok": load("mlx_audio.codec.models.s3tok_utils", "mlx_audio/codec/models/s3tokenizer/utils.py"),
        "funasr": load("mlx_audio.stt.models.funasr_audio", "mlx_audio/stt/models/funasr/audio.py"),
    }

    def extract(rel, names, extra=None):
        """exec selected top-level definitions of a reference file (verbatim AST nodes)."""
        src = open(os.path.join(REF, rel)).read()
        tree = ast.parse(src)
        keep = [n for n in tree.body
                if (isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names)
                or (isinstance(n, ast.Assign) and any(isinstance(t, ast.Name) and t.id in names for t in n.targets))]
        ns = {"mx": mx, "nn": nn, "np": np, "math": math, "stft": dsp.stft, "istft": dsp.istft,
              "hanning": dsp.hanning, "mel_filters": dsp.mel_filters}
        ns.update(extra or {})
        exec(compile(ast.Module(body=keep, type_ignores=[]), rel, "exec"), ns)
        return types.SimpleNamespace(**{n: ns[n] for n in names})

    def extract_methods(rel, cls, names, extra=None):
        """exec selected methods of a reference class (verbatim AST nodes) as free functions taking `self`."""
        tree = ast.parse(open(os.path.join(REF, rel)).read())
        c = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls)
        keep = [n for n in c.body if isinstance(n, ast.FunctionDef) and n.name in names]
        for n in keep:
            n.decorator_list = []
        ns = {"mx": mx, "nn": nn, "np": np, "math": math, "Union": tuple, "Optional": tuple}
        ns.update(extra or {})
        exec(compile(ast.Module(body=keep, type_ignores=[]), rel, "exec"), ns)
        return types.SimpleNamespace(**{n: ns[n] for n in names})

    mods["extract"], mods["extract_methods"] = extract, extract_methods
    mods["vocos"] = extract("mlx_audio/codec/models/vocos/vocos.py", ["ISTFTHead"])
    mods["kokoro"] = extract("mlx_audio/tts/models/kokoro/istftnet.py", ["mlx_angle", "mlx_unwrap", "MLXSTFT"])
    mods["qwen3"] = extract("mlx_audio/tts/models/qwen3_tts/qwen3_tts.py", ["mel_spectrogram"])
    mods["hift_s3gen"] = extract("mlx_audio/codec/models/s3gen/hifigan.py", ["stft", "istft"])
    mods["hift_cosy3"] = extract("mlx_audio/tts/models/cosyvoice3/hifigan.py", ["stft", "istft"], {"Tuple": tuple})
    mods["sortformer"] = extract("mlx_audio/vad/models/sortformer/sortformer.py",
                                 ["_LOG_GUARD", "_NORM_CONSTANT", "preemphasis_filter", "extract_mel_features"])
    return mx, mods


def synth(seed, n, sr=16000):
    """The §8(d) synthetic: 0.1*N(0,1) + 0.2*(sin 440 Hz + sin 3 kHz)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / sr
    x = 0.1 * rng.standard_normal(n) + 0.2 * (np.sin(2 * np.pi * 440 * t) + np.sin(2 * np.pi * 3000 * t))
    return x.astype(np.float32)


def hift_goldens(mx, R):
    """tests/golden/refshim_hift.npz: the HiFT model-local stft / istft pairs."""
    A = lambda v: np.asarray(v)
    g = {}
    rng = np.random.default_rng(77)
    for n_fft, hop in ((16, 4), (20, 5)):
        w = np.asarray(R["dsp"].hanning(n_fft + 1)[:-1])  # periodic Hann, as scipy get_window(fftbins=True)
        x = np.stack([synth(300 + i, 1203, 24000) * (0.5 + i) for i in range(3)])
        mag = np.exp(rng.normal(0, 1.5, (3, n_fft // 2 + 1, 151))).astype(np.float32)  # some values beyond the 1e2 clip
        mag[0, 2, 5], mag[1, 3, 7] = 250.0, -0.5  # exercise both clip bounds
        ph = rng.uniform(-np.pi, np.pi, mag.shape).astype(np.float32)
        g[f"n{n_fft}|w"], g[f"n{n_fft}|x"], g[f"n{n_fft}|mag"], g[f"n{n_fft}|phase"] = w, x, mag, ph
        for name in ("hift_s3gen", "hift_cosy3"):
            re, im = R[name].stft(mx.array(x), n_fft, hop, mx.array(w))
            g[f"n{n_fft}|{name}|re"], g[f"n{n_fft}|{name}|im"] = A(re), A(im)
            g[f"n{n_fft}|{name}|y"] = A(R[name].istft(mx.array(mag), mx.array(ph), n_fft, hop, mx.array(w)))
    np.savez_compressed(os.path.join(OUT, "refshim_hift.npz"), **g)


def post_goldens(mx, R):
    """tests/golden/refshim_post.npz: the steps AFTER the path — FunASR log-mel + apply_lfr +
    apply_cmvn (funasr/audio.py:32-169) and Whisper's segment builder pad_or_trim(mel[seek:seek+n], N_FRAMES,
    axis=-2).astype(float16) (whisper/whisper.py:990-996), from the reference's own functions."""
    F, W = R["funasr"], R["whisper"]
    A = np.asarray
    g = {}
    x = synth(300, 16000 * 3 + 77)
    g["funasr|x"] = x
    lm = F.log_mel_spectrogram(mx.array(x))
    g["funasr|logmel"] = A(lm)
    g["funasr|lfr"] = A(F.apply_lfr(lm))
    g["funasr|lfr_5_3"] = A(F.apply_lfr(lm, 5, 3))
    short = lm[:4]
    g["funasr|lfr_short"] = A(F.apply_lfr(short))
    rng = np.random.default_rng(12)
    mean = rng.standard_normal(560).astype(np.float32)
    istd = (0.5 + rng.random(560)).astype(np.float32)
    g["funasr|cmvn_mean"], g["funasr|cmvn_istd"] = mean, istd
    g["funasr|lfr_cmvn"] = A(F.apply_cmvn(F.apply_lfr(lm), mx.array(mean), mx.array(istd)))
    g["funasr|lfr_cmvn_utt"] = A(F.apply_cmvn(F.apply_lfr(lm)))  # per-utterance mean / std (funasr/audio.py:160-164)
    mel = W.log_mel_spectrogram(synth(301, 16000 * 4), n_mels=80)
    g["whisper|mel"] = A(mel)
    for seek, size in ((0, 400), (100, 300), (250, 150)):
        seg = W.pad_or_trim(mel[seek: seek + size], 500, axis=-2).astype(mx.float16)
        g[f"whisper|seg|{seek}|{size}"] = A(seg)
    np.savez_compressed(os.path.join(OUT, "refshim_post.npz"), **g)


def kaldi_goldens(mx, R):
    """tests/golden/refshim_kaldi.npz: the Kaldi-compatible part of mlx_audio/dsp.py (439-676), dither = 0."""
    A = lambda v: np.asarray(v)
    dsp = R["dsp"]
    g = {}
    x48 = (synth(401, 30000, 48000) * 8000.0).astype(np.float32)  # int16-scale amplitudes, as MossFormer2 feeds
    x16 = (synth(402, 9000, 16000) * 8000.0 + 37.5).astype(np.float32)  # with a DC offset
    g["x48"], g["x16"] = x48, x16
    cases = {
        "moss": (x48, dict(sample_rate=48000, win_len=1920, win_inc=384, num_mels=60, win_type="hamming", preemphasis=0.97)),
        "povey16": (x16, dict(sample_rate=16000, win_len=400, win_inc=160, num_mels=80, win_type="povey", preemphasis=0.97)),
        "hann_nopre": (x16, dict(sample_rate=16000, win_len=400, win_inc=160, num_mels=24, win_type="hanning",
                                 preemphasis=0.0, low_freq=0.0, high_freq=-400.0)),
        "rect_noclip": (x16, dict(sample_rate=16000, win_len=512, win_inc=128, num_mels=40, win_type="rectangular",
                                  preemphasis=0.5, snip_edges=False)),
        "short": (x16[:300], dict(sample_rate=16000, win_len=400, win_inc=160, num_mels=23)),
    }
    for name, (x, kw) in cases.items():
        g[f"fbank|{name}"] = A(dsp.compute_fbank_kaldi(mx.array(x), dither=0.0, **kw))
    bins, cf = dsp.get_mel_banks_kaldi(60, 2048, 48000.0, 20.0, 0.0)
    g["banks|60_2048"], g["banks|60_2048|cf"] = A(bins), A(cf)
    bins, cf = dsp.get_mel_banks_kaldi(80, 512, 16000.0, 20.0, -400.0)
    g["banks|80_512"], g["banks|80_512|cf"] = A(bins), A(cf)
    f = g["fbank|moss"].T.copy()
    g["deltas|edge5"] = A(dsp.compute_deltas_kaldi(mx.array(f), win_length=5))
    g["deltas|const9"] = A(dsp.compute_deltas_kaldi(mx.array(f), win_length=9, mode="constant"))
    g["deltas|3d"] = A(dsp.compute_deltas_kaldi(mx.array(f[:24].reshape(2, 12, -1)), win_length=3))
    np.savez_compressed(os.path.join(OUT, "refshim_kaldi.npz"), **g)


def variants_goldens(mx, R):
    """tests/golden/refshim_variants.npz: the thin dsp callers of SURVEY §8a row a12, each from the reference's own code."""
    A = np.asarray
    dsp, extract, extract_methods = R["dsp"], R["extract"], R["extract_methods"]
    g = {}
    # S3Gen mel (codec/models/s3gen/mel.py:25-100)
    s3 = extract("mlx_audio/codec/models/s3gen/mel.py", ["_reflect_pad_2d", "mel_spectrogram"])
    x = np.stack([synth(500, 12000, 24000), 0.3 * synth(501, 12000, 24000)])
    g["s3gen|x"], g["s3gen|y"] = x, A(s3.mel_spectrogram(mx.array(x)))
    g["s3gen|y1d"] = A(s3.mel_spectrogram(mx.array(x[1])))
    # IndexTTS (tts/models/indextts/mel.py)
    spec = importlib.util.spec_from_file_location("mlx_audio.tts_indextts_mel", os.path.join(REF, "mlx_audio/tts/models/indextts/mel.py"))
    it = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(it)
    x = synth(502, 9000, 24000)
    g["indextts|x"], g["indextts|y"] = x, A(it.log_mel_spectrogram(mx.array(x)))
    g["indextts|ypad"] = A(it.log_mel_spectrogram(x, padding=500))
    # Spark BiCodec (tts/models/spark/bicodec.py:20-49)
    sp = extract("mlx_audio/tts/models/spark/bicodec.py", ["mel_spectrogram"], {"Optional": tuple})
    x = synth(503, 16000)
    g["spark|x"], g["spark|y"] = x, A(sp.mel_spectrogram(mx.array(x)))
    # Chatterbox voice encoder (tts/models/chatterbox/voice_encoder/melspec.py:13-77)
    cfg = extract("mlx_audio/tts/models/chatterbox/voice_encoder/config.py", ["VoiceEncConfig"],
                  {"dataclass": __import__("dataclasses").dataclass})
    ve = extract("mlx_audio/tts/models/chatterbox/voice_encoder/melspec.py", ["melspectrogram"], {"VoiceEncConfig": cfg.VoiceEncConfig})
    x = np.stack([synth(504, 8000), 0.5 * synth(505, 8000)])
    g["ve|x"] = x
    g["ve|amp"] = A(ve.melspectrogram(mx.array(x), cfg.VoiceEncConfig()))
    g["ve|amp1d"] = A(ve.melspectrogram(mx.array(x[0]), cfg.VoiceEncConfig()))
    g["ve|db_norm"] = A(ve.melspectrogram(mx.array(x), cfg.VoiceEncConfig(mel_power=1.0, mel_type="db", normalized_mels=True)))
    g["ve|db"] = A(ve.melspectrogram(mx.array(x), cfg.VoiceEncConfig(mel_type="db")))
    # Soprano decoder head (tts/models/soprano/decoder.py:14-49), n_fft 2048 / hop 512
    so = extract("mlx_audio/tts/models/soprano/decoder.py", ["ISTFTHead"])
    lin = (0.5 * np.random.default_rng(15).standard_normal((1, 12, 2050))).astype(np.float32)
    g["soprano|head_in"], g["soprano|head_out"] = lin, A(so.ISTFTHead(8, 2048, 512)(mx.array(lin)))
    # GLM-ASR (stt/models/glmasr/glmasr.py:547-589)
    glm = extract_methods("mlx_audio/stt/models/glmasr/glmasr.py", "Model", ["_preprocess_audio"])
    me = types.SimpleNamespace(sample_rate=16000, config=types.SimpleNamespace(whisper_config=types.SimpleNamespace(num_mel_bins=128)))
    x = synth(506, 16000 * 2 + 40)
    g["glmasr|x"], g["glmasr|y"] = x, A(glm._preprocess_audio(me, x))
    # Smart-Turn (vad/models/smart_turn/smart_turn.py:158-229)
    st = extract_methods("mlx_audio/vad/models/smart_turn/smart_turn.py", "Model", ["_prepare_audio_array", "prepare_input_features"],
                         {"log_mel_spectrogram": R["whisper"].log_mel_spectrogram})
    pc = types.SimpleNamespace(sampling_rate=16000, max_audio_seconds=2, n_fft=400, hop_length=160, n_mels=80, normalize_audio=True)
    me = types.SimpleNamespace(config=types.SimpleNamespace(processor_config=pc), dtype=mx.float32, _resample=lambda a, s, t: a)
    me._prepare_audio_array = lambda audio, sample_rate=None: st._prepare_audio_array(me, audio, sample_rate=sample_rate)
    for name, n in (("short", 20000), ("long", 40000)):
        x = synth(507, n) + 0.05
        g[f"smart|{name}|x"], g[f"smart|{name}|y"] = x, A(st.prepare_input_features(me, x))
    # LFM2 audio preprocessor (sts/models/lfm_audio/processor.py:34-140), dither 0
    lf = extract("mlx_audio/sts/models/lfm_audio/processor.py", ["AudioPreprocessor"], {"PreprocessorConfig": object})
    pcfg = types.SimpleNamespace(sample_rate=16000, normalize="per_feature", window_size=0.025, window_stride=0.01, window="hann",
                                 features=128, n_fft=512, log=True, dither=0.0, preemph=0.97)
    x = np.stack([synth(508, 16000 + 53), 0.2 * synth(509, 16000 + 53)])
    g["lfm2|x"], g["lfm2|y"] = x, A(lf.AudioPreprocessor(pcfg)(mx.array(x)))
    g["lfm2|y1d"] = A(lf.AudioPreprocessor(pcfg)(mx.array(x[1])))
    # LFM2 detokenizer iSTFT (sts/models/lfm_audio/detokenizer.py:468-507), n_fft 1280 / hop 320
    dt = extract_methods("mlx_audio/sts/models/lfm_audio/detokenizer.py", "LFM2AudioDetokenizer", ["_istft"])
    rng = np.random.default_rng(16)
    mag = np.exp(0.5 * rng.standard_normal((2, 9, 641))).astype(np.float32)
    ph = rng.uniform(-np.pi, np.pi, mag.shape).astype(np.float32)
    w = A(dsp.hanning(1280, True))
    me = types.SimpleNamespace(n_fft=1280, hop_length=320, window=mx.array(w))
    g["lfm2|mag"], g["lfm2|phase"], g["lfm2|w"], g["lfm2|wave"] = mag, ph, w, A(dt._istft(me, mx.array(mag), mx.array(ph)))
    # MossFormer2-SE chunk (sts/models/mossformer2_se/model.py:396-428): stft(center=False) -> (F, T) planes -> ISTFTCache.istft
    x = (synth(510, 9600, 48000) * 8000.0).astype(np.float32)
    w = dsp.hamming(1920, periodic=False)
    sc = dsp.stft(mx.array(x), 1920, 384, 1920, w, center=False)
    re, im = A(mx.real(sc).T), A(mx.imag(sc).T)
    g["moss|x"], g["moss|re"], g["moss|im"] = x, re, im
    cache = dsp.ISTFTCache()
    g["moss|y"] = A(cache.istft(mx.array(re).reshape(1, *re.shape), mx.array(im).reshape(1, *im.shape), 1920, 384, 1920, w,
                                center=False, audio_length=9600))[0]
    # Chatterbox-Turbo HiFT (tts/models/chatterbox_turbo/models/s3gen/hifigan.py:418-537): un-centred _stft, S3Gen-style _istft
    ct = extract_methods("mlx_audio/tts/models/chatterbox_turbo/models/s3gen/hifigan.py", "HiFTGenerator", ["_stft", "_istft"],
                         {"Tuple": tuple})
    w16 = A(dsp.hanning(16, True))
    me = types.SimpleNamespace(istft_params={"n_fft": 16, "hop_len": 4}, stft_window=mx.array(w16))
    x = np.stack([synth(511, 1203, 24000), 0.4 * synth(512, 1203, 24000)])
    re, im = ct._stft(me, mx.array(x))
    g["cturbo|x"], g["cturbo|re"], g["cturbo|im"] = x, A(re), A(im)
    re, im = ct._stft(me, mx.array(x[:, :9]))  # shorter than n_fft: one zero-extended frame
    g["cturbo|short|re"], g["cturbo|short|im"] = A(re), A(im)
    rng = np.random.default_rng(17)
    mag = np.exp(rng.normal(0, 1.5, (2, 9, 120))).astype(np.float32)
    mag[0, 3, 4] = 300.0
    ph = rng.uniform(-np.pi, np.pi, mag.shape).astype(np.float32)
    g["cturbo|mag"], g["cturbo|phase"], g["cturbo|y"] = mag, ph, A(ct._istft(me, mx.array(mag), mx.array(ph)))
    # CAMPPlus x-vector front-end (codec/models/s3gen/xvector.py:12-150)
    xv = extract("mlx_audio/codec/models/s3gen/xvector.py", ["_povey_window", "_next_power_of_2", "kaldi_fbank"])
    x = synth(513, 16000 + 123) + 0.02
    g["xvector|x"], g["xvector|y"] = x, A(xv.kaldi_fbank(mx.array(x)))
    g["xvector|y40"] = A(xv.kaldi_fbank(mx.array(x[None, :4000]), num_mel_bins=40))
    np.savez_compressed(os.path.join(OUT, "refshim_variants.npz"), **g)


def main():
    mx, R = _load_reference()
    if "--variants-only" in sys.argv:
        variants_goldens(mx, R)
        return
    if "--hift-only" in sys.argv:
        hift_goldens(mx, R)
        return
    if "--kaldi-only" in sys.argv:
        kaldi_goldens(mx, R)
        return
    dsp = R["dsp"]
    os.makedirs(OUT, exist_ok=True)
    A = lambda v: np.asarray(v)

    # ---- windows + filterbanks ---------------------------------------------------------
    g = {}
    for kind in ("hanning", "hamming", "blackman", "bartlett"):
        for size in (16, 20, 21, 400, 401, 1024):
            for per in (False, True):
                g[f"win|{kind}|{size}|{int(per)}"] = A(getattr(dsp, kind)(size, per))
    fb_cases = {
        "whisper80": (16000, 400, 80, 0, None, "slaney", None),
        "whisper128": (16000, 400, 128, 0, None, "slaney", None),
        "parakeet80": (16000, 512, 80, 0, None, "per_feature", None),
        "vocos100": (24000, 1024, 100, 0, None, None, "htk"),
        "qwen3tts128": (24000, 1024, 128, 0.0, 12000.0, "slaney", "slaney"),
        "voxtral128": (16000, 400, 128, 0, 8000, "slaney", "slaney"),
        "funasr80": (16000, 400, 80, 0, None, "slaney", "htk"),
        "s3gen80": (24000, 1920, 80, 0, 8000, "slaney", "slaney"),
        "spark128": (16000, 1024, 128, 10, 8000, "slaney", "slaney"),
    }
    for k, a in fb_cases.items():
        g[f"fb|{k}"] = A(dsp.mel_filters(*a))
    np.savez_compressed(os.path.join(OUT, "refshim_tables.npz"), **g)

    # ---- stft --------------------------------------------------------------------------
    g = {}
    stft_cases = [
        # name, L, n_fft, hop, win, window(str|("arr",fn,size,periodic)), center, pad_mode
        ("w400", 4000, 400, 160, None, ("arr", "hanning", 400, False), True, "reflect"),
        ("p512", 4000, 512, 160, 400, "hann", True, "reflect"),
        ("v1024", 6000, 1024, None, 256, ("arr", "hanning", 1024, False), True, "reflect"),
        ("k20", 600, 20, 5, 20, "hann", True, "reflect"),
        ("h16", 333, 16, 4, 16, ("arr", "hanning", 17, False, -1), True, "reflect"),
        ("nc1920", 9600, 1920, 384, 1920, ("arr", "hamming", 1920, False), False, "reflect"),
        ("const512", 3000, 512, 160, 400, "hamming", True, "constant"),
        ("def800", 5000, 800, None, None, "hann", True, "reflect"),
        ("black300", 2000, 300, 75, 200, "blackman", True, "reflect"),
        ("bart64", 999, 64, 16, 64, "bartlett", True, "constant"),
        ("edge201", 201, 400, 160, None, "hann", True, "reflect"),
        ("edge_nc", 400, 400, 160, None, "hann", False, "reflect"),
        ("odd_hop", 3001, 400, 161, None, "hann", True, "reflect"),
    ]

    def mkwin(spec):
        if isinstance(spec, str):
            return spec
        fn = getattr(dsp, spec[1])
        w = fn(spec[2], spec[3])
        if len(spec) > 4:
            w = w[: spec[4]]
        return w

    for i, (name, L, n_fft, hop, win, wspec, center, pad_mode) in enumerate(stft_cases):
        x = synth(100 + i, L)
        y = dsp.stft(mx.array(x), n_fft, hop, win, mkwin(wspec), center, pad_mode)
        g[f"stft|{name}|x"] = x
        g[f"stft|{name}|y"] = A(y).astype(np.complex64)
    np.savez_compressed(os.path.join(OUT, "refshim_stft.npz"), **g)

    # ---- istft -------------------------------------------------------------------------
    g = {}
    rng = np.random.default_rng(7)

    def rspec(F, T):
        return (rng.standard_normal((F, T)) + 1j * rng.standard_normal((F, T))).astype(np.complex64)

    istft_cases = [
        # name, F, T, hop, win, window, center, length, normalized
        ("k20", 11, 200, 5, 20, "hann", True, None, False),
        ("h16n", 9, 150, 4, 16, "hann", True, None, True),
        ("v1024", 513, 24, 256, 1024, ("arr", "hanning", 1024, False), True, None, False),
        ("s2048", 1025, 7, 512, 2048, ("arr", "hanning", 2048, False), True, None, False),
        ("e1280", 641, 9, 320, 1280, ("arr", "hanning", 1280, False), True, None, False),
        ("nc_norm", 33, 40, 16, 64, "hamming", False, None, True),
        ("len", 33, 40, 16, 64, "hann", True, 500, False),
        ("odd21", 11, 50, 5, 21, ("arr", "hanning", 21, False), True, None, False),
        ("short_w", 33, 30, 16, 64, ("arr", "hanning", 48, False), True, None, False),
        ("black", 51, 33, 25, 100, "blackman", True, None, False),
    ]
    for name, F, T, hop, win, wspec, center, length, normalized in istft_cases:
        s = rspec(F, T)
        if name == "odd21":
            # irfft of 11 bins gives 20 samples; a 21-tap window cannot broadcast — the reference
            # raises a shape error here.  Record that instead of values.
            try:
                dsp.istft(mx.array(s), hop, win, mkwin(wspec), center, length, normalized)
                g["istft|odd21|raises"] = np.array(0)
            except Exception:
                g["istft|odd21|raises"] = np.array(1)
            continue
        y = dsp.istft(mx.array(s), hop, win, mkwin(wspec), center, length, normalized)
        g[f"istft|{name}|x"] = s
        g[f"istft|{name}|y"] = A(y).astype(np.float32)

    cache = dsp.ISTFTCache()
    for name, B, n_fft, hop, T, wkind, center, alen in [
        ("c64", 3, 64, 16, 50, "hamming", True, 700),
        ("c1920", 2, 1920, 384, 6, "hamming", True, None),
        ("c20nc", 4, 20, 5, 120, "hanning", False, None),
    ]:
        F = n_fft // 2 + 1
        re = rng.standard_normal((B, F, T)).astype(np.float32)
        im = rng.standard_normal((B, F, T)).astype(np.float32)
        w = getattr(dsp, wkind)(n_fft, False)
        y = cache.istft(mx.array(re), mx.array(im), n_fft, hop, n_fft, w, center, alen)
        g[f"icache|{name}|re"], g[f"icache|{name}|im"], g[f"icache|{name}|y"] = re, im, A(y).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "refshim_istft.npz"), **g)

    # ---- model front-ends --------------------------------------------------------------
    g = {}
    x = synth(200, 16000)
    g["whisper|x"] = x
    g["whisper|80"] = A(R["whisper"].log_mel_spectrogram(x, n_mels=80))
    g["whisper|128"] = A(R["whisper"].log_mel_spectrogram(mx.array(x), n_mels=128))
    g["whisper|80pad"] = A(R["whisper"].log_mel_spectrogram(x, n_mels=80, padding=8000))
    xs = x.copy()
    xs[5000:] = 0.0  # long digital silence -> the max-8 clamp is active
    g["whisper|sil|x"] = xs
    g["whisper|sil|80"] = A(R["whisper"].log_mel_spectrogram(xs, n_mels=80))

    PA = R["parakeet"].PreprocessArgs
    pa = PA(sample_rate=16000, normalize="per_feature", window_size=0.025, window_stride=0.01,
            window="hann", features=80, n_fft=512, dither=1e-5)
    x = synth(201, 24000)
    g["parakeet|x"] = x
    g["parakeet|pf"] = A(R["parakeet"].log_mel_spectrogram(mx.array(x), pa))
    pa2 = PA(sample_rate=16000, normalize="all_features", window_size=0.025, window_stride=0.01,
             window="hamming", features=64, n_fft=512, dither=0.0, pad_to=30000, pad_value=0.0, preemph=0.0)
    g["parakeet|global"] = A(R["parakeet"].log_mel_spectrogram(mx.array(x), pa2))

    x = synth(202, 12000)
    fbv = R["voxtral"].compute_mel_filters()
    g["voxtral|x"] = x
    g["voxtral|y"] = A(R["voxtral"].compute_mel_spectrogram(mx.array(x), mx.array(fbv)))

    x = synth(203, 12000, sr=24000)
    g["vocos|x"] = x
    g["vocos|mel"] = A(R["vocos_mel"].log_mel_spectrogram(mx.array(x)))
    head = R["vocos"].ISTFTHead(8, 1024, 256)
    lin = (0.5 * np.random.default_rng(5).standard_normal((1, 20, 1026))).astype(np.float32)
    g["vocos|head_in"] = lin
    g["vocos|head_out"] = A(head(mx.array(lin)))
    z = np.zeros(120000, np.float32)
    mel0 = A(R["vocos_mel"].log_mel_spectrogram(mx.array(z)))
    g["vocos|zeros_mel_shape"] = np.array(mel0.shape)
    head_z = R["vocos"].ISTFTHead(8, 1024, 256)
    g["vocos|zeros_audio_len"] = np.array(A(head_z(mx.array(np.zeros((1, mel0.shape[1], 1026), np.float32)))).shape)

    K = R["kokoro"].MLXSTFT(filter_length=20, hop_length=5, win_length=20)
    x = np.stack([synth(204, 1500, 24000), synth(205, 1500, 24000)])
    mag, ph = K.transform(mx.array(x))
    g["kokoro|x"], g["kokoro|mag"], g["kokoro|phase"] = x, A(mag), A(ph)
    rng2 = np.random.default_rng(9)
    m2 = np.minimum(np.exp(0.5 * rng2.standard_normal((2, 11, 301))), 1e2).astype(np.float32)
    p2 = np.sin(rng2.standard_normal((2, 11, 301))).astype(np.float32)
    g["kokoro|inv_mag"], g["kokoro|inv_phase"] = m2, p2
    g["kokoro|inv_y"] = A(K.inverse(mx.array(m2), mx.array(p2)))
    p3 = rng2.uniform(-math.pi, math.pi, (1, 11, 301)).astype(np.float32)
    g["kokoro|inv_phase_wrapped"] = p3
    g["kokoro|inv_y_wrapped"] = A(K.inverse(mx.array(m2[:1]), mx.array(p3)))
    g["kokoro|unwrap"] = A(R["kokoro"].mlx_unwrap(mx.array(p3[0]), axis=1))

    np.random.seed(42)  # the reference test's own input: tts/tests/test_qwen3_tts.py:164-167
    xq = np.random.randn(12000).astype(np.float32)
    g["qwen3|y"] = A(R["qwen3"].mel_spectrogram(mx.array(xq)))
    t = np.arange(12000, dtype=np.float32) / 24000.0
    g["qwen3|sine_y"] = A(R["qwen3"].mel_spectrogram(mx.array(np.sin(2 * np.pi * 1000 * t).astype(np.float32))))

    x = synth(206, 8000)
    g["s3|x"] = x
    g["s3|y"] = A(R["s3tok"].log_mel_spectrogram(mx.array(x)))
    xb = np.stack([synth(207, 8000), 0.05 * synth(208, 8000)])
    g["s3|xb"] = xb
    g["s3|compat"] = A(R["s3tok"].log_mel_spectrogram_compat(mx.array(xb), n_mels=80))

    x = np.stack([synth(209, 9000), synth(210, 9000)])
    g["sortformer|x"] = x
    g["sortformer|y"] = A(R["sortformer"].extract_mel_features(mx.array(x)))
    np.savez_compressed(os.path.join(OUT, "refshim_models.npz"), **g)
    hift_goldens(mx, R)
    kaldi_goldens(mx, R)
    post_goldens(mx, R)
    variants_goldens(mx, R)

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
