#!/usr/bin/env python
"""Per-config measurements for BASELINE.json configs C1, C3, C4, C5 (C2 is bench.py's headline).

For every config: device-resident synthetic input, >= 3 warm-ups, CUDA-event timing on the
launch stream, audio-hours/s, and achieved algorithmic HBM GB/s as a fraction of MEASURED_PEAKS.json.
Prints one JSON line per config and writes them to the path given by --out."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mlx_audio_plus_b200 import _lib as L  # noqa: E402
from mlx_audio_plus_b200._arrays import Ingested  # noqa: E402
from mlx_audio_plus_b200.dsp import hanning, mel_filters  # noqa: E402
from mlx_audio_plus_b200.frontend import FrontendPlan, IstftPlan  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def timeit(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def synth(B, n, sr, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    t = torch.arange(n, device="cuda", dtype=torch.float64) / sr
    tone = (0.2 * (torch.sin(2 * np.pi * 440 * t) + torch.sin(2 * np.pi * 3000 * t))).float()
    x = torch.empty((B, n), dtype=torch.float32, device="cuda")
    for c0 in range(0, B, 64):
        c1 = min(B, c0 + 64)
        scale = (0.5 + (torch.arange(c0, c1, device="cuda") % 7).float() / 7)[:, None]
        x[c0:c1] = (0.1 * torch.randn((c1 - c0, n), generator=g, device="cuda") + tone[None]) * scale
    return x


def fwd_case(name, plan, x, sr, steps, length=None):
    B, n = x.shape
    ing = Ingested("torch", True, x, None, x.device)
    out = plan.run(ing, length=length)
    ms = timeit(lambda: plan.run(ing, length=length), steps)
    by = x.numel() * 4 + out.numel() * out.element_size()
    hours = B * n / sr / 3600.0
    return {"config": name, "kernel": plan.kernel_name, "batch": B, "samples": n, "ms": ms,
            "audio_hours_per_s": hours / (ms * 1e-3), "algorithmic_GBps": by / (ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": by / (ms * 1e-3) / 1e9 / peak(), "out_shape": list(out.shape)}


def inv_case(name, plan, spec, sr, steps):
    B, F, T = spec.shape
    ing = Ingested("torch", True, spec, None, spec.device)
    out = plan.run(ing)
    ms = timeit(lambda: plan.run(ing), steps)
    by = spec.numel() * 8 + out.numel() * 4
    hours = out.numel() / sr / 3600.0
    return {"config": name, "kernel": plan.kernel_name, "batch": B, "frames": T, "ms": ms,
            "audio_hours_per_s": hours / (ms * 1e-3), "algorithmic_GBps": by / (ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": by / (ms * 1e-3) / 1e9 / peak(), "out_shape": list(out.shape)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--match", default="", help="section G: run only the shapes whose name contains this text")
    a = ap.parse_args()
    res = []

    def want(k):
        return not a.only or k in a.only.split(",")

    whisper = dict(n_fft=400, hop=160, window=np.asarray(hanning(400)), drop_last=True, spec_kind=L.SPEC_POWER,
                   log_kind=L.LOG_LOG10, guard_kind=L.GUARD_MAX, guard_eps=1e-10, clamp_kind=L.CLAMP_CLIP_MAX,
                   clamp_value=8.0, affine_add=4.0, affine_div=4.0)
    if want("C1"):
        plan = FrontendPlan(filterbank=np.asarray(mel_filters(16000, 400, 80, norm="slaney", mel_scale=None)), **whisper)
        res.append(fwd_case("C1 whisper-80, one 30 s clip (latency)", plan, synth(1, 480000, 16000, 1234), 16000, 50))
    if want("C3"):
        plan = FrontendPlan(n_fft=512, hop=160, window=np.asarray(hanning(400)), preemph=0.97, spec_kind=L.SPEC_POWER,
                            filterbank=np.asarray(mel_filters(16000, 512, 80, norm="per_feature", mel_scale=None)),
                            log_kind=L.LOG_LN, guard_kind=L.GUARD_ADD, guard_eps=1e-5, norm_kind=L.NORM_PER_FEATURE,
                            norm_ddof=0, norm_eps=1e-5)
        res.append(fwd_case("C3 parakeet, one 1-hour file", plan, synth(1, 57_600_000, 16000, 1236), 16000, a.steps))
        res.append(fwd_case("C3 parakeet, 16 x 1-hour files", plan, synth(16, 57_600_000, 16000, 1236), 16000, 3))
    if want("C5"):
        plan = FrontendPlan(n_fft=1024, hop=256, window=np.asarray(hanning(1024)), drop_last=True, spec_kind=L.SPEC_MAGNITUDE,
                            filterbank=np.asarray(mel_filters(24000, 1024, 100, norm=None, mel_scale="htk")),
                            log_kind=L.LOG_LN, guard_kind=L.GUARD_MAX, guard_eps=1e-5)
        for B in (1, 64, 1024, 8192):
            res.append(fwd_case(f"C5 vocos mel forward, B={B} x 5 s", plan, synth(B, 120000, 24000, 1238), 24000, a.steps))
        iplan = IstftPlan(n_fft=1024, hop=256, window=np.asarray(hanning(1024)), center=True)
        g = torch.Generator(device="cuda")
        g.manual_seed(7)
        for B in (1, 64, 1024):
            mag = torch.exp(0.5 * torch.randn((B, 513, 468), generator=g, device="cuda")).clamp(max=1e2)
            ph = torch.randn((B, 513, 468), generator=g, device="cuda")
            spec = torch.complex(mag * torch.cos(ph), mag * torch.sin(ph)).contiguous()
            res.append(inv_case(f"C5 vocos istft head, B={B} x (513,468)", iplan, spec, 24000, a.steps))
    if want("C4"):
        iplan = IstftPlan(n_fft=20, hop=5, window=np.asarray(hanning(21)[:-1]), center=True)
        g = torch.Generator(device="cuda")
        g.manual_seed(9)
        for B in (1, 64, 1024):
            mag = torch.exp(0.5 * torch.randn((B, 11, 24001), generator=g, device="cuda")).clamp(max=1e2)
            ph = torch.sin(torch.randn((B, 11, 24001), generator=g, device="cuda"))
            spec = torch.complex(mag * torch.cos(ph), mag * torch.sin(ph)).contiguous()
            res.append(inv_case(f"C4 kokoro istft, B={B} x (11,24001)", iplan, spec, 24000, a.steps))
    if want("S"):
        # dsp.stft itself (complex64 (T, F) rows): bytes = samples in + 8 * T * F out.  torch.stft (cuFFT, same
        # framing) is timed beside it for context only.
        for nm, n_fft, hop, B, n, sr in (("400/160, 1024 x 30 s", 400, 160, 1024, 480000, 16000),
                                         ("512/160, 4 x 1 h", 512, 160, 4, 57_600_000, 16000),
                                         ("1024/256, 4096 x 5 s", 1024, 256, 4096, 120000, 24000),
                                         ("800/200 (the module defaults), 1024 x 30 s", 800, 200, 1024, 480000, 16000)):
            w = np.asarray(hanning(n_fft))
            x = synth(B, n, sr, 1240)
            plan = FrontendPlan(n_fft=n_fft, hop=hop, window=w, spec_kind=L.SPEC_COMPLEX)
            r = fwd_case(f"S dsp.stft {nm}", plan, x, sr, a.steps)
            tw = torch.from_numpy(w).cuda()
            r["torch_stft_ms"] = timeit(lambda: torch.stft(x, n_fft, hop, n_fft, tw, center=True, pad_mode="reflect",
                                                           return_complex=True), max(2, a.steps // 2))
            res.append(r)
            del x
            torch.cuda.empty_cache()
    if want("R"):
        # the step in front of the path: decoded PCM -> resample -> mono float32, one kernel.
        # bytes = interleaved PCM in + float32 mono out; scipy (the reference's own call) timed on one host core beside it.
        from mlx_audio_plus_b200.stt.utils import load_audio
        for nm, orig, ch, secs in (("44.1 kHz stereo int16 -> 16 kHz mono, 1 h", 44100, 2, 3600),
                                   ("48 kHz mono int16 -> 16 kHz, 1 h", 48000, 1, 3600)):
            n = orig * secs
            g = torch.Generator(device="cuda")
            g.manual_seed(11)
            pcm = (torch.randn((n, ch), generator=g, device="cuda") * 3000).to(torch.int16)
            out = load_audio(pcm=pcm, sample_rate=orig, sr=16000)
            ms = timeit(lambda: load_audio(pcm=pcm, sample_rate=orig, sr=16000), a.steps)
            by = pcm.numel() * 2 + out.numel() * 4
            r = {"config": f"R load_audio {nm}", "kernel": "resample_kernel", "batch": 1, "samples": n, "ms": ms,
                 "audio_hours_per_s": secs / 3600.0 / (ms * 1e-3), "algorithmic_GBps": by / (ms * 1e-3) / 1e9,
                 "frac_of_hbm_peak": by / (ms * 1e-3) / 1e9 / peak(), "out_shape": list(out.shape)}
            try:  # the reference's own call (stt/utils.py:21-29: scipy.signal.resample_poly per channel, then the channel mean)
                import time
                from math import gcd

                from scipy.signal import resample_poly
                sub = pcm[: orig * 60].cpu().numpy().astype(np.float32) / 32768.0
                g_ = gcd(orig, 16000)
                t0 = time.perf_counter()
                resample_poly(sub, 16000 // g_, orig // g_, axis=0, padtype="edge").mean(axis=1)
                r["scipy_cpu_1core_audio_hours_per_s"] = (60 / 3600.0) / (time.perf_counter() - t0)
            except Exception as e:  # noqa: BLE001
                r["scipy_cpu_1core_audio_hours_per_s"] = str(e)
            res.append(r)
            del pcm
            torch.cuda.empty_cache()
    if want("W"):
        # Whisper as the reference's transcription loop calls it: padding = N_SAMPLES zero samples behind every clip
        # (all-padding frames are filled, not transformed) and the encoder's float16 straight from the epilogue
        from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram as wmel
        x = synth(1024, 480000, 16000, 1242)
        for nm, kw in (("padding=0, float32", dict(padding=0)), ("padding=N_SAMPLES, float32", dict(padding=480000)),
                       ("padding=N_SAMPLES, float16 epilogue", dict(padding=480000, dtype="float16"))):
            out = wmel(x, n_mels=128, **kw)
            ms = timeit(lambda: wmel(x, n_mels=128, **kw), a.steps)
            by = x.numel() * 4 + out.numel() * out.element_size()
            res.append({"config": f"W whisper-128 1024 x 30 s, {nm}", "kernel": "fast_logmel_400x160 (+ const_rows)", "batch": 1024,
                        "ms": ms, "audio_hours_per_s": 1024 * 30 / 3600.0 / (ms * 1e-3), "algorithmic_GBps": by / (ms * 1e-3) / 1e9,
                        "frac_of_hbm_peak": by / (ms * 1e-3) / 1e9 / peak(), "out_shape": list(out.shape)})
        del x
        torch.cuda.empty_cache()
    if want("P"):
        # the steps right after the path: Whisper (B, 3000, 128) float32 -> float16 segments; FunASR LFR + CMVN
        from mlx_audio_plus_b200._post import lfr, rows_pad_cast
        mel = torch.randn((1024, 3000, 128), device="cuda")
        out = rows_pad_cast(mel, 0, 3000, 3000, "float16")
        ms = timeit(lambda: rows_pad_cast(mel, 0, 3000, 3000, "float16"), a.steps)
        by = mel.numel() * 4 + out.numel() * 2
        res.append({"config": "P whisper segments 1024 x (3000,128) f32 -> f16", "kernel": "rows_pad_cast_kernel", "batch": 1024, "ms": ms,
                    "audio_hours_per_s": 1024 * 30 / 3600.0 / (ms * 1e-3), "algorithmic_GBps": by / (ms * 1e-3) / 1e9,
                    "frac_of_hbm_peak": by / (ms * 1e-3) / 1e9 / peak(), "out_shape": list(out.shape)})
        feats = torch.randn((64, 360000, 80), device="cuda")
        sh, sc = torch.randn(560, device="cuda"), torch.rand(560, device="cuda") + 0.5
        out = lfr(feats, 7, 6, sh, sc)
        ms = timeit(lambda: lfr(feats, 7, 6, sh, sc), a.steps)
        by = feats.numel() * 4 + out.numel() * 4
        res.append({"config": "P funasr LFR 7/6 + CMVN, 64 x 1 h of (T,80)", "kernel": "lfr_kernel", "batch": 64, "ms": ms,
                    "audio_hours_per_s": 64.0 / (ms * 1e-3), "algorithmic_GBps": by / (ms * 1e-3) / 1e9,
                    "frac_of_hbm_peak": by / (ms * 1e-3) / 1e9 / peak(), "out_shape": list(out.shape)})
        del mel, feats
        torch.cuda.empty_cache()
    if want("V"):
        # SURVEY 8a row a12 callers and the Hugging Face extractor of Qwen3-ASR: (M, T)-layout / run-time-table instances
        from mlx_audio_plus_b200._post import cmvn_utterance
        from mlx_audio_plus_b200.codec.models.s3gen.mel import mel_spectrogram as s3gen_mel
        from mlx_audio_plus_b200.codec.models.s3tokenizer.utils import log_mel_spectrogram_compat as s3_compat
        from mlx_audio_plus_b200.sts.models.lfm_audio.processor import AudioPreprocessor, PreprocessorConfig
        from mlx_audio_plus_b200.stt.models.qwen3_asr.feature_extractor import WhisperFeatureExtractor
        from mlx_audio_plus_b200.tts.models.spark.bicodec import mel_spectrogram as spark_mel
        x = synth(1024, 480000, 16000, 1243)
        fe = WhisperFeatureExtractor(feature_size=128)
        lfm = AudioPreprocessor(PreprocessorConfig(dither=0.0))
        cases = [
            ("V qwen3-asr HF WhisperFeatureExtractor 1024 x 30 s -> (B,128,3000)", "fast_logmel_400x160 (generated mel, (M,T) write-out)",
             lambda: fe(x, sampling_rate=16000, padding=True, truncation=False, return_attention_mask=True,
                        return_tensors="cuda")["input_features"]),
            ("V s3tokenizer compat 1024 x 30 s -> (B,128,3000), one max over the batch", "fast_logmel_400x160 (generated mel, (M,T) write-out)",
             lambda: s3_compat(x, 128)),
            ("V lfm2 preprocessor 1024 x 30 s -> (B,3001,128), valid-frame statistics", "fast_logmel_512x160 + normalise",
             lambda: lfm(x)),
            ("V spark mel 1024 x 30 s (1024/320, 128 mel, magnitude)", "fast_logmel_1024x320", lambda: spark_mel(x)),
        ]
        for name, kern, fn in cases:
            out = fn()
            ms = timeit(fn, a.steps)
            by = x.numel() * 4 + out.numel() * out.element_size()
            res.append({"config": name, "kernel": kern, "batch": 1024, "ms": ms, "audio_hours_per_s": 1024 * 30 / 3600.0 / (ms * 1e-3),
                        "algorithmic_GBps": by / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": by / (ms * 1e-3) / 1e9 / peak(),
                        "out_shape": list(out.shape)})
        del x
        torch.cuda.empty_cache()
        x24 = synth(1024, 240000, 24000, 1244)
        out = s3gen_mel(x24)
        ms = timeit(lambda: s3gen_mel(x24), a.steps)
        by = x24.numel() * 4 + out.numel() * 4
        res.append({"config": "V s3gen mel 1024 x 10 s (1920/480, 80 mel) -> (B,80,T)", "kernel": "frontend_generic_kernel", "batch": 1024, "ms": ms,
                    "audio_hours_per_s": 1024 * 10 / 3600.0 / (ms * 1e-3), "algorithmic_GBps": by / (ms * 1e-3) / 1e9,
                    "frac_of_hbm_peak": by / (ms * 1e-3) / 1e9 / peak(), "out_shape": list(out.shape)})
        del x24
        feats = torch.randn((64, 60000, 560), device="cuda")  # 64 x 1 h of LFR features
        out = cmvn_utterance(feats)
        ms = timeit(lambda: cmvn_utterance(feats), a.steps)
        by = feats.numel() * 4 * 3  # read for the statistics, read + write for the apply pass
        res.append({"config": "V funasr per-utterance CMVN, 64 x 1 h of (T/6,560)", "kernel": "cmvn_stats_kernel + cmvn_apply_kernel", "batch": 64,
                    "ms": ms, "audio_hours_per_s": 64.0 / (ms * 1e-3), "algorithmic_GBps": by / (ms * 1e-3) / 1e9,
                    "frac_of_hbm_peak": by / (ms * 1e-3) / 1e9 / peak(), "out_shape": list(out.shape)})
        del feats
        torch.cuda.empty_cache()
    if want("G"):
        # shapes the generic (any-n_fft) kernels serve: S3Gen mel 1920/480, Soprano head 2048/512, LFM2 detokenizer 1280/320,
        # MossFormer2-SE chunk stft / istft 1920/384, Kaldi fbank (512-point after zero extension) — SURVEY 8a rows a10 / a12
        from mlx_audio_plus_b200.codec.models.s3gen.mel import mel_spectrogram as s3gen_mel
        from mlx_audio_plus_b200.dsp import compute_fbank_kaldi
        from mlx_audio_plus_b200.sts.models.lfm_audio.detokenizer import istft_same
        from mlx_audio_plus_b200.tts.models.soprano.decoder import ISTFTHead as SopranoHead
        g = torch.Generator(device="cuda")
        g.manual_seed(21)

        def rec(name, kern, fn, in_bytes, secs):
            if a.match and a.match not in name:
                return
            out = fn()
            ms = timeit(fn, a.steps)
            by = in_bytes + out.numel() * out.element_size()
            res.append({"config": name, "kernel": kern, "ms": ms, "audio_hours_per_s": secs / 3600.0 / (ms * 1e-3),
                        "algorithmic_GBps": by / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": by / (ms * 1e-3) / 1e9 / peak(), "out_shape": list(out.shape)})

        x24 = synth(1024, 240000, 24000, 1244)
        rec("G s3gen mel 1024 x 10 s (1920/480, 80 mel) -> (B,80,T)", "forward 1920/480", lambda: s3gen_mel(x24), x24.numel() * 4, 1024 * 10)
        del x24
        xl = torch.randn((64, 469, 2050), generator=g, device="cuda") * 0.5
        sh = SopranoHead(512, 2048, 512)
        rec("G soprano head 64 x (469, 2050) -> 2048/512 istft", "inverse 2048/512 (polar, log-magnitude)", lambda: sh(xl)[0], xl.numel() * 4, 64 * 468 * 512 / 32000)
        del xl
        mag = torch.exp(0.3 * torch.randn((256, 375, 641), generator=g, device="cuda"))
        ph = torch.randn((256, 375, 641), generator=g, device="cuda")
        w = torch.hann_window(1280, device="cuda")
        rec("G lfm2 detokenizer 256 x (375, 641) -> 1280/320 istft", "inverse 1280/320 (polar)", lambda: istft_same(mag, ph, w, 1280, 320), 2 * mag.numel() * 4, 256 * 375 * 320 / 24000)
        del mag, ph
        x48 = synth(256, 480000, 48000, 1245)
        from mlx_audio_plus_b200.dsp import ISTFTCache, hamming, stft
        wm = hamming(1920, periodic=False)
        rec("G mossformer stft 256 x 10 s (1920/384, complex)", "forward 1920/384 complex", lambda: stft(x48, 1920, 384, 1920, wm, center=False), x48.numel() * 4, 2560)
        spec = stft(x48, 1920, 384, 1920, wm, center=False)  # (B, T, F)
        re, im = spec.real.transpose(1, 2).contiguous(), spec.imag.transpose(1, 2).contiguous()
        del spec
        cache = ISTFTCache()
        rec("G mossformer ISTFTCache.istft 256 x (961, 1246) planes (1920/384)", "inverse 1920/384 (planar)",
            lambda: cache.istft(re, im, 1920, 384, 1920, wm, center=False, audio_length=480000), 2 * re.numel() * 4, 2560)
        del re, im
        x16 = synth(1, 16000 * 3600, 16000, 1246)[0] * 8000.0  # compute_fbank_kaldi takes ONE waveform (dsp.py:607-608)
        rec("G kaldi fbank 1 h (400/160 -> 512-point, 80 mel, povey)", "forward kaldi 512",
            lambda: compute_fbank_kaldi(x16, sample_rate=16000, win_len=400, win_inc=160, num_mels=80, win_type="povey", dither=0.0), x16.numel() * 4, 3600)
        del x16
        torch.cuda.empty_cache()
    for r in res:
        print(json.dumps(r), flush=True)
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
