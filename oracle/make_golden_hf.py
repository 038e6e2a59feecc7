#!/usr/bin/env python
"""ORACLE / TEST INFRASTRUCTURE — generates tests/golden/hf_whisper_fe.npz by running the REAL third-party dependency the
reference calls for Qwen3-ASR / Qwen3-ForcedAligner features (mlx_audio/stt/models/qwen3_asr/qwen3_asr.py:800-846):
`transformers.WhisperFeatureExtractor` (the version installed in the build image; printed into the fixture), with the
reference's own call arguments.  Run in the build container:  python oracle/make_golden_hf.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.make_golden import synth  # noqa: E402


def main():
    import transformers
    from transformers import WhisperFeatureExtractor

    g = {"transformers_version": np.array(transformers.__version__)}
    fe = WhisperFeatureExtractor(feature_size=128)  # Qwen3-ASR: num_mel_bins 128
    g["fb128"] = fe.mel_filters
    g["fb80"] = WhisperFeatureExtractor(feature_size=80).mel_filters
    kw = dict(sampling_rate=16000, return_attention_mask=True, truncation=False, padding=True, return_tensors="np")
    x = synth(600, 16000 * 2 + 77)
    g["one|x"] = x
    o = fe(x, **kw)  # qwen3_asr.py:835-842
    g["one|features"], g["one|mask"] = o["input_features"], o["attention_mask"]
    a, b, c = synth(601, 24000), 0.3 * synth(602, 16000), synth(603, 9999)
    g["batch|a"], g["batch|b"], g["batch|c"] = a, b, c
    o = fe([a, b, c], **kw)
    g["batch|features"], g["batch|mask"] = o["input_features"], o["attention_mask"]
    o = fe(x, sampling_rate=16000, return_tensors="np", max_length=48000)  # the class defaults: max_length padding, truncation
    g["default|features"] = o["input_features"]
    o = fe(np.concatenate([x, x]), sampling_rate=16000, return_tensors="np", max_length=48000, return_attention_mask=True,
           do_normalize=True)
    g["trunc_norm|features"], g["trunc_norm|mask"] = o["input_features"], o["attention_mask"]
    np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "hf_whisper_fe.npz"), **g)
    print({k: getattr(v, "shape", v) for k, v in g.items()})


if __name__ == "__main__":
    main()
