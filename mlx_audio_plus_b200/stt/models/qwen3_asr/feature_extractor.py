"""Drop-in for the feature step of Qwen3-ASR / Qwen3-ForcedAligner (and any caller of Hugging Face's
`WhisperFeatureExtractor`): mlx_audio/stt/models/qwen3_asr/qwen3_asr.py:800-846 and qwen3_forced_aligner.py:589-630 call

    self._feature_extractor(audio_np, sampling_rate=16000, return_attention_mask=True, truncation=False, padding=True,
                            return_tensors="np")  ->  {"input_features": (B, n_mels, T), "attention_mask": (B, T)}

with transformers' extractor (a third-party dependency of the reference: `transformers`, any 4.x / 5.x; semantics restated
from transformers/models/whisper/feature_extraction_whisper.py and audio_utils.mel_filter_bank).  Same call signature and
result keys; the padded batch goes through ONE launch of the fused 400/160 kernel: periodic Hann-400, reflect-centred frames,
power spectrum, Slaney/slaney filterbank up to 8 kHz computed in float64 as transformers does, log10(max(., 1e-10)), last frame
dropped, per-clip max - 8 clamp, (x + 4) / 4, (B, n_mels, T) layout."""
from __future__ import annotations

import numpy as np

from ...._arrays import _is_torch
from ...._wrap import as_batch, run_frontend
from .... import _lib as L
from ....dsp import hanning


def mel_filter_bank_slaney(num_frequency_bins: int, num_mel_filters: int, min_frequency: float, max_frequency: float,
                           sampling_rate: int) -> np.ndarray:
    """transformers.audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney") -> (num_frequency_bins, num_mel_filters)
    float64: Slaney mel scale (linear below 1 kHz, logarithmic above), triangles from the frequency differences, area
    normalisation 2 / (f[i+2] - f[i])."""
    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * (27.0 / np.log(6.4)), 3.0 * f / 200.0)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), 200.0 * m / 3.0)

    mel_freqs = np.linspace(hz_to_mel(min_frequency), hz_to_mel(max_frequency), num_mel_filters + 2)
    filter_freqs = mel_to_hz(mel_freqs)
    fft_freqs = np.linspace(0, sampling_rate // 2, num_frequency_bins)
    filter_diff = np.diff(filter_freqs)
    slopes = np.expand_dims(filter_freqs, 0) - np.expand_dims(fft_freqs, 1)
    down = -slopes[:, :-2] / filter_diff[:-1]
    up = slopes[:, 2:] / filter_diff[1:]
    fb = np.maximum(np.zeros(1), np.minimum(down, up))
    enorm = 2.0 / (filter_freqs[2 : num_mel_filters + 2] - filter_freqs[:num_mel_filters])
    return fb * np.expand_dims(enorm, 0)


class WhisperFeatureExtractor:
    model_input_names = ["input_features"]

    def __init__(self, feature_size=80, sampling_rate=16000, hop_length=160, chunk_length=30, n_fft=400, padding_value=0.0,
                 dither=0.0, return_attention_mask=False, **kwargs):
        self.feature_size, self.sampling_rate, self.hop_length = feature_size, sampling_rate, hop_length
        self.chunk_length, self.n_fft, self.padding_value, self.dither = chunk_length, n_fft, padding_value, dither
        self.return_attention_mask = return_attention_mask
        self.n_samples = chunk_length * sampling_rate
        self.nb_max_frames = self.n_samples // hop_length
        self.mel_filters = mel_filter_bank_slaney(1 + n_fft // 2, feature_size, 0.0, 8000.0, sampling_rate)
        # (n_mels, F) float32, made once and read-only: the plan cache fingerprints a read-only table once, not per call
        self._fb = np.ascontiguousarray(self.mel_filters.T, dtype=np.float32)
        self._fb.setflags(write=False)

    # -- transformers' SequenceFeatureExtractor.pad for one float feature per time step ------------------------------------
    def _pad(self, clips, padding, max_length, truncation, pad_to_multiple_of):
        if padding is True or padding == "longest":
            strategy = "longest"
        elif padding == "max_length":
            strategy = "max_length"
        elif padding is False or padding is None or padding == "do_not_pad":
            strategy = "do_not_pad"
        else:
            raise ValueError(f"unknown padding strategy {padding!r}")
        if truncation:
            if max_length is None:
                raise ValueError("When setting ``truncation=True``, make sure that ``max_length`` is defined.")
            lim = max_length
            if pad_to_multiple_of and lim % pad_to_multiple_of:
                lim = (lim // pad_to_multiple_of + 1) * pad_to_multiple_of
            clips = [c[:lim] for c in clips]
        lengths = [int(c.shape[0]) for c in clips]
        if strategy == "longest":
            target = max(lengths)
        elif strategy == "max_length":
            target = max_length
        else:
            if len(set(lengths)) > 1:
                raise ValueError("do_not_pad with clips of different lengths cannot be returned as one array")
            target = lengths[0]
        if strategy != "do_not_pad" and pad_to_multiple_of and target % pad_to_multiple_of:
            target = (target // pad_to_multiple_of + 1) * pad_to_multiple_of
        if any(n > target for n in lengths):
            raise ValueError("a clip is longer than the padding target; pass truncation=True")
        return target, lengths

    def __call__(self, raw_speech, truncation=True, pad_to_multiple_of=None, return_tensors=None, return_attention_mask=None,
                 padding="max_length", max_length=None, sampling_rate=None, do_normalize=None, device=None, **kwargs):
        if sampling_rate is not None and sampling_rate != self.sampling_rate:
            raise ValueError(f"The model corresponding to this feature extractor: {self.__class__.__name__} was trained using a"
                             f" sampling rate of {self.sampling_rate}. Please make sure that the provided `raw_speech` input"
                             f" was sampled with {self.sampling_rate} and not {sampling_rate}.")
        import torch

        on_device = _is_torch(raw_speech) and raw_speech.is_cuda
        max_length = max_length if max_length else self.n_samples
        two_d = False
        if _is_torch(raw_speech) or isinstance(raw_speech, np.ndarray):
            if raw_speech.ndim > 2:
                raise ValueError(f"Only mono-channel audio is supported for input to {self}")
            two_d = raw_speech.ndim == 2
            if two_d:  # a rectangular batch: one representative row stands for all of them in the length logic
                if truncation:
                    lim = max_length
                    if pad_to_multiple_of and lim % pad_to_multiple_of:
                        lim = (lim // pad_to_multiple_of + 1) * pad_to_multiple_of
                    raw_speech = raw_speech[:, :lim]
                clips = [raw_speech[0]] * int(raw_speech.shape[0])
            else:
                clips = [raw_speech]
        elif isinstance(raw_speech, (list, tuple)) and len(raw_speech) and (
                isinstance(raw_speech[0], (np.ndarray, list, tuple)) or _is_torch(raw_speech[0])):
            clips = [c if hasattr(c, "shape") else np.asarray(c, np.float32) for c in raw_speech]
        else:
            clips = [np.asarray(raw_speech, dtype=np.float32)]
        target, lengths = self._pad(clips, padding, max_length, truncation and not two_d, pad_to_multiple_of)
        if on_device:
            dev = raw_speech.device
        elif torch.cuda.is_available():
            dev = torch.device("cuda", torch.cuda.current_device())
        else:
            raise L.B2AError("b200audio: no CUDA device — this library has no CPU fallback")
        len_t = torch.tensor(lengths, dtype=torch.int64, device=dev)
        if two_d:  # a rectangular batch: no per-clip copies
            t = raw_speech if _is_torch(raw_speech) else torch.from_numpy(np.ascontiguousarray(raw_speech, dtype=np.float32))
            t = t.to(device=dev, dtype=torch.float32)
            if lengths[0] == target:
                batch = t.contiguous()
            else:
                batch = torch.full((len(clips), target), float(self.padding_value), dtype=torch.float32, device=dev)
                batch[:, : lengths[0]] = t
        else:
            batch = torch.full((len(clips), target), float(self.padding_value), dtype=torch.float32, device=dev)
            for i, (c, n) in enumerate(zip(clips, lengths)):
                t = c if _is_torch(c) else torch.from_numpy(np.ascontiguousarray(np.asarray(c, dtype=np.float32)))
                batch[i, :n] = t[:n].to(device=dev, dtype=torch.float32)
        if do_normalize:  # zero_mean_unit_var_norm over the valid samples, padding back to padding_value: one launch
            from ...._post import rows_normalize

            batch = rows_normalize(batch, len_t, den_kind=0, eps=1e-7, pad_value=float(self.padding_value))
        if self.dither != 0.0:
            batch = batch + self.dither * torch.randn_like(batch)
        ing, _ = as_batch(batch)
        feats = run_frontend(
            ing, hanning(self.n_fft, True), self._fb, n_fft=self.n_fft,
            hop=self.hop_length, center=True, pad_mode="reflect", drop_last=True, spec_kind=L.SPEC_POWER,
            log_kind=L.LOG_LOG10, guard_kind=L.GUARD_MAX, guard_eps=1e-10, clamp_kind=L.CLAMP_CLIP_MAX, clamp_value=8.0,
            affine_add=4.0, affine_div=4.0, out_layout=L.LAYOUT_MT)  # (B, n_mels, T)
        out = {"input_features": feats}
        if return_attention_mask if return_attention_mask is not None else self.return_attention_mask:
            # the sample-level mask sampled every hop samples (attention_mask[:, ::hop]), built at frame resolution directly
            starts = torch.arange(0, target, self.hop_length, device=dev)
            m = (starts[None, :] < len_t[:, None]).to(torch.int32)
            if target % self.hop_length != 0:  # L // hop + 1 frames minus the dropped one
                m = m[:, :-1]
            out["attention_mask"] = m.contiguous()
        if return_tensors in ("cuda", "device"):  # extension: stay on the GPU
            return out
        if return_tensors == "pt":
            return {k: v.cpu() for k, v in out.items()}
        return {k: v.cpu().numpy() for k, v in out.items()}
