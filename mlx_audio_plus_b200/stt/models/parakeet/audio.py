"""Drop-in for mlx_audio/stt/models/parakeet/audio.py (PreprocessArgs, log_mel_spectrogram): NeMo-style
features — pre-emphasis, right-padded window, power, Slaney-scale filterbank WITHOUT area normalisation
(norm=args.normalize is forwarded to mel_filters, audio.py:59-61), ln(x+1e-5), per-feature (ddof=0) or
global normalisation.  Statistics are accumulated in float64 on the device."""
from __future__ import annotations

from dataclasses import dataclass

from ...._arrays import emit
from ...._wrap import as_batch, run_frontend
from .... import _lib as L
from ....dsp import STR_TO_WINDOW_FN, hanning, mel_filters


@dataclass
class PreprocessArgs:  # reference audio.py:16-36
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
    def win_length(self) -> int:
        return int(self.window_size * self.sample_rate)

    @property
    def hop_length(self) -> int:
        return int(self.window_stride * self.sample_rate)


def log_mel_spectrogram(x, args: PreprocessArgs):
    """(L,) -> (1, T, features) (reference audio.py:39-78).  `dither` is ignored, as in the reference."""
    ing, was_1d = as_batch(x)
    L_in = ing.data.shape[1]
    length = max(L_in, args.pad_to) if args.pad_to > 0 else L_in  # audio.py:42-45
    window_fn = STR_TO_WINDOW_FN.get(args.window, None)  # no .lower() here (audio.py:47)
    window = window_fn(args.win_length) if window_fn else hanning(args.win_length)
    preemph = getattr(args, "preemph", 0.97)
    fb = mel_filters(args.sample_rate, args.n_fft, args.features, norm=args.normalize, mel_scale=None)
    out = run_frontend(
        ing, window, fb, length=length, pad_value=float(args.pad_value),
        n_fft=args.n_fft, hop=args.hop_length, center=True, pad_mode="reflect",
        preemph=float(preemph) if preemph > 0 else 0.0, spec_kind=L.SPEC_POWER,
        log_kind=L.LOG_LN, guard_kind=L.GUARD_ADD, guard_eps=1e-5,
        norm_kind=L.NORM_PER_FEATURE if args.normalize == "per_feature" else L.NORM_GLOBAL,
        norm_ddof=0, norm_eps=1e-5)
    res = emit(ing, out)  # (1, T, M) for a 1-D input, (B, T, M) for a batch
    # audio.py:78 returns in the INPUT dtype (parakeet.py:184,227 may hand bfloat16).  The arithmetic here stays float32
    # throughout (the reference rounds the power spectrum to the input dtype before the mel product, audio.py:58); only
    # the result is cast.
    od = ing.orig_dtype
    if od is not None and ing.family == "torch" and od != res.dtype and od.is_floating_point:
        res = res.to(od)
    return res
