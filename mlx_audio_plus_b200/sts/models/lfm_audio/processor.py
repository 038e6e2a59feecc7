"""Drop-in for AudioPreprocessor of mlx_audio/sts/models/lfm_audio/processor.py:34-140 (the NeMo-style front-end of
LFM2-Audio): optional dither, pre-emphasis, CONSTANT centre padding, power spectrum, Slaney/slaney filterbank,
ln(x + 2**-24), per-feature normalisation whose mean / Bessel-corrected std come from the first len // hop frames only
and are applied to ALL frames.  The whole batch runs in one pass of the fused kernel; the valid-frame statistic uses the
split form of the C ABI (b2a_frontend_partial over [0, n) with statistics, over [n, T) without, b2a_frontend_finalize with
n as the frame count)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ...._arrays import emit
from ...._wrap import as_batch
from .... import _lib as L
from ....dsp import _resolve_window, mel_filters
from ....frontend import FrontendPlan, _device_index, cached_plan

_LOG_ZERO_GUARD = 5.96e-8


@dataclass
class PreprocessorConfig:  # sts/models/lfm_audio/config.py:12-36
    sample_rate: int = 16000
    normalize: str = "per_feature"
    window_size: float = 0.025
    window_stride: float = 0.01
    window: str = "hann"
    features: int = 128
    n_fft: int = 512
    log: bool = True
    frame_splicing: int = 1
    dither: float = 1e-05
    pad_to: int = 0
    pad_value: float = 0.0
    preemph: float = 0.97

    @property
    def hop_length(self) -> int:
        return int(self.sample_rate * self.window_stride)

    @property
    def win_length(self) -> int:
        return int(self.sample_rate * self.window_size)


class AudioPreprocessor:
    def __init__(self, config: PreprocessorConfig = PreprocessorConfig()):
        self.config = config
        self._mel_filters = mel_filters(sample_rate=config.sample_rate, n_fft=config.n_fft, n_mels=config.features,
                                        f_min=0.0, f_max=config.sample_rate // 2, norm="slaney", mel_scale="slaney")

    @property
    def hop_length(self) -> int:
        return int(self.config.sample_rate * self.config.window_stride)

    @property
    def win_length(self) -> int:
        return int(self.config.sample_rate * self.config.window_size)

    def __call__(self, audio):
        import torch

        c = self.config
        ing, single = as_batch(audio)
        dev = ing.device if ing.on_device else torch.device("cuda", _device_index())  # raises without a GPU: no CPU fallback
        x = ing.data if ing.on_device else torch.from_numpy(ing.data).to(dev)
        if c.dither > 0:  # processor.py:79-83 (a fresh normal draw per call; not reproducible against the reference's RNG)
            x = x + c.dither * torch.randn_like(x)
        B, Lx = int(x.shape[0]), int(x.shape[1])
        per_feature = c.normalize == "per_feature"
        plan = cached_plan(
            FrontendPlan, dev.index if dev.index is not None else 0, _resolve_window(c.window, self.win_length, False),
            np.asarray(self._mel_filters, dtype=np.float32), n_fft=c.n_fft, hop=self.hop_length, center=True,
            pad_mode="constant", preemph=float(c.preemph) if c.preemph > 0 else 0.0, spec_kind=L.SPEC_POWER,
            log_kind=L.LOG_LN if c.log else L.LOG_NONE, guard_kind=L.GUARD_ADD if c.log else L.GUARD_NONE,
            guard_eps=_LOG_ZERO_GUARD if c.log else 0.0, norm_kind=L.NORM_PER_FEATURE if per_feature else L.NORM_NONE,
            norm_ddof=1, norm_eps=1e-5)
        T, M = plan.out_frames(Lx), plan.n_out
        n = min(Lx // self.hop_length, T)  # frames that enter the statistics (processor.py:121-122)
        with torch.cuda.device(dev):
            x = x.contiguous()
            out = torch.empty((B, T, M), dtype=torch.float32, device=dev)
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if not per_feature or n == T:
                a = plan._args(x.data_ptr(), Lx, Lx, Lx, B, out.data_ptr(), frame_count=T)
                L.check(L.lib.b2a_frontend_forward(plan._h, C.byref(a), st))
            else:
                if n < 2:
                    raise ValueError("Input is too short for a per-feature statistic over len // hop frames")
                cmax, sums = plan.stats_tensors(B, dev)
                cmax2, sums2 = plan.stats_tensors(B, dev)
                a = plan._args(x.data_ptr(), Lx, Lx, Lx, B, out.data_ptr(), frame_begin=0, frame_count=n,
                               clip_max=cmax.data_ptr(), feat_sums=sums.data_ptr())
                a.out_clip_stride = T * M
                L.check(L.lib.b2a_frontend_partial(plan._h, C.byref(a), st))
                b = plan._args(x.data_ptr(), Lx, Lx, Lx, B, out.data_ptr() + 4 * n * M, frame_begin=n, frame_count=T - n,
                               clip_max=cmax2.data_ptr(), feat_sums=sums2.data_ptr())
                b.out_clip_stride = T * M
                L.check(L.lib.b2a_frontend_partial(plan._h, C.byref(b), st))
                f = plan._args(x.data_ptr(), Lx, Lx, Lx, B, out.data_ptr(), frame_begin=0, frame_count=T,
                               clip_max=cmax.data_ptr(), feat_sums=sums.data_ptr())
                f.out_clip_stride = T * M
                L.check(L.lib.b2a_frontend_finalize(plan._h, C.byref(f), n, st))
        res = out[0] if single else out
        if not ing.on_device:
            res = res.cpu().numpy()
        return emit(ing, res)
