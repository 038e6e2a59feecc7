import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
from mlx_audio_plus_b200 import _lib as L
from mlx_audio_plus_b200.dsp import hanning, mel_filters
from mlx_audio_plus_b200.frontend import FrontendPlan
from oracle import wrappers_oracle as W
from bench import synth_clip_np
B, CL = int(sys.argv[1]), 480000
x = np.stack([synth_clip_np(i) for i in range(8)])
xb = np.tile(x, (B // 8, 1))
plan = FrontendPlan(n_fft=400, hop=160, window=np.asarray(hanning(400)), drop_last=True, spec_kind=L.SPEC_POWER,
    filterbank=np.asarray(mel_filters(16000, 400, 128, norm="slaney", mel_scale=None)), log_kind=L.LOG_LOG10,
    guard_kind=L.GUARD_MAX, guard_eps=1e-10, clamp_kind=L.CLAMP_CLIP_MAX, clamp_value=8.0, affine_add=4.0, affine_div=4.0)
ref = np.stack([W.whisper_log_mel(x[i], 128) for i in range(8)])
xd = torch.from_numpy(xb).cuda()
from mlx_audio_plus_b200._arrays import Ingested
yd = plan.run(Ingested("torch", True, xd, None, xd.device)).cpu().numpy()
yh = plan.run(Ingested("numpy", False, xb))
for name, y in (("device", yd), ("host", yh)):
    errs = [np.abs(y[i] - ref[i % 8]).max() for i in range(B)]
    bad = [i for i, e in enumerate(errs) if e > 1e-4]
    print(name, "max err", max(errs), "bad clips", len(bad), bad[:20])
