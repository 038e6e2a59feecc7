import sys, numpy as np, torch
sys.path.insert(0, '.')
from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram
from oracle import wrappers_oracle as W, dsp_oracle as D
from bench import synth_clip_np
x = synth_clip_np(0)
ref = W.whisper_log_mel(x, 128)
y = log_mel_spectrogram(torch.from_numpy(x).cuda(), 128).cpu().numpy()
# float64 truth
xp = np.pad(x.astype(np.float64), 200, mode="reflect")
idx = np.arange(400)[None, :] + 160 * np.arange(3001)[:, None]
w = D.hanning(400).astype(np.float64)
S = np.fft.rfft(xp[idx] * w)[:-1]
mel = (np.abs(S) ** 2) @ D.mel_filters(16000, 400, 128, norm="slaney", mel_scale=None).astype(np.float64).T
t = np.log10(np.maximum(mel, 1e-10)); t = np.maximum(t, t.max() - 8); t = (t + 4) / 4
e = np.abs(y - ref); i = np.unravel_index(e.argmax(), e.shape)
print("max |gpu-oracle|", e.max(), "at", i, "gpu", y[i], "oracle", ref[i], "truth", t[i])
print("max |gpu-truth|", np.abs(y - t).max(), "max |oracle-truth|", np.abs(ref - t).max())
print("mean |gpu-truth|", np.abs(y - t).mean(), "mean |oracle-truth|", np.abs(ref - t).mean())
j = np.unravel_index(np.abs(ref - t).argmax(), e.shape); print("oracle worst at", j, ref[j], t[j], y[j])
j = np.unravel_index(np.abs(y - t).argmax(), e.shape); print("gpu worst at", j, ref[j], t[j], y[j])
print("mel power at worst:", mel[i], "max mel", mel.max(), "min mel", mel.min())
