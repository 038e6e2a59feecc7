import sys, numpy as np, torch, scipy.fft
sys.path.insert(0, '.')
from mlx_audio_plus_b200.dsp import stft
from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram
from oracle import dsp_oracle as D, wrappers_oracle as W
from bench import synth_clip_np
x = synth_clip_np(0)
ref = W.whisper_log_mel(x, 128)
xd = torch.from_numpy(x).cuda()
y = log_mel_spectrogram(xd, 128).cpu().numpy()
fb = D.mel_filters(16000, 400, 128, norm="slaney", mel_scale=None)
def tail(S):
    mel = (np.abs(S[:-1]) ** 2).astype(np.float32) @ fb.T
    t = np.log10(np.maximum(mel, 1e-10)); t = np.maximum(t, t.max() - 8); return ((t + 4) / 4).astype(np.float32), mel
g = stft(xd, 400, 160, window=D.hanning(400)).cpu().numpy()
yg, melg = tail(g)
fr = D.frames_of(x, 400, 160) * D.hanning(400)
ys, mels = tail(scipy.fft.rfft(fr.astype(np.float32)))
for nm, a in (("kernel", y), ("gpu-stft + numpy tail", yg), ("scipy-stft + numpy tail", ys)):
    e = np.abs(a - ref)
    print("%-24s max %.3e mean %.3e  n>2e-5: %d  n>5e-5: %d" % (nm, e.max(), e.mean(), (e > 2e-5).sum(), (e > 5e-5).sum()))
e = np.abs(y - ref); i = np.unravel_index(e.argmax(), e.shape)
print("worst", i, "kernel", y[i], "gpu-stft-tail", yg[i], "scipy-tail", ys[i], "ref", ref[i], "mel(g)", melg[i], "mel(s)", mels[i])
m = i[1]; nz = np.nonzero(fb[m])[0]; print("mel taps", nz, fb[m][nz])
print("gpu |X|^2 at taps", np.abs(g[i[0], nz]) ** 2, "scipy", np.abs(scipy.fft.rfft(fr[i[0]].astype(np.float32))[nz]) ** 2)
tr = np.fft.rfft(fr[i[0]].astype(np.float64)); print("truth", np.abs(tr[nz]) ** 2)
