import sys, numpy as np, torch, scipy.fft
sys.path.insert(0, '.')
from mlx_audio_plus_b200.dsp import stft
from oracle import dsp_oracle as D
from bench import synth_clip_np
x = synth_clip_np(0)[:160*2000]
w = D.hanning(400)
fr = D.frames_of(x, 400, 160) * w
tr = np.fft.rfft(fr.astype(np.float64))
sp = scipy.fft.rfft(fr.astype(np.float32))
g = stft(torch.from_numpy(x).cuda(), 400, 160, window=w).cpu().numpy()
wk = np.abs(tr) < 0.2
for nm, a in (("scipy", sp), ("gpu", g)):
    e = np.abs(a - tr)
    print(nm, "all: max %.3e mean %.3e | weak: max %.3e mean %.3e | even-frames mean %.3e odd-frames mean %.3e" % (
        e.max(), e.mean(), e[wk].max(), e[wk].mean(), e[0::2].mean(), e[1::2].mean()))
e = np.abs(g - tr)
print("per-bin mean err (gpu) top10 bins:", np.argsort(e.mean(0))[-10:], np.sort(e.mean(0))[-10:])
es = np.abs(sp - tr)
print("per-bin mean err (scipy) top10 bins:", np.argsort(es.mean(0))[-10:], np.sort(es.mean(0))[-10:])
