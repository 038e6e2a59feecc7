"""stft_small_kernel timing: Kokoro's forward transform (n_fft 20, hop 5) over 1024 x 5 s at 24 kHz."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from mlx_audio_plus_b200.dsp import stft, hanning
x = torch.randn((1024, 120000), device="cuda")
w = np.asarray(hanning(21)[:-1])
y = stft(x, 20, 5, 20, w)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): stft(x, 20, 5, 20, w)
e0.record()
for _ in range(10): y = stft(x, 20, 5, 20, w)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
by = x.numel() * 4 + y.numel() * 8
print("stft_small 20/5 1024 x 5 s: %.3f ms, %.0f GB/s algorithmic = %.3f of 6545" % (ms, by / ms / 1e6, by / ms / 1e6 / 6545), tuple(y.shape))
