import time, torch, numpy as np, os, sys
sys.path.insert(0, os.getcwd())
from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram
from mlx_audio_plus_b200 import _lib as L
x = torch.randn(480000, device="cuda") * 0.1
for m in (80, 128):
    for _ in range(30): log_mel_spectrogram(x, n_mels=m)
    torch.cuda.synchronize()
    n0 = L.lib.b2a_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(300): log_mel_spectrogram(x, n_mels=m)
    e1.record(); torch.cuda.synchronize()
    b2b = e0.elapsed_time(e1) / 300 * 1e3
    nl = (L.lib.b2a_launch_count() - n0) / 300
    ts = []
    for _ in range(300):
        t0 = time.perf_counter(); log_mel_spectrogram(x, n_mels=m); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print("mels %d: back-to-back %.1f us/call, launches/call %.1f, synchronised latency %.1f us" % (m, b2b, nl, np.median(ts) * 1e6))
