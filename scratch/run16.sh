python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python - <<'P'
import torch, sys
sys.path.insert(0, '.')
from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram
import os
x = torch.randn(1024, 480000, device='cuda') * 0.1
def t(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n
a = t(lambda: log_mel_spectrogram(x, n_mels=128, padding=480000))
os.environ['B2A_NO_PAD_SKIP']='1'
b = t(lambda: log_mel_spectrogram(x, n_mels=128, padding=480000))
print('1024 x 30 s, padding=N_SAMPLES: skip', a, 'ms; full', b, 'ms')
P
python bench.py --clips 4096 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2', d['ms_per_step'], d['roofline']['kernel_ms'])"
