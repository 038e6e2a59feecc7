#!/bin/bash
# fused single-launch forward: full suite, then headline A/B (fused vs three launches) at 4096 and 512 clips, C1 latency
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/f_pytest.log
for env in "" "B2A_NO_FUSED_FORWARD=1"; do
  for clips in 4096 512; do
    env $env python bench.py --clips $clips --no-extras --no-e2e --no-cpu-baseline 2>gpurun_out/f.err | python -c "import json,sys;d=json.loads(sys.stdin.read());print('[$env] clips $clips ms %.4f kernel %.4f launches %d'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['gpu_launches']))" || tail -3 gpurun_out/f.err
  done
done
python - <<'P'
import time, torch, numpy as np, os, sys
sys.path.insert(0, os.getcwd())
from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram
from mlx_audio_plus_b200 import _lib as L
x = torch.randn(480000, device="cuda") * 0.1
for _ in range(30): log_mel_spectrogram(x, n_mels=80)
torch.cuda.synchronize()
n0 = L.lib.b2a_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(300): log_mel_spectrogram(x, n_mels=80)
e1.record(); torch.cuda.synchronize()
print("C1 back-to-back us/call %.1f, launches/call %.1f" % (e0.elapsed_time(e1) / 300 * 1e3, (L.lib.b2a_launch_count() - n0) / 300))
ts = []
for _ in range(200):
    t0 = time.perf_counter(); log_mel_spectrogram(x, n_mels=80); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("C1 synchronised latency us %.1f" % (np.median(ts) * 1e6))
P
