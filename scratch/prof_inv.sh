cat > /tmp/inv_bench.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
from mlx_audio_plus_b200._arrays import Ingested
from mlx_audio_plus_b200.dsp import hanning
from mlx_audio_plus_b200.frontend import IstftPlan
B=int(sys.argv[1]) if len(sys.argv)>1 else 1024
iplan = IstftPlan(n_fft=1024, hop=256, window=np.asarray(hanning(1024)), center=True)
g = torch.Generator(device="cuda"); g.manual_seed(7)
mag = torch.exp(0.5 * torch.randn((B, 513, 468), generator=g, device="cuda")).clamp(max=1e2)
ph = torch.randn((B, 513, 468), generator=g, device="cuda")
spec = torch.complex(mag * torch.cos(ph), mag * torch.sin(ph)).contiguous()
ing = Ingested("torch", True, spec, None, spec.device)
for _ in range(3): out = iplan.run(ing)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): out = iplan.run(ing)
e1.record(); torch.cuda.synchronize()
print("ms", e0.elapsed_time(e1)/10)
PY
python /tmp/inv_bench.py 1024
ncu --set full --clock-control none --import-source on -k regex:fast_istft -s 2 -c 1 -f -o gpurun_out/prof_inv python /tmp/inv_bench.py 256 > gpurun_out/ncu_inv.log 2>&1
tail -1 gpurun_out/ncu_inv.log
