python -m pytest tests -q -m gpu -x -k "stft or whisper or parakeet or framing or frontends" 2>&1 | tail -2
python benchmarks/bench_configs.py --only S --steps 6 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'], '|', d['kernel'], '| ms', round(d['ms'], 4), '| frac', round(d['frac_of_hbm_peak'], 3))"
python bench.py --clips 4096 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2', d['ms_per_step'], d['roofline']['kernel_ms'])"
