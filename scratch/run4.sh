mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -2
for i in 1 2; do python bench.py --clips 4096 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; done
python benchmarks/bench_configs.py --out gpurun_out/configs_new.json 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'], '|', d['kernel'], '| ms', round(d['ms'], 4), '| ah/s', round(d['audio_hours_per_s'], 1), '| frac', round(d['frac_of_hbm_peak'], 3), d.get('torch_stft_ms', ''))"
