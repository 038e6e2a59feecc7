python -m pytest tests -q -m gpu 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('packed', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
python benchmarks/bench_configs.py --only C3,C4,C5 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print(d['config'][:44].ljust(44), d['kernel'].ljust(22), '%9.3f ms %9.1f ah/s %7.1f GB/s  %.3f' % (d['ms'], d['audio_hours_per_s'], d['algorithmic_GBps'], d['frac_of_hbm_peak']))
    except Exception as e: print(l.rstrip()[:200])"
