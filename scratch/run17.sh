python -m pytest tests -q -m gpu 2>&1 | tail -2
python benchmarks/bench_configs.py --only W --steps 10 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'], '| ms', round(d['ms'], 4), '| frac', round(d['frac_of_hbm_peak'], 3))"
