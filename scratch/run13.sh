python -m pytest tests -q -m gpu -x -k "istft or vocos or hift or polar" 2>&1 | tail -2
python benchmarks/bench_configs.py --only C5 --steps 10 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'istft' in d['config']: print(d['config'], '| ms', round(d['ms'], 4), '| frac', round(d['frac_of_hbm_peak'], 3))"
