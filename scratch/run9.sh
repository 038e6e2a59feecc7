for e in "" "B2A_NO_MELSPEC=1"; do
env $e python benchmarks/bench_configs.py --only C5,C3 --steps 6 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'B=8192' in d['config'] or 'B=1024 x 5' in d['config'] or '16 x' in d['config']: print('$e', d['config'], '|', d['kernel'], '| ms', round(d['ms'], 4), '| frac', round(d['frac_of_hbm_peak'], 3))"
done
