python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python benchmarks/bench_configs.py --only S --steps 6 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'], '|', d['kernel'], '| ms', round(d['ms'], 4), '| frac', round(d['frac_of_hbm_peak'], 3), '| torch', round(d['torch_stft_ms'],3))"
B2A_NO_FAST_STFT=1 python benchmarks/bench_configs.py --only S --steps 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if '800' in d['config']: print('generic:', d['config'], '|', d['kernel'], '| ms', round(d['ms'], 4), '| frac', round(d['frac_of_hbm_peak'], 3))"
