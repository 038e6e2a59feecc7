for L in "" /root/repo/mlx_audio_plus_b200/lib/lib_k3.so; do
[ -n "$L" ] && export B2A_LIB=$L
python -m pytest tests -q -m gpu -x -k "istft or vocos or hift or polar" 2>&1 | tail -1
python benchmarks/bench_configs.py --only C5 --steps 10 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'istft' in d['config'] and 'B=1 ' not in d['config']: print('$L', d['config'], '| ms', round(d['ms'], 4), '| frac', round(d['frac_of_hbm_peak'], 3))"
done
