for n in orig u0; do
  if [ "$n" = orig ]; then L=/root/repo/mlx_audio_plus_b200/lib/libb200audio.so; else L=/root/repo/mlx_audio_plus_b200/lib/lib_$n.so; fi
  echo "== $n: $(B2A_LIB=$L python -m pytest tests -q -m gpu -x -k 'whisper or stft or parakeet or vocos or logmel' 2>&1 | tail -1)"
  for i in 1 2; do
    B2A_LIB=$L python bench.py --clips 4096 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$n C2', d['ms_per_step'], d['roofline']['kernel_ms'])"
  done
  B2A_LIB=$L python benchmarks/bench_configs.py --only C3,C5,S 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'B=1 ' in d['config'] or 'B=64 ' in d['config'] or 'one 1-hour' in d['config'] or 'istft' in d['config']: continue
    print('$n', d['config'], '| ms', round(d['ms'], 4))"
done
