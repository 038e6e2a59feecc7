python -m pytest tests -q -m gpu -x -k "funasr or mel_segment" 2>&1 | tail -8
python benchmarks/bench_configs.py --only P --steps 5 2>&1 | grep -v Warning | cut -c1-420
