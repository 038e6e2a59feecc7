python -m pytest tests -q -m gpu -x -k "whisper or logmel or framing" 2>&1 | tail -1
python bench.py --clips 4096 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms', d['ms_per_step'], 'frac', d['roofline']['frac'])"
