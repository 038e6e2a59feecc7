timeout 120 python -m pytest tests -q -m gpu -x -k "whisper or logmel or framing" 2>&1 | tail -2
timeout 120 python bench.py --clips 4096 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('WS ms', d['ms_per_step'], 'frac', d['roofline']['frac'])"
B2A_NO_WS=1 timeout 120 python bench.py --clips 4096 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('sym ms', d['ms_per_step'], 'frac', d['roofline']['frac'])"
