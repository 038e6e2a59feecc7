python -m pytest tests -q -m gpu 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
python bench.py --impl reference --steps 2 --warmup 3 2>/dev/null | cut -c1-200
