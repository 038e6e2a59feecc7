python -m pytest tests -q -m gpu 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('float2 tw', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
