python -m pytest tests -q -m gpu 2>&1 | tail -2
for sk in 0 4; do
B2A_MEL_BY_MEL=1 B2A_NO_BULK_STORE=1 B2A_SKIP=$sk python bench.py --clips 2048 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('lane=mel skip=$sk kernel_ms', round(d['roofline']['kernel_ms'],3))"
done
B2A_SKIP=0 python bench.py --clips 2048 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('melf kernel_ms', round(d['roofline']['kernel_ms'],3))"
