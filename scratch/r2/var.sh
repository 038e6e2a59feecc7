#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python bench.py --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print('sampler   ms %.3f kernel %.3f'%(d['ms_per_step'], d['roofline']['kernel_ms']))"; done
for i in 1 2 3 4 5 6; do B2A_BENCH_NO_SAMPLER=1 python bench.py --no-extras --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print('nosampler ms %.3f kernel %.3f'%(d['ms_per_step'], d['roofline']['kernel_ms']))"; done
