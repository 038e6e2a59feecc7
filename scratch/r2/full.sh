#!/bin/bash
# full GPU suite + smoke + headline bench (kernel only)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/full_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/full_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/full_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/full_smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/full_bench.json'));print('ms/step %.3f kernel_ms %.3f frac %.3f e2e %.1f'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value']))"
