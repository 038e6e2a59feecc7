#!/bin/bash
# generic-kernel work: parity of everything (full -m gpu suite), then section G of the per-config bench, with optional env toggles
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/g_pytest.log
for envs in "$@"; do
echo "== $envs"
env ${envs//,/ } python benchmarks/bench_configs.py --only G --steps 5 --out gpurun_out/g_new.json 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.rstrip()[-300:]); continue
    print(d['config'][:70], '| ms %.3f | ah/s %.0f | frac %.3f' % (d['ms'], d['audio_hours_per_s'], d['frac_of_hbm_peak']))
"
done
