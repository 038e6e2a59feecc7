#!/bin/bash
# resampler: parity tests, then bench_configs section R with optional env toggles
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "resample or load_audio" 2>&1 | tail -3
for envs in "$@"; do
echo "== $envs"
env ${envs//,/ } python benchmarks/bench_configs.py --only R --steps 10 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.rstrip()[-300:]); continue
    print(d['config'][:70], '| ms %.3f | ah/s %.0f | frac %.3f' % (d['ms'], d['audio_hours_per_s'], d['frac_of_hbm_peak']))
"
done
