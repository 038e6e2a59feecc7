#!/bin/bash
# wl.sh <workload> <clips> [ENV=VAL ...] : one workload of bench.py, kernel-only, with optional env toggles
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
w=$1; c=$2; shift 2
env "$@" timeout 300 python bench.py --workload $w --clips $c --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/wl.json 2> gpurun_out/wl.err || tail -3 gpurun_out/wl.err
python -c "
import json;d=json.load(open('gpurun_out/wl.json'));r=d['roofline'];print('$w $*', 'ms %.3f kernel_ms %.3f binding %.3f hbm %.3f kernel %s'%(d['ms_per_step'], r['kernel_ms'], r['binding_frac'], r['frac'], d['config']['kernel']))"
