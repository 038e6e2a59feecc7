#!/bin/bash
# A/B of development libraries: ab.sh <pytest -k filter or "-"> lib1[:ENV=VAL,...] lib2 ...   (paths relative to the repo root)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
filt="$1"; shift
for spec in "$@"; do
  lib="${spec%%:*}"; envs=""
  if [[ "$spec" == *:* ]]; then envs="${spec#*:}"; envs="${envs//,/ }"; fi
  tag=$(basename "$lib" .so)_$(echo "$envs" | tr -c 'A-Za-z0-9=\n' '_')
  echo "=== $lib [$envs]"
  if [ "$filt" != "-" ]; then
    env B2A_LIB="$PWD/$lib" $envs timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$filt" 2>&1 | tail -3
  fi
  env B2A_LIB="$PWD/$lib" $envs timeout 90 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  echo "rc=$?"; tail -3 gpurun_out/ab_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/ab_$tag.json'));print('ms/step %.3f kernel_ms %.3f frac %.3f'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
done
