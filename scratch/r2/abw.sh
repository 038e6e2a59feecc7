#!/bin/bash
# abw.sh <pytest -k filter or "-"> lib:workload:clips[:ENV=VAL,...] ... : one bench.py workload per spec, kernel-only, on a development library
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
filt="$1"; shift
if [ "$filt" != "-" ]; then timeout 600 python -m pytest tests -m gpu -x -q -k "$filt" 2>&1 | tail -4; fi
for spec in "$@"; do
  IFS=: read lib w c envs <<< "$spec"; envs="${envs//,/ }"
  tag=${lib}_${w}
  env B2A_LIB="$PWD/mlx_audio_plus_b200/lib/$lib.so" $envs timeout 300 python bench.py --workload $w --clips $c --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/abw_$tag.json 2> gpurun_out/abw_$tag.err || tail -3 gpurun_out/abw_$tag.err
  python -c "
import json;d=json.load(open('gpurun_out/abw_$tag.json'));r=d['roofline'];print('$lib $w $envs', 'ms %.4f kernel_ms %.4f binding %.3f hbm %.3f kernel %s'%(d['ms_per_step'], r['kernel_ms'], r['binding_frac'], r['frac'], d['config']['kernel']))"
done
