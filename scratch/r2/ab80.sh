#!/bin/bash
cd "$GRAFT_REPO_ROOT"
lib="$1"
for ws in 0 1; do
  for wl in whisper80_30s whisper128_30s; do
    env B2A_LIB="$PWD/$lib" B2A_WS=$ws timeout 90 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab80.json 2> gpurun_out/ab80.err
    python -c "import json;d=json.load(open('gpurun_out/ab80.json'));print('ws=$ws $wl ms/step %.3f kernel_ms %.3f frac %.3f'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
  done
done
env B2A_LIB="$PWD/$lib" B2A_TMA_OUT=0 timeout 90 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab80.json 2> gpurun_out/ab80.err
python -c "import json;d=json.load(open('gpurun_out/ab80.json'));print('ws=1 tma_out=0 whisper128 ms/step %.3f kernel_ms %.3f frac %.3f'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
