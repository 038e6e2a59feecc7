#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
C4="python bench.py --workload kokoro_istft --clips 1024 --steps 3 --no-cpu-baseline --no-e2e"
$C4 > gpurun_out/p3_c4_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:istft_small -s 3 -c 1 -f -o gpurun_out/r02_k4_polar $C4 > gpurun_out/p3_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
C3="python bench.py --workload parakeet_64x1h --clips 4 --steps 3 --no-cpu-baseline --no-e2e"
$C3 > gpurun_out/p3_c3_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:normalise -s 3 -c 1 -f -o gpurun_out/r02_normalise $C3 > gpurun_out/p3_ncu_norm.log 2>&1; echo "ncu norm rc=$?"
