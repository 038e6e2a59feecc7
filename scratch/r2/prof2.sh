#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
C5="python bench.py --workload vocos_mel --clips 2048 --steps 3 --no-cpu-baseline --no-e2e"
$C5 > gpurun_out/p2_c5_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:fast_logmel -s 3 -c 1 -f -o gpurun_out/r02b_k1_1024 $C5 > gpurun_out/p2_ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
C5I="python bench.py --workload vocos_istft --clips 1024 --steps 3 --no-cpu-baseline --no-e2e"
$C5I > gpurun_out/p2_c5i_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:fast_istft -s 3 -c 1 -f -o gpurun_out/r02b_k3_1024 $C5I > gpurun_out/p2_ncu_c5i.log 2>&1; echo "ncu c5i rc=$?"
