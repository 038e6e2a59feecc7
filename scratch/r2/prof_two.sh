#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
W="python bench.py --workload vocos_istft --clips 1024 --steps 3 --no-cpu-baseline --no-e2e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:istft16 -s 3 -c 1 -f -o gpurun_out/r02e_k3_16_2cta $W > gpurun_out/p2_ncu_a.log 2>&1; echo "ncu rc=$?"
B2A_X_CARVE=72 timeout 900 ncu --set full --clock-control none --import-source on -k regex:istft16 -s 3 -c 1 -f -o gpurun_out/r02e_k3_16_1cta $W > gpurun_out/p2_ncu_b.log 2>&1; echo "ncu rc=$?"
