#!/bin/bash
# ncu --set full captures (with source) of the two generic kernels on the shapes of bench_configs section G
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
G="python benchmarks/bench_configs.py --only G --steps 2"
$G > gpurun_out/pg_plain.log 2>&1; echo "plain rc=$?"; tail -7 gpurun_out/pg_plain.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:frontend_generic -c 1 -f -o gpurun_out/r02_k2_fwd_1920 $G > gpurun_out/pg_ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_generic -c 1 -f -o gpurun_out/r02_k3b_inv_2048 $G > gpurun_out/pg_ncu_inv.log 2>&1; echo "ncu inv rc=$?"
