#!/bin/bash
# prof_one.sh <out name> <kernel regex> <bench.py workload> <clips> : ncu --set full capture (with source) of one bench.py workload's kernel
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
W="python bench.py --workload $3 --clips $4 --steps 3 --no-cpu-baseline --no-e2e"
$W > gpurun_out/p1_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/$1 $W > gpurun_out/p1_ncu.log 2>&1; echo "ncu rc=$?"
