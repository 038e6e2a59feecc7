#!/bin/bash
# prof_g.sh <name> <kernel regex> <match text> : ncu --set full capture (with source) of one generic-kernel shape of bench_configs section G
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
while [ $# -ge 3 ]; do
G="python benchmarks/bench_configs.py --only G --steps 2 --match $3"
$G > gpurun_out/pg_plain_$1.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/pg_plain_$1.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f -o gpurun_out/$1 $G > gpurun_out/pg_ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
shift 3
done
