#!/bin/bash
# prof.sh <lib path relative to repo> <kernel regex> <out tag> [ENV=VAL ...]: plain run, then one ncu --set full capture with source
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
lib="$1"; rx="$2"; tag="$3"; shift 3
env B2A_LIB="$PWD/$lib" "$@" timeout 300 python bench.py --clips 512 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_plain.log 2>&1 &&
env B2A_LIB="$PWD/$lib" "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -o gpurun_out/${tag} -f python bench.py --clips 512 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${tag}_ncu.log
