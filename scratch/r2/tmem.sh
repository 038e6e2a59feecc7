#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 120 ./scratch/r2/ubench_tmem > gpurun_out/ubench_tmem.log 2>&1
echo rc=$?; cat gpurun_out/ubench_tmem.log
