#!/bin/bash
# third-session captures: new parity tests, then ncu --set full (with source) of the kernels this session changed
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "phase_ranges or interior_copy" 2>&1 | tail -4
prof() {  # name, kernel regex, command...
  local name=$1 rx=$2; shift 2
  "$@" > gpurun_out/p3_plain_$name.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/p3_ncu_$name.log 2>&1; echo "ncu $name rc=$?"
}
prof r02f_k4_polar istft_small python bench.py --workload kokoro_istft --clips 1024 --steps 3 --no-cpu-baseline --no-e2e
prof r02f_k3_1024 fast_istft_kernel python bench.py --workload vocos_istft --clips 1024 --steps 3 --no-cpu-baseline --no-e2e
prof r02f_k2_fwd_1920 frontend_generic python benchmarks/bench_configs.py --only G --steps 2 --match s3gen
prof r02f_k3b_inv_1920 istft_generic python benchmarks/bench_configs.py --only G --steps 2 --match ISTFTCache
