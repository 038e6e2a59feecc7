#!/bin/bash
# 1 GPU: failed tests re-run + ncu --set full of the production K1 kernel (source page) + C3 / C5 kernels
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "float64_truth or c4_full or c5_full or silent" > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/d_pytest.log
CMD="python bench.py --clips 512 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/d_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:fast_logmel -s 3 -c 1 -f -o gpurun_out/r02_k1_tma $CMD > gpurun_out/d_ncu_k1.log 2>&1; echo "ncu k1 rc=$?"; tail -2 gpurun_out/d_ncu_k1.log
C3="python bench.py --workload parakeet_64x1h --clips 4 --steps 3 --no-cpu-baseline --no-e2e"
$C3 > gpurun_out/d_c3_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:fast_logmel -s 3 -c 1 -f -o gpurun_out/r02_k1_512 $C3 > gpurun_out/d_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"; tail -2 gpurun_out/d_ncu_c3.log
C5="python bench.py --workload vocos_mel --clips 2048 --steps 3 --no-cpu-baseline --no-e2e"
$C5 > gpurun_out/d_c5_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:fast_logmel -s 3 -c 1 -f -o gpurun_out/r02_k1_1024 $C5 > gpurun_out/d_ncu_c5.log 2>&1; echo "ncu c5 rc=$?"; tail -2 gpurun_out/d_ncu_c5.log
C5I="python bench.py --workload vocos_istft --clips 1024 --steps 3 --no-cpu-baseline --no-e2e"
$C5I > gpurun_out/d_c5i_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:fast_istft -s 3 -c 1 -f -o gpurun_out/r02_k3_1024 $C5I > gpurun_out/d_ncu_c5i.log 2>&1; echo "ncu c5i rc=$?"; tail -2 gpurun_out/d_ncu_c5i.log
ls -la gpurun_out/*.ncu-rep | awk '{print $5,$9}'
