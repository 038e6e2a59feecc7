#!/bin/bash
# round 2: first run of the warp-specialised kernel — parity subset, A/B timing, ncu full capture with source
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "whisper or qwen or funasr or glm or s3tok or hf" > gpurun_out/ws1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/ws1_pytest.log
tail -5 gpurun_out/ws1_pytest.log
for ws in 0 1; do
  B2A_WS=$ws timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ws1_bench_ws$ws.json 2> gpurun_out/ws1_bench_ws$ws.err
  echo "ws=$ws rc=$?"; python -c "import json;d=json.load(open('gpurun_out/ws1_bench_ws$ws.json'));print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
B2A_WS=1 timeout 300 python bench.py --clips 512 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ws1_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fast_logmel_ws -s 3 -c 1 -o gpurun_out/ws1_prof -f python bench.py --clips 512 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ws1_ncu.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out | tail -12
