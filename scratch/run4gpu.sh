mkdir -p gpurun_out
nproc; python -c "import os; print(len(os.sched_getaffinity(0)))"; cat /sys/devices/system/node/node*/cpulist; nvidia-smi topo -m 2>/dev/null | head -14
for n in 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${n}gpu_numa.json 2> gpurun_out/bench_${n}gpu_numa.err
tail -c 300 gpurun_out/bench_${n}gpu_numa.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_${n}gpu_numa.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['e2e'])"
done
