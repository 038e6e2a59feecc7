mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
tail -c 300 gpurun_out/bench_${n}gpu.err; cut -c1-260 gpurun_out/bench_${n}gpu.json; echo
done
