mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_parallel.py -q -m gpu 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 600 gpurun_out/bench_2gpu.err; cut -c1-700 gpurun_out/bench_2gpu.json
