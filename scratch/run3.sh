mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x -k "stft" 2>&1 | tail -5
python benchmarks/bench_configs.py --only S --steps 6 2>&1 | tee gpurun_out/stft_fast.txt | cut -c1-420
B2A_NO_FAST_STFT=1 python benchmarks/bench_configs.py --only S --steps 3 2>&1 | tee gpurun_out/stft_generic.txt | cut -c1-420
