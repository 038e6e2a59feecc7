rm -f gpurun_out/clk.txt
B2A_CLOCKS=gpurun_out/clk.txt python bench.py --clips 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
B2A_SMEM_PAD=30000 B2A_CLOCKS=gpurun_out/clk.txt python bench.py --clips 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
awk '!seen[substr($0,1,60)]++' gpurun_out/clk.txt | cut -c40-200; sort -u gpurun_out/clk.txt | cut -c60-200 | head -8
