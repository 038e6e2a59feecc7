for d in 0 3000 20000000; do
echo -n "stagger $d: "; B2A_STAGGER=$d python bench.py --clips 4096 --steps 5 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms', d['ms_per_step'])"
done
rm -f gpurun_out/clk.txt
for d in 0 3000; do B2A_STAGGER=$d B2A_CLOCKS=gpurun_out/clk.txt python bench.py --clips 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1; done
sort gpurun_out/clk.txt | uniq -c | cut -c1-200
