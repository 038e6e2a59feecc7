python -m pytest tests -q -m gpu -x -k "whisper or logmel or framing or stft" 2>&1 | tail -3
rm -f gpurun_out/clk.txt
run() { echo "== $*" >> gpurun_out/clk.txt; env "$@" B2A_CLOCKS=gpurun_out/clk.txt python bench.py --clips 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1; env "$@" python bench.py --clips 4096 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms', d['ms_per_step'])" >> gpurun_out/clk.txt; }
run A=1
run B2A_NO_MELSPEC=1
awk '!seen[$0]++' gpurun_out/clk.txt | cut -c1-160 | grep -v "^fast" ; grep -A1 "^==" gpurun_out/clk.txt | cut -c60-200
