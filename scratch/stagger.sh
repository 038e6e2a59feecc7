for d in 0 1000 2000 3000 4000 5000 6000; do
echo -n "stagger $d: "; B2A_STAGGER=$d python bench.py --clips 4096 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms', d['ms_per_step'])"
done
