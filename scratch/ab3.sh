mkdir -p gpurun_out
for n in "$@"; do
  if [ "$n" = orig ]; then L=/root/repo/mlx_audio_plus_b200/lib/libb200audio.so; else L=/root/repo/mlx_audio_plus_b200/lib/lib_$n.so; fi
  for i in 1 2; do
    B2A_LIB=$L python bench.py --clips 4096 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$n', d['ms_per_step'], d['roofline']['kernel_ms'])"
  done
done 2>&1 | tee gpurun_out/ab3.txt
