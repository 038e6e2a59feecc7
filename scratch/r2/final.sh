#!/bin/bash
# round-2 final validation on one GPU: full suite, smoke, default bench line (all workloads, CPU baselines), reference arm, ncu launch list
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; echo "reference rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/final_bench.json'))
print('headline ms %.3f value %.0f frac %.3f e2e %.1f e2e16 %.1f cpu %.2f' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e_i16_f16']['value'], d['cpu_baseline']['value']))
for k,v in d['workloads'].items():
    print(k, 'ms %.3f value %.0f bound %s binding_frac %.3f e2e %s cpu %s' % (v['ms_per_step'], v['value'], v['roofline']['bound'], v['roofline']['binding_frac'], v.get('e2e',{}).get('value'), (v.get('cpu_baseline') or {}).get('value')))
print(d['variants']); print(d['e2e_api'])
r=json.load(open('gpurun_out/final_reference.json')); print('reference arm', r['value'], r['cpu_baseline']['cores'])
P
CMD="python bench.py --clips 512 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/final_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv $CMD > gpurun_out/final_ncu_l.log 2>&1; echo "ncu list rc=$?"
