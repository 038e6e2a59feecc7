#!/bin/bash
# 1 GPU: full GPU suite, default bench line (all workloads), reference arm, sanitizer runs, ncu launch list
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/a_pytest.log
timeout 900 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/a_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/a_bench.json'))
print('headline ms %.3f value %.0f frac %.3f e2e %.1f e2e16 %.1f cpu %.2f' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e_i16_f16']['value'], d['cpu_baseline']['value']))
for k,v in d['workloads'].items():
    print(k, 'ms %.3f value %.0f bound %s binding_frac %.3f hbm_frac %.3f e2e %s cpu %s' % (v['ms_per_step'], v['value'], v['roofline']['bound'], v['roofline']['binding_frac'], v['roofline']['frac'], v.get('e2e',{}).get('value'), (v.get('cpu_baseline') or {}).get('value')))
print(d['variants']); print(d['e2e_api'])
P
timeout 300 python tests/sanitize_cases.py > gpurun_out/a_sanitize_plain.log 2>&1; echo "sanitize plain rc=$?"; tail -3 gpurun_out/a_sanitize_plain.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitize_cases.py > gpurun_out/a_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/a_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/sanitize_cases.py k1_400 k1_512 k1c_stft k3_istft k4_small > gpurun_out/a_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/a_racecheck.log
CMD="python bench.py --clips 512 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/a_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/a_launches.csv $CMD > gpurun_out/a_ncu_l.log 2>&1; echo "ncu list rc=$?"
