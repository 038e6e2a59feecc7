#!/bin/bash
# 1 GPU: full GPU suite + default bench (after: polar hoist, silent-tile skip, fused head)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/c_pytest.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/c_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/c_bench.json'))
print('headline ms %.3f value %.0f frac %.3f e2e %.1f e2e16 %.1f' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e_i16_f16']['value']))
for k,v in d['workloads'].items():
    print(k, 'ms %.3f value %.0f bound %s binding_frac %.3f hbm_frac %.3f e2e %s' % (v['ms_per_step'], v['value'], v['roofline']['bound'], v['roofline']['binding_frac'], v['roofline']['frac'], v.get('e2e',{}).get('value')))
print(d['variants'])
P
