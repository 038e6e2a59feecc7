#!/bin/bash
# 2 GPUs: the real multi-rank tests (NCCL), then the default bench line at N = 2 (strong + weak + frame-sharded C3)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py tests/test_gpu_reentrancy.py -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/b_pytest.log
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/b_bench_n2.json 2> gpurun_out/b_bench_n2.err; echo "bench n2 rc=$?"; grep -c "NCCL INFO" gpurun_out/b_bench_n2.err; tail -3 gpurun_out/b_bench_n2.err | cut -c1-300
python - <<'P'
import json
d=json.load(open('gpurun_out/b_bench_n2.json'))
print('headline', d['scaling'], 'ms %.3f value %.0f frac %.3f e2e %.1f e2e16 %.1f weak %s' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e_i16_f16']['value'], d['weak']))
print(d['e2e_i16_f16'].get('pinned_copy_gbs_all_ranks_active'))
for k,v in d['workloads'].items():
    print(k, 'ms %.3f value %.0f binding_frac %.3f e2e %s' % (v['ms_per_step'], v['value'], v['roofline']['binding_frac'], v.get('e2e',{}).get('value')), v.get('collective',''))
P
