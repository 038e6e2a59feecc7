#!/bin/bash
# N GPUs (gpurun --gpus N): multi-rank tests + default bench line at N ranks.  usage: run_n.sh N
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=$1
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/n${N}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/n${N}_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/n${N}_bench.err | cut -c1-300
python - $N <<'P'
import json,sys
N=sys.argv[1]
d=json.load(open(f'gpurun_out/n{N}_bench.json'))
print('headline', d['scaling'], 'n', d['n_gpus'], 'ms %.3f value %.0f frac %.3f e2e %.1f e2e16 %.1f' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e_i16_f16']['value']), 'weak', d['weak'])
print(d['e2e_i16_f16'].get('pinned_copy_gbs_all_ranks_active'))
for k,v in d['workloads'].items():
    print(k, 'ms %.3f value %.0f binding_frac %.3f e2e %s' % (v['ms_per_step'], v['value'], v['roofline']['binding_frac'], v.get('e2e',{}).get('value')), v.get('collective',''))
print(d['variants'])
P
