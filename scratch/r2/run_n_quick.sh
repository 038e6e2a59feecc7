#!/bin/bash
# N GPUs, bounded: the multi-rank test for world = N only, then the default bench line.  usage: run_n_quick.sh N [skipbench]
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=$1
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "[$N]" > gpurun_out/q${N}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/q${N}_pytest.log
if [ -z "$2" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N > gpurun_out/q${N}_bench.json 2> gpurun_out/q${N}_bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/q${N}_bench.json'));print('n', d['n_gpus'], d['scaling'], 'ms %.3f value %.0f e2e %.1f e2e16 %.1f weak %.0f'%(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e_i16_f16']['value'], d['weak']['value']));print({k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['workloads'].items()})"
fi
