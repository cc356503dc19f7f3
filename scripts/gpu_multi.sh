#!/bin/bash
# multi-GPU visit: parity check + bench at N GPUs.  usage: gpu_multi.sh <tag> <N>
tag=${1:-mg}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/${tag}_check.log 2>&1
tail -6 gpurun_out/${tag}_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${tag}_bench_n$N.json') if l.startswith('{')][-1])
    print('c4', d['n_gpus'], d['value'], d['ms_per_step'], 'parity', d['parity_check'])
    print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk!='cases'}) for k,v in d['parity'].items()})
    print('mode_a', d['mode_a'])
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/${tag}_bench_n$N.err').read()[-3000:])
PY
