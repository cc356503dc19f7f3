#!/bin/bash
# c5 strong-scaling (mode A) only: usage gpu_modea.sh <tag> <N>
tag=${1:-ma}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload c5 --strong --steps 20 --warmup 3 --no-parity-check > gpurun_out/${tag}_c5_n$N.json 2> gpurun_out/${tag}_c5_n$N.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${tag}_c5_n$N.json') if l.startswith('{')][-1])
    print('c5 strong', d['n_gpus'], d['value'], 'ms/step', d['ms_per_step'], 'half-step us', d['ms_per_step']*1e3/8)
except Exception as e:
    print('failed', e); print(open('gpurun_out/${tag}_c5_n$N.err').read()[-2000:])
PY
