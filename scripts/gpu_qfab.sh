#!/bin/bash
# A/B of the walker-partitioned quadratic-form move on the SAME box: in-tree library vs bayadera_b200/variants/libbay_oldqf.so
# (build the other variant first: nvcc ... -shared -o bayadera_b200/variants/libbay_oldqf.so <other tree>/bayadera_b200/csrc/engine.cu)
N=${1:-2}
mkdir -p gpurun_out
for v in new old new old; do
  if [ $v = old ]; then export BAYADERA_B200_LIB=$PWD/bayadera_b200/variants/libbay_oldqf.so; else unset BAYADERA_B200_LIB; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 2 --warmup 3 --no-parity-check --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v', d['mode_a']['value'], d['mode_a']['half_step_us'], 'c4', d['value'])
"
done
