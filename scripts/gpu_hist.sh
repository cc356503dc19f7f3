#!/bin/bash
# histogram kernels: parity tests, summary bench (lane-private R=4 / R=2 / atomic kernel), launch list.  usage: gpu_hist.sh <tag>
tag=${1:-hist}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next.py -m gpu -q -k "histogram or dataset or evidence or quadform" > gpurun_out/${tag}_pytest.log 2>&1
tail -5 gpurun_out/${tag}_pytest.log
for v in r4 r2 r16; do
  unset BAY_HIST_VARIANT BAY_HIST_ATOMS
  [ $v = r2 ] && export BAY_HIST_VARIANT=1
  [ $v = r16 ] && export BAY_HIST_VARIANT=2
  [ $v = atoms ] && export BAY_HIST_ATOMS=1
  timeout 600 python bench.py --workload summary --steps 20 > gpurun_out/${tag}_bench_summary_$v.json 2> gpurun_out/${tag}_bench_summary_$v.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench_summary_$v.json').read().strip().splitlines()[-1])
    for k,v in d['results'].items():
        if 'hist' in k: print('$v', k, v)
except Exception as e:
    print('failed', e); print(open('gpurun_out/${tag}_bench_summary_$v.err').read()[-2000:])
PY
done
unset BAY_HIST_VARIANT BAY_HIST_ATOMS
