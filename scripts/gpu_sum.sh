#!/bin/bash
tag=${1:-sum}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
tail -12 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --workload summary --steps 20 > gpurun_out/${tag}_bench_summary.json 2> gpurun_out/${tag}_bench_summary.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench_summary.json').read().strip().splitlines()[-1])
    for k,v in d['results'].items(): print(k, v)
except Exception as e:
    print('failed', e); print(open('gpurun_out/${tag}_bench_summary.err').read()[-2000:])
PY
