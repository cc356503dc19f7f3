#!/bin/bash
tag=${1:-glm}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_glm.py -m gpu -q -s > gpurun_out/${tag}_pytest.log 2>&1
grep -E "dlogp|level|accept dec|passed|failed|FAILED|Error" gpurun_out/${tag}_pytest.log | head -20
cat > /tmp/show.py <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['parity'], d['clocks'], 'frac', d['roofline']['frac'])
PY
timeout 400 python bench.py --no-cpu-baseline --no-mode-a > gpurun_out/${tag}_bench_c4_tA.json 2> gpurun_out/${tag}_bench_c4_tA.err
python /tmp/show.py gpurun_out/${tag}_bench_c4_tA.json terms4 || tail -5 gpurun_out/${tag}_bench_c4_tA.err
BAY_GLM_TERMS=5 timeout 400 python bench.py --no-cpu-baseline --no-mode-a > gpurun_out/${tag}_bench_c4_tB.json 2> gpurun_out/${tag}_bench_c4_tB.err
python /tmp/show.py gpurun_out/${tag}_bench_c4_tB.json terms5 || tail -5 gpurun_out/${tag}_bench_c4_tB.err
