#!/bin/bash
# A/B of the GLM epilogue variants (BAY_GLM_POLY8 = 0 / 2 (in-tree) / 4): accuracy test + bench each
tag=${1:-poly}
mkdir -p gpurun_out
cat > /tmp/show.py <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['parity'].get('glm_dlogp_vs_fp64'), d['clocks'], 'frac', d['roofline']['frac'])
PY
timeout 900 python -m pytest tests/test_gpu_glm.py -m gpu -q -s > gpurun_out/${tag}_pytest.log 2>&1
grep -E "dlogp|level|accept dec|passed|failed|FAILED|Error" gpurun_out/${tag}_pytest.log | head -20
for v in p2 p0 p4; do
  if [ $v = p2 ]; then unset BAYADERA_B200_LIB; else export BAYADERA_B200_LIB=$PWD/bayadera_b200/variants/libbay_$v.so; fi
  timeout 400 python bench.py --no-cpu-baseline --no-mode-a > gpurun_out/${tag}_bench_c4_$v.json 2> gpurun_out/${tag}_bench_c4_$v.err
  python /tmp/show.py gpurun_out/${tag}_bench_c4_$v.json $v || tail -5 gpurun_out/${tag}_bench_c4_$v.err
done
