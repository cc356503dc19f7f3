#!/bin/bash
# quadform (c5) tensor-core kernel: parity tests + bench + ncu
tag=${1:-qf}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "mvn or quadform or mirror or persistent" > gpurun_out/${tag}_pytest.log 2>&1
tail -15 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --workload c5 --no-cpu-baseline --no-parity-check --steps 20 > gpurun_out/${tag}_bench_c5.json 2> gpurun_out/${tag}_bench_c5.err
head -c 300 gpurun_out/${tag}_bench_c5.json; tail -3 gpurun_out/${tag}_bench_c5.err
bash scripts/gpu_ncu.sh ${tag}_prof k_quadform 20 1 python bench.py --workload c5 --no-cpu-baseline --no-parity-check --steps 3 --warmup 3 > /dev/null 2>&1
