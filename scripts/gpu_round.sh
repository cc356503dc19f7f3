#!/bin/bash
# One GPU-box visit: GPU tests, smoke, the default bench line and the other single-GPU workloads.
# Usage (from the repo root):  gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [tag]'
tag=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
(time python -m pytest tests -m gpu -q --durations=15) > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
python bench.py > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err; tail -c 1500 gpurun_out/${tag}_bench_c4.json
for w in c5 c3 c1; do
  python bench.py --workload $w --no-cpu-baseline --no-parity-check --steps 20 > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  head -c 400 gpurun_out/${tag}_bench_$w.json; echo
done
python bench.py --workload summary --steps 20 > gpurun_out/${tag}_bench_summary.json 2>&1; cat gpurun_out/${tag}_bench_summary.json
