#!/bin/bash
tag=${1:-ra}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rowadd.py tests/test_gpu_context.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
tail -25 gpurun_out/${tag}_pytest.log
