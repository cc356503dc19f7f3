#!/bin/bash
# compute-sanitizer over a subset of the GPU parity tests (the reference has no race detection; its
# histogram kernel syncs inside divergent control flow and its accept counters rely on stream order).
# Run on the GPU box:   gpurun --timeout 1500 -- 'bash scripts/sanitize.sh'
# Logs land in gpurun_out/sanitizer_<tool>.log; copy the summaries into profiles/.
set -u
mkdir -p gpurun_out
TESTS=${TESTS:-"tests/test_gpu_parity.py::test_golden_uniform_positions tests/test_gpu_parity.py::test_golden_raw_launches_and_accu_step tests/test_gpu_parity.py::test_histogram_counts_bit_exact_with_cycles tests/test_gpu_parity.py::test_histogram_and_moments_multidim tests/test_gpu_parity.py::test_acor_fixtures tests/test_gpu_glm.py::test_glm_steplocked_vs_oracle tests/test_gpu_next.py tests/test_gpu_parity.py::test_dataset_histogram_repeated_bins tests/test_gpu_parity.py::test_dataset_engine_ragged_shapes tests/test_gpu_parity.py::test_quadform_tensor_core_move_vs_oracle tests/test_gpu_rowadd.py::test_gaussian_posterior_logdensity_and_moves_match_oracle"}
for tool in memcheck racecheck; do
    timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
        python -m pytest $TESTS -x -q -m gpu > gpurun_out/sanitizer_$tool.log 2>&1
    echo "$tool exit=$?" | tee -a gpurun_out/sanitizer_$tool.log
    grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3
done
