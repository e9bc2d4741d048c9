#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the shared-memory-heavy kernels added this round
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 \
    python -m pytest -q -m gpu --timeout 1400 -x \
    "tests/test_gpu_midx.py::test_kmeans_matches_reference_golden" \
    "tests/test_gpu_midx.py::test_construct_index_is_a_stable_sort[1000-37]" \
    "tests/test_gpu_midx.py::test_construct_index_is_a_stable_sort[2049-256]" \
    "tests/test_gpu_midx.py::test_kmeans_assign_and_update_vs_oracle" \
    "tests/test_gpu_sampling_methods.py::test_masked_uniform_bit_exact[case0]" \
    "tests/test_gpu_sampling_methods.py::test_masked_uniform_bit_exact[case3]" \
    "tests/test_gpu_pair.py::test_golden_steps" \
    > gpurun_out/sanitizer_race.log 2>&1
echo "sanitizer exit $?" >> gpurun_out/sanitizer_race.log
grep -E "RACECHECK SUMMARY|hazard|passed|failed|sanitizer exit|Error" gpurun_out/sanitizer_race.log | head -20
