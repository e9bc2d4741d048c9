#!/bin/bash
# compute-sanitizer memcheck over the small-shape GPU tests of the kernels added this round
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 \
    python -m pytest -q -m gpu --timeout 1400 -x \
    "tests/test_gpu_midx.py::test_kmeans_matches_reference_golden" \
    "tests/test_gpu_midx.py::test_segment_cdf_and_search_vs_oracle" \
    "tests/test_gpu_midx.py::test_sampler_update_and_draw_match_reference_golden" \
    "tests/test_gpu_midx.py::test_construct_index_is_a_stable_sort[1000-37]" \
    "tests/test_gpu_midx.py::test_construct_index_is_a_stable_sort[2049-256]" \
    "tests/test_gpu_midx.py::test_kmeans_assign_and_update_vs_oracle" \
    "tests/test_gpu_sampling_methods.py::test_masked_uniform_bit_exact[case0]" \
    "tests/test_gpu_sampling_methods.py::test_masked_uniform_bit_exact[case3]" \
    "tests/test_gpu_sampling_methods.py::test_score_ids_streaming_kernel" \
    "tests/test_gpu_sampling_methods.py::test_sampling_methods_replay_reference_golden" \
    "tests/test_gpu_pair.py::test_golden_steps" \
    "tests/test_gpu_pair.py::test_kernel_variants_match_golden" \
    "tests/test_gpu_pair.py::test_edge_shapes_vs_oracle" \
    "tests/test_gpu_rowopt.py::test_apply_matches_rows" \
    > gpurun_out/sanitizer.log 2>&1
echo "sanitizer exit $?" >> gpurun_out/sanitizer.log
grep -E "ERROR SUMMARY|Invalid|passed|failed|sanitizer exit|Error" gpurun_out/sanitizer.log | head -20
