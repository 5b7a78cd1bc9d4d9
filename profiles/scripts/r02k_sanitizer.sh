# compute-sanitizer over a reduced selection of the GPU tests (dense, sparse, streamed, max-sum, private model, size factors)
mkdir -p gpurun_out
SEL='tests/test_gpu_parity.py::test_golden_example_data tests/test_gpu_parity.py::test_smallest_trees tests/test_gpu_parity.py::test_streamed_ragged_chunks_are_deterministic tests/test_gpu_parity.py::test_reference_call_sequence_and_store_restore tests/test_gpu_parity.py::test_linear_partials_specific_behaviour tests/test_gpu_maxsum.py::test_trace_is_invalidated_by_changes_and_needs_resident_core tests/test_gpu_private.py::test_private_core_rejects_unsupported_modes tests/test_gpu_sizefactors.py::test_sharded_selection_equals_single_handle_sort'
SEL2='tests/test_gpu_parity.py::test_sparse_residency_is_bit_identical_to_dense tests/test_gpu_parity.py::test_streamed_evaluation_equals_resident tests/test_gpu_maxsum.py::test_max_sum_partials_and_back_pointers tests/test_gpu_private.py::test_accept_time_reestimation_matches_the_oracle'
for tool in memcheck racecheck initcheck synccheck; do
  S=$(date +%s)
  timeout 900 compute-sanitizer --tool $tool --target-processes all --error-exitcode 0 --print-limit 30 python -m pytest $SEL $( [ $tool = memcheck ] && echo $SEL2 ) -x -q -p no:cacheprovider > gpurun_out/r02k_$tool.log 2>&1
  echo "$tool rc=$? $(( $(date +%s) - S ))s: $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/r02k_$tool.log | tr '\n' ' ')"
done
