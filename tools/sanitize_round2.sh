# compute-sanitizer over the round-2 kernels (run on the GPU box): small-shape parity tests only (memcheck is 10-50x slower)
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
K='lowres or unaligned or scatter or feature_sources or reduction_none or shared_pass_equals or losses_match_oracle or proto_labeller_matches_oracle or selectors_match_oracle or grouped_launches or invalid_ids or dense_kernels_match or regime_switch or batched_call'
timeout 1500 $S --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_acquisition.py tests/test_gpu_losses.py tests/test_gpu_losses_dense.py tests/test_gpu_labeller.py tests/test_gpu_scatter_compat.py -m gpu -q -x -k "$K" > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/r2_memcheck.log
grep -E "passed|failed|ERROR SUMMARY|memcheck exit" gpurun_out/r2_memcheck.log | tail -5
timeout 900 $S --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_losses.py tests/test_gpu_losses_dense.py tests/test_gpu_labeller.py -m gpu -q -x -k "golden or dense_kernels_match" > gpurun_out/r2_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/r2_racecheck.log
grep -E "passed|failed|RACECHECK SUMMARY|racecheck exit" gpurun_out/r2_racecheck.log | tail -5
