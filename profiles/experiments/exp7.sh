mkdir -p gpurun_out/exp7
python -m pytest tests -m gpu -x -q -k "leaf_scale_sums or full_evaluation or streamed_evaluation" > gpurun_out/exp7/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp7/pytest.log
SIEVE_B200_LEAF_ACC=1 python -m pytest tests -m gpu -x -q -k "leaf_scale_sums or full_evaluation or streamed_evaluation or leaf_kernel" > gpurun_out/exp7/pytest_acc.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp7/pytest_acc.log
for acc in 0 1; do
  SIEVE_B200_LEAF_ACC=$acc python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp7/acc$acc.json 2> gpurun_out/exp7/acc$acc.err
done
