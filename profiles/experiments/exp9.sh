mkdir -p gpurun_out/exp9
for sh in 0 1 2; do
  SIEVE_B200_LEAF_SHAPE=$sh python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp9/shape$sh.json 2> gpurun_out/exp9/shape$sh.err
done
SIEVE_B200_LEAF_SHAPE=1 python -m pytest tests -m gpu -x -q -k "leaf_scale_sums or full_evaluation or streamed_evaluation or leaf_kernel" > gpurun_out/exp9/pytest_s1.log 2>&1
SIEVE_B200_LEAF_SHAPE=2 python -m pytest tests -m gpu -x -q -k "leaf_scale_sums or full_evaluation or streamed_evaluation or leaf_kernel" > gpurun_out/exp9/pytest_s2.log 2>&1
