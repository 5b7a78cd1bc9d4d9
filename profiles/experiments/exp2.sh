mkdir -p gpurun_out/exp2
python -m pytest tests -m gpu -x -q > gpurun_out/exp2/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp2/pytest.log
for sh in 0 1 2; do
  SIEVE_B200_LEAF_SHAPE=$sh python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp2/shape$sh.json 2> gpurun_out/exp2/shape$sh.err
done
