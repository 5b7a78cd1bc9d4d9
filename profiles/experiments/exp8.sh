mkdir -p gpurun_out/exp8
python -m pytest tests -m gpu -x -q > gpurun_out/exp8/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp8/pytest.log
for chk in 0 1; do
  SIEVE_B200_LEAF_CHECKED=$chk python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp8/chk$chk.json 2> gpurun_out/exp8/chk$chk.err
done
