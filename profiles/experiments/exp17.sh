mkdir -p gpurun_out/exp17
timeout 100 python -m pytest tests -m gpu -x -q -k "pipelined_walk or sparse_residency" > gpurun_out/exp17/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp17/pytest.log
for t in 2 3 4; do
  SIEVE_B200_TMA=$t timeout 60 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp17/tma$t.json 2> gpurun_out/exp17/tma$t.err
done
