mkdir -p gpurun_out/exp15
timeout 120 python -m pytest tests -m gpu -x -q -k "pipelined_walk" > gpurun_out/exp15/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp15/pytest.log
for t in 0 2 3 4; do
  SIEVE_B200_TMA=$t timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp15/tma$t.json 2> gpurun_out/exp15/tma$t.err
done
