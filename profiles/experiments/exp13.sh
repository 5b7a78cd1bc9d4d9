mkdir -p gpurun_out/exp13
python -m pytest tests -m gpu -x -q -k "pipelined_walk or sparse_residency or streamed" > gpurun_out/exp13/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp13/pytest.log
for rp in 0 1 4; do
  SIEVE_B200_RP=$rp python bench.py --steps 8 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp13/s100k_rp${rp}.json 2> gpurun_out/exp13/s100k_rp${rp}.err
done
for rp in 1 4; do
  SIEVE_BENCH_SITES=125000 SIEVE_B200_RP=$rp python bench.py --steps 8 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp13/s125k_rp${rp}.json 2> gpurun_out/exp13/s125k_rp${rp}.err
done
