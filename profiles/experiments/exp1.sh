mkdir -p gpurun_out/exp1
for s in 56832 85248 100000 113664 125000 170496; do
  SIEVE_BENCH_SITES=$s python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp1/s$s.json 2> gpurun_out/exp1/s$s.err
done
