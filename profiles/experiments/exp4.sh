mkdir -p gpurun_out/exp4
for mb in 0 7 8; do
  SIEVE_B200_LEAF_TAB=1 SIEVE_B200_PRUNE_MINB=$mb python bench.py --steps 8 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp4/mb${mb}.json 2> gpurun_out/exp4/mb${mb}.err
done
for mb in 0 7 8; do
  SIEVE_BENCH_SITES=125000 SIEVE_B200_LEAF_TAB=1 SIEVE_B200_PRUNE_MINB=$mb python bench.py --steps 8 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp4/s125k_mb${mb}.json 2> gpurun_out/exp4/s125k_mb${mb}.err
done
