mkdir -p gpurun_out/exp12
for pf in 0 1; do for mb in 6 7 8; do
  SIEVE_B200_PRUNE_PF=$pf SIEVE_B200_PRUNE_MINB=$mb python bench.py --steps 8 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp12/s100k_pf${pf}_mb${mb}.json 2> gpurun_out/exp12/s100k_pf${pf}_mb${mb}.err
done; done
for pf in 0 1; do for mb in 6 7 8; do
  SIEVE_BENCH_SITES=125000 SIEVE_B200_PRUNE_PF=$pf SIEVE_B200_PRUNE_MINB=$mb python bench.py --steps 8 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp12/s125k_pf${pf}_mb${mb}.json 2> gpurun_out/exp12/s125k_pf${pf}_mb${mb}.err
done; done
