mkdir -p gpurun_out/exp3
for sh in 0 1; do for tab in 1 2; do for pf in 0 1; do
  SIEVE_B200_LEAF_SHAPE=$sh SIEVE_B200_LEAF_TAB=$tab SIEVE_B200_LEAF_PF=$pf python bench.py --steps 8 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp3/s${sh}_t${tab}_p${pf}.json 2> gpurun_out/exp3/s${sh}_t${tab}_p${pf}.err
done; done; done
