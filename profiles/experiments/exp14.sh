mkdir -p gpurun_out/exp14
for T in 8 16 32; do
  SIEVE_B200_TRANSIENT_LEAVES=$T python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/exp14/T$T.json 2> gpurun_out/exp14/T$T.err
done
