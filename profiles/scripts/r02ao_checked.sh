# memory-safety substitute on the FINAL code: whole GPU suite on the bounds-checked build with poisoned pools, and the product build with poisoned pools
mkdir -p gpurun_out
SIEVE_B200_POISON=1 SIEVE_B200_LIB=$PWD/build/variants/libsieve_b200_checked.so python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02ao_checked.log; tail -2 gpurun_out/r02ao_checked.log
SIEVE_B200_POISON=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02ao_poison.log; tail -2 gpurun_out/r02ao_poison.log
