# substitutes for compute-sanitizer (closed on this pool): (1) the whole GPU suite with every pool poisoned (0xFF bytes: NaN /
# -1) at allocation -- a read of uninitialised device memory changes a result; (2) the same with the bounds-checked build
# (-DSV_CHECK_BOUNDS: every pool / table index of the walk and leaf kernels asserted on the device, __trap on failure)
mkdir -p gpurun_out
SIEVE_B200_POISON=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02m_poison.log; tail -2 gpurun_out/r02m_poison.log
SIEVE_B200_POISON=1 SIEVE_B200_LIB=$PWD/build/variants/libsieve_b200_checked.so python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02m_checked.log; tail -3 gpurun_out/r02m_checked.log
grep -c "SV_BOUND failed" gpurun_out/r02m_checked.log
