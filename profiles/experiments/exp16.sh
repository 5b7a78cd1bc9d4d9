mkdir -p gpurun_out/exp16
SIEVE_B200_TMA=3 timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:^k_prune_tma$' --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_r01k_tma python profiles/ncu_target.py > gpurun_out/exp16/ncu.log 2>&1
