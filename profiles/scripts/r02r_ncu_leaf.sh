mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_leaf" -c 2 -o gpurun_out/r02r_leaf -f python profiles/ncu_target.py > gpurun_out/r02r_ncu.log 2>&1; tail -3 gpurun_out/r02r_ncu.log
