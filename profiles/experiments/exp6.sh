mkdir -p gpurun_out/exp6
python profiles/ncu_target.py > gpurun_out/exp6/target_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k 'regex:^k_(leaf|prune|colsum|reduce)$' -f -o gpurun_out/prof_r01i python profiles/ncu_target.py > gpurun_out/exp6/ncu_full.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp6/bench3.json 2> gpurun_out/exp6/bench3.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01i_launches_bench_steps3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp6/ncu_list.log 2>&1
