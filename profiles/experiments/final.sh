mkdir -p gpurun_out/final
python -m pytest tests -m gpu -x -q > gpurun_out/final/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final/smoke.log
( time python bench.py > gpurun_out/final/bench_default.json ) 2> gpurun_out/final/bench_default.err
python profiles/ncu_target.py > gpurun_out/final/target_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:^k_(leaf|prune|colsum|reduce)$' -f -o gpurun_out/prof_r01j python profiles/ncu_target.py > gpurun_out/final/ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_' -c 400 --csv --log-file gpurun_out/r01j_launches_bench_steps3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mcmc-states 0 > gpurun_out/final/ncu_list.log 2>&1
( time SIEVE_BENCH_CELLS=10000 SIEVE_BENCH_SITES=125000 python bench.py --sparse --steps 5 --warmup 3 --no-cpu-baseline --no-variants --mcmc-states 200 > gpurun_out/final/cfg4_sparse.json ) 2> gpurun_out/final/cfg4_sparse.err
