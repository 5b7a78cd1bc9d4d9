mkdir -p gpurun_out/exp10
python -m pytest tests -m gpu -x -q > gpurun_out/exp10/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp10/pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp10/b.json 2> gpurun_out/exp10/b.err
SIEVE_B200_LEAF_TAB=2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp10/b_tab2.json 2> gpurun_out/exp10/b_tab2.err
