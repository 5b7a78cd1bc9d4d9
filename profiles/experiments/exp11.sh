mkdir -p gpurun_out/exp11
python -m pytest tests -m gpu -x -q > gpurun_out/exp11/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp11/pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp11/s100k.json 2> gpurun_out/exp11/s100k.err
SIEVE_BENCH_SITES=125000 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/exp11/s125k.json 2> gpurun_out/exp11/s125k.err
