mkdir -p gpurun_out/final3
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/final3/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final3/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final3/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final3/smoke.log
( time python bench.py > gpurun_out/final3/bench_default.json ) 2> gpurun_out/final3/bench_default.err
