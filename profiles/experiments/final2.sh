mkdir -p gpurun_out/final2
python -m pytest tests -m gpu -x -q > gpurun_out/final2/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final2/pytest.log
( time python bench.py > gpurun_out/final2/bench_default.json ) 2> gpurun_out/final2/bench_default.err
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final2/bench_ref.json ) 2> gpurun_out/final2/bench_ref.err
