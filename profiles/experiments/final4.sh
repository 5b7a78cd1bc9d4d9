mkdir -p gpurun_out/final5
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/final5/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final5/pytest.log
