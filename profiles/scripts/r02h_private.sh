mkdir -p gpurun_out
python -m pytest tests/test_gpu_private.py -x -q 2>&1 | tail -40 > gpurun_out/r02h_private.log; tail -30 gpurun_out/r02h_private.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02h_pytest.log; tail -5 gpurun_out/r02h_pytest.log
