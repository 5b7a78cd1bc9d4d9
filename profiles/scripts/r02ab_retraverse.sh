# sieve_private_retraverse_changed: the new test, then the whole private and parity suites (k_prune's epilogue index changed)
python -m pytest tests/test_gpu_private.py -m gpu -x -q 2>&1 | tail -15
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
