# leaf read-back of underflowed genotypes, walk-kernel bit-identity, launch list of a bench step (kernels of the library only), host profile of the drop-in sequence
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "leaf_kernel_all_cells or walk_kernels_are or golden or private" 2>&1 | tail -5
python -m pytest tests/test_gpu_private.py -m gpu -x -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|k_walk|k_leaf|k_prune" -c 120 --csv --log-file gpurun_out/r02z_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star > gpurun_out/r02z_ncu.log 2>&1
tail -1 gpurun_out/r02z_ncu.log | cut -c1-200
SIEVE_B200_HOST_PROF=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-north-star > gpurun_out/r02z_dropin.json 2> gpurun_out/r02z_dropin.err
grep "sv_flush host" gpurun_out/r02z_dropin.err | tail -12
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02z_dropin.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'dropin', d['e2e_dropin']['value'], d['e2e_dropin']['ms_per_step'])
PY
