python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02b_pytest.log
for d in 0 2 4 8 16; do
  SIEVE_B200_PF_DIST=$d python bench.py --steps 20 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/r02b_bench_pf$d.json 2> gpurun_out/r02b_bench_pf$d.err
done
for d in 0 4; do
  SIEVE_B200_PF_DIST=$d SIEVE_BENCH_CELLS=10000 SIEVE_BENCH_SITES=125000 python bench.py --sparse --steps 5 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants > gpurun_out/r02b_cfg4_pf$d.json 2> gpurun_out/r02b_cfg4_pf$d.err
done
tail -5 gpurun_out/r02b_pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02b_*pf*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f ms/step %.3f e2e %.2f'%(d['value'], d['ms_per_step'], d['e2e']['value']), d['kernels'], d['parity'])
    except Exception as e:
        print(f, 'ERR', e)
PY
