mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02u_pytest.log; tail -3 gpurun_out/r02u_pytest.log
SIEVE_B200_POISON=1 SIEVE_B200_LIB=$PWD/build/variants/libsieve_b200_checked.so python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02u_checked.log; tail -2 gpurun_out/r02u_checked.log
S=$(date +%s)
python bench.py > gpurun_out/r02u_bench_default.json 2> gpurun_out/r02u_bench_default.err
echo "bench default rc=$? wall=$(( $(date +%s) - S ))s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02u_bench_default.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'dropin', (d.get('e2e_dropin') or {}).get('value'))
print('kernels', d['kernels'])
print('roofline', d['roofline'])
print('mcmc', {k:(v['states_per_s'] if isinstance(v,dict) else v) for k,v in (d.get('mcmc_states_per_s') or {}).items()})
ns=d.get('north_star') or {}
print({k:ns.get(k) for k in ('sites_per_gpu','residency','ms_per_evaluation','share_evals_per_s','tree_only_ms','prune_ms')}, ns.get('leaf_stress'))
print('ns mcmc', {k:v['states_per_s'] for k,v in (ns.get('mcmc') or {}).items()})
print('cpu', d.get('cpu_baseline'))
PY
