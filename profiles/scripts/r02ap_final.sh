# final state of round 2: GPU suite, smoke, ncu --set full of the final kernels + traffic file for THIS build, default bench, launch list
mkdir -p gpurun_out
S=$(date +%s)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02ap_pytest.log; tail -2 gpurun_out/r02ap_pytest.log
echo "pytest wall=$(( $(date +%s) - S ))s"
python __graft_entry__.py smoke 2>&1 | tail -2
python profiles/ncu_target.py > gpurun_out/r02ap_target.log 2>&1; tail -2 gpurun_out/r02ap_target.log
ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_leaf|k_colsum|k_reduce" -c 16 -o gpurun_out/r02ap_final -f python profiles/ncu_target.py > gpurun_out/r02ap_ncu.log 2>&1; tail -1 gpurun_out/r02ap_ncu.log
python profiles/make_ncu_traffic.py gpurun_out/r02ap_final.ncu-rep > gpurun_out/r02ap_traffic.log 2>&1
cp profiles/ncu_traffic.json gpurun_out/r02ap_ncu_traffic.json
python profiles/summarize_ncu.py gpurun_out/r02ap_final.ncu-rep > gpurun_out/r02ap_final_kernels_ncu.txt
S=$(date +%s)
python bench.py > gpurun_out/r02ap_bench_default.json 2> gpurun_out/r02ap_bench_default.err
echo "bench default rc=$? wall=$(( $(date +%s) - S ))s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ap_bench_default.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'dropin', (d.get('e2e_dropin') or {}).get('value'))
print('kernels', d['kernels'])
print('roofline', d['roofline'])
print('mcmc', {k:round(v['states_per_s']) for k,v in d['mcmc'].items()})
ns=d['north_star']; print('ns', ns['ms_per_evaluation'], ns['leaf_stress']['leaf_ms'], {k:round(v['states_per_s']) for k,v in ns['mcmc'].items()})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|k_walk|k_leaf|k_prune" -c 120 --csv --log-file gpurun_out/r02ap_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star > gpurun_out/r02ap_launches.log 2>&1
SIEVE_B200_GENERIC=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('generic kernel: value %.1f prune %.3f'%(d['value'], d['kernels']['prune_ms']))"
