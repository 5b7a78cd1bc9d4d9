# round-2 state after the container was re-created: GPU suite, ncu --set full of the final kernels, traffic file, default bench, launch list
mkdir -p gpurun_out
S=$(date +%s)
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02y_pytest.log; tail -3 gpurun_out/r02y_pytest.log
echo "pytest wall=$(( $(date +%s) - S ))s"
python profiles/ncu_target.py > gpurun_out/r02y_target.log 2>&1; tail -3 gpurun_out/r02y_target.log
ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_leaf|k_colsum|k_reduce" -c 16 -o gpurun_out/r02y_final -f python profiles/ncu_target.py > gpurun_out/r02y_ncu.log 2>&1; tail -3 gpurun_out/r02y_ncu.log
python profiles/make_ncu_traffic.py gpurun_out/r02y_final.ncu-rep > gpurun_out/r02y_traffic.log 2>&1
cp profiles/ncu_traffic.json gpurun_out/r02y_ncu_traffic.json
python profiles/summarize_ncu.py gpurun_out/r02y_final.ncu-rep > gpurun_out/r02y_final_kernels_ncu.txt
S=$(date +%s)
python bench.py > gpurun_out/r02y_bench_default.json 2> gpurun_out/r02y_bench_default.err
echo "bench default rc=$? wall=$(( $(date +%s) - S ))s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02y_bench_default.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'dropin', (d.get('e2e_dropin') or {}).get('value'))
print('kernels', d['kernels'])
print('roofline', d['roofline'])
print('cpu', d.get('cpu_baseline'))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02y_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star > gpurun_out/r02y_launches.log 2>&1
tail -2 gpurun_out/r02y_launches.log
