mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|k_walk|k_leaf|k_prune" -c 120 --csv --log-file gpurun_out/r02w_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star > gpurun_out/r02w_ncu.log 2>&1
Q="--steps 20 --warmup 3 --no-cpu-baseline --mcmc-states 200 --no-variants --no-dropin --no-north-star"
for t in 8 12 16; do SIEVE_B200_TRANSIENT_LEAVES=$t python bench.py $Q > gpurun_out/r02w_tr$t.json 2> gpurun_out/r02w_tr$t.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02w_tr*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value %.2f prune %.3f tree-only %.3f'%(d['value'], d['kernels']['prune_ms'], d['kernels']['tree_only_eval_ms']), {k:round(v['states_per_s']) for k,v in d['mcmc'].items()})
PY
