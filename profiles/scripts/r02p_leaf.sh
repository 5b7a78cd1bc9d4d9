mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02p_pytest.log; tail -6 gpurun_out/r02p_pytest.log
Q="--steps 20 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star"
run() { name=$1; shift; env "$@" python bench.py $Q > gpurun_out/r02p_$name.json 2> gpurun_out/r02p_$name.err; }
run analytic A=1
run table SIEVE_B200_LEAF_ANALYTIC=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02p_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f ms/step %.3f e2e %.2f prune %.3f tree-only %.3f leaf %.3f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernels']['prune_ms'], d['kernels']['tree_only_eval_ms'], d['kernels']['leaf_ms']), d['parity']['max_rel_site_logl'], d['parity']['max_rel_site_const'], d['parity']['logP'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
