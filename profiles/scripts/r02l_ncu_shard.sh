mkdir -p gpurun_out
Q="--steps 20 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star"
python bench.py $Q > gpurun_out/r02l_base.json 2> gpurun_out/r02l_base.err
python bench.py $Q > gpurun_out/r02l_base2.json 2> gpurun_out/r02l_base2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02l_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f ms/step %.3f e2e %.2f prune %.3f tree-only %.3f leaf %.3f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernels']['prune_ms'], d['kernels']['tree_only_eval_ms'], d['kernels']['leaf_ms']), d['parity']['max_rel_site_logl'], d['parity']['logP'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
python profiles/ncu_target.py 10000 125000 sparse > gpurun_out/r02l_target.log 2>&1; tail -3 gpurun_out/r02l_target.log
ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_leaf" -c 4 -o gpurun_out/r02l_shard -f python profiles/ncu_target.py 10000 125000 sparse > gpurun_out/r02l_ncu.log 2>&1; tail -3 gpurun_out/r02l_ncu.log
