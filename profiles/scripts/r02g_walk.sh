mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -x -q 2>&1 | tail -8 > gpurun_out/r02g_pytest.log; tail -4 gpurun_out/r02g_pytest.log
Q="--steps 20 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star"
run() { name=$1; shift; env "$@" python bench.py $Q > gpurun_out/r02g_$name.json 2> gpurun_out/r02g_$name.err; }
run walk A=1
run kprune SIEVE_B200_WALK=0
run walk_pf0 SIEVE_B200_PF_DIST=0
run walk_b7 SIEVE_B200_PRUNE_MINB=7
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02g_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f ms/step %.3f e2e %.2f prune %.3f tree-only %.3f leaf %.3f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernels']['prune_ms'], d['kernels']['tree_only_eval_ms'], d['kernels']['leaf_ms']), d['parity']['max_rel_site_logl'], d['parity']['logP'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
ncu --set full --clock-control none --import-source on -k regex:"k_walk" -c 2 -o gpurun_out/r02g_prof -f python profiles/ncu_target.py > gpurun_out/r02g_ncu.log 2>&1
