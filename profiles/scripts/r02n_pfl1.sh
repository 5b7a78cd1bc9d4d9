mkdir -p gpurun_out
Q="--steps 20 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star"
run() { name=$1; shift; env "$@" python bench.py $Q > gpurun_out/r02n_$name.json 2> gpurun_out/r02n_$name.err; }
run base A=1
for d in 2 3 4 6 8; do run pfl1_d$d SIEVE_B200_LIB=$PWD/build/variants/libsieve_b200_PFL1.so SIEVE_B200_PF_DIST=$d; done
run base_d5 SIEVE_B200_PF_DIST=5
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02n_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f ms/step %.3f e2e %.2f prune %.3f tree-only %.3f leaf %.3f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernels']['prune_ms'], d['kernels']['tree_only_eval_ms'], d['kernels']['leaf_ms']), d['parity']['max_rel_site_logl'], d['parity']['logP'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
