mkdir -p gpurun_out
Q="--steps 20 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star"
run() { name=$1; shift; env "$@" python bench.py $Q > gpurun_out/r02i_$name.json 2> gpurun_out/r02i_$name.err; }
run base A=1
for v in NS1B10 NS1B12 NS3B4 NS4B3 T64 T256; do
  run $v SIEVE_B200_LIB=$PWD/build/variants/libsieve_b200_$v.so SIEVE_B200_PRUNE_MINB=6
done
for t in 12 16 24; do run tr$t SIEVE_B200_TRANSIENT_LEAVES=$t; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02i_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f ms/step %.3f e2e %.2f prune %.3f tree-only %.3f leaf %.3f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernels']['prune_ms'], d['kernels']['tree_only_eval_ms'], d['kernels']['leaf_ms']), d['parity']['max_rel_site_logl'], d['parity']['logP'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02i_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star > gpurun_out/r02i_ncu.log 2>&1
