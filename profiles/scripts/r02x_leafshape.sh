mkdir -p gpurun_out
Q="--steps 20 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star"
run() { name=$1; shift; env "$@" python bench.py $Q > gpurun_out/r02x_$name.json 2> gpurun_out/r02x_$name.err; }
run base A=1
for v in T384I2B3 T512I3 T320I2B4 T640I2B2; do run L$v SIEVE_B200_LIB=$PWD/build/variants/libsieve_b200_L$v.so; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02x_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f ms/step %.3f e2e %.2f prune %.3f leaf %.3f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernels']['prune_ms'], d['kernels']['leaf_ms']), d['parity']['max_rel_site_logl'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
