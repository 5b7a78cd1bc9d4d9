# k_leaf_lin block shapes re-swept after the row compaction (threads x sites per thread, blocks per SM)
mkdir -p gpurun_out
Q="--steps 30 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star"
for v in base t256i4 t384i2 t512i3 t256i2 t1024i1 base2; do
  lib=""; case $v in base*) ;; *) lib=$PWD/build/variants/libsieve_b200_$v.so;; esac
  SIEVE_B200_LIB=$lib python bench.py $Q > gpurun_out/r02aj_$v.json 2> gpurun_out/r02aj_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02aj_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f ms/step %.4f prune %.4f leaf %.4f frac %.3f'%(d['value'], d['ms_per_step'], d['kernels']['prune_ms'], d['kernels']['leaf_ms'], d['kernels']['leaf_frac']), d['parity']['logP'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
