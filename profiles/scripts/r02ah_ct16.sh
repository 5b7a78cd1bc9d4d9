# k_leaf_lin with compact 16-byte cell-table rows (default now) against the 32-byte rows (-DSV_LEAF_CT32); leaf parity tests first
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "leaf or golden or full_evaluation or smallest" 2>&1 | tail -3
Q="--steps 30 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star"
for v in ct16 ct32 ct16b ct32b; do
  lib=""; case $v in ct32*) lib=$PWD/build/variants/libsieve_b200_ct32.so;; esac
  SIEVE_B200_LIB=$lib python bench.py $Q > gpurun_out/r02ah_$v.json 2> gpurun_out/r02ah_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02ah_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f ms/step %.4f prune %.4f leaf %.4f frac %.3f'%(d['value'], d['ms_per_step'], d['kernels']['prune_ms'], d['kernels']['leaf_ms'], d['kernels']['leaf_frac']), d['parity']['max_rel_site_logl'], d['parity']['logP'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
