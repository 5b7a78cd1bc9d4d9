# k_leaf_lin with the exponents packed into the mantissa doubles (32 + 16 byte rows, three 16-byte arrays) against the 48-byte rows
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -x -q -k "leaf or golden or full_evaluation or smallest or nan or shapes or table" 2>&1 | tail -3
Q="--steps 30 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star"
for v in packed unpacked packedb unpackedb; do
  lib=""; case $v in unpacked*) lib=$PWD/build/variants/libsieve_b200_unpacked.so;; esac
  SIEVE_B200_LIB=$lib python bench.py $Q > gpurun_out/r02ai_$v.json 2> gpurun_out/r02ai_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02ai_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f ms/step %.4f prune %.4f leaf %.4f frac %.3f'%(d['value'], d['ms_per_step'], d['kernels']['prune_ms'], d['kernels']['leaf_ms'], d['kernels']['leaf_frac']), d['parity']['max_rel_site_logl'], d['parity']['logP'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
