# walk micro-variants: op descriptor as one 256-bit load; binary exponents with ld/st.global instead of generic LD/ST; both
mkdir -p gpurun_out
Q="--steps 30 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star"
for v in base desc256 gs gsd base2; do
  lib=""; case $v in base|base2) ;; *) lib=$PWD/build/variants/libsieve_b200_$v.so;; esac
  SIEVE_B200_LIB=$lib python bench.py $Q > gpurun_out/r02af_$v.json 2> gpurun_out/r02af_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02af_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f ms/step %.4f prune %.4f tree-only %.4f leaf %.3f'%(d['value'], d['ms_per_step'], d['kernels']['prune_ms'], d['kernels']['tree_only_eval_ms'], d['kernels']['leaf_ms']), d['parity']['max_rel_site_logl'], d['parity']['logP'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
