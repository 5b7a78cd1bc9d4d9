# one site per lane in the walk (NS = 1): 58 registers -> 8 blocks/SM, capped at 56 / 51 -> 9 / 10 blocks/SM
mkdir -p gpurun_out
Q="--steps 20 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star"
run() { # tag lib minb
  SIEVE_B200_LIB=$2 SIEVE_B200_PRUNE_MINB=$3 python bench.py $Q > gpurun_out/r02aa_$1.json 2> gpurun_out/r02aa_$1.err
}
run base "" ""
run ns1_b8 $PWD/build/variants/libsieve_b200_ns1m10.so 8
run ns1_b10 $PWD/build/variants/libsieve_b200_ns1m10.so 6
run ns1_b9 $PWD/build/variants/libsieve_b200_ns1m9.so 6
run ns1_b7 $PWD/build/variants/libsieve_b200_ns1m9.so 7
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02aa_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f ms/step %.3f prune %.3f tree-only %.3f leaf %.3f'%(d['value'], d['ms_per_step'], d['kernels']['prune_ms'], d['kernels']['tree_only_eval_ms'], d['kernels']['leaf_ms']), d['parity']['max_rel_site_logl'], d['parity']['logP'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
