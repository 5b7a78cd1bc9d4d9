# block size x L2-prefetch distance of the walk kernel, then the whole test suite and one full default bench line
mkdir -p gpurun_out
Q="--steps 20 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star"
for t in 32 64 128 256; do
  lib=""; [ $t != 128 ] && lib="$PWD/build/variants/libsieve_b200_T$t.so"
  for d in 0 4; do
    SIEVE_B200_LIB=$lib SIEVE_B200_PF_DIST=$d python bench.py $Q > gpurun_out/r02c_T${t}_pf$d.json 2> gpurun_out/r02c_T${t}_pf$d.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02c_T*_pf*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f ms/step %.3f e2e %.2f prune %.3f tree-only %.3f leaf %.3f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernels']['prune_ms'], d['kernels']['tree_only_eval_ms'], d['kernels']['leaf_ms']), d['parity']['max_rel_site_logl'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02c_pytest.log
tail -4 gpurun_out/r02c_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench_full.json 2> gpurun_out/r02c_bench_full.err
tail -c 3000 gpurun_out/r02c_bench_full.json; tail -5 gpurun_out/r02c_bench_full.err
