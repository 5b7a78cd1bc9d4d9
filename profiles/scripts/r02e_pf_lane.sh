mkdir -p gpurun_out
Q="--steps 20 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star"
for v in nopf def0 def2 def3 def4 def6 def8 def12; do
  lib=""; pf=0
  [ $v = nopf ] && lib="$PWD/build/variants/libsieve_b200_nopf.so"
  [ $v != nopf ] && pf=${v#def}
  SIEVE_B200_LIB=$lib SIEVE_B200_PF_DIST=$pf python bench.py $Q > gpurun_out/r02e_${v}.json 2> gpurun_out/r02e_${v}.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02e_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f ms/step %.3f e2e %.2f prune %.3f tree-only %.3f leaf %.3f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['kernels']['prune_ms'], d['kernels']['tree_only_eval_ms'], d['kernels']['leaf_ms']), d['parity']['max_rel_site_logl'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
SIEVE_B200_PF_DIST=4 ncu --set full --clock-control none --import-source on -k regex:k_prune -c 2 -o gpurun_out/r02e_prof_def4 -f python profiles/ncu_target.py > gpurun_out/r02e_ncu_def4.log 2>&1
