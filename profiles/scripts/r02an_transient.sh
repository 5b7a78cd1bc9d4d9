# transient threshold re-swept on the final walk: fewer stored partials (less DRAM write traffic) against more recomputation in local moves
mkdir -p gpurun_out
Q="--steps 30 --warmup 3 --no-cpu-baseline --mcmc-states 400 --no-variants --no-dropin --no-north-star"
for t in 8 12 16 24 32; do SIEVE_B200_TRANSIENT_LEAVES=$t python bench.py $Q > gpurun_out/r02an_tr$t.json 2> gpurun_out/r02an_tr$t.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02an_tr*.json'), key=lambda x:int(x.split('tr')[-1].split('.')[0])):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f[-9:], 'value %.2f prune %.3f tree-only %.3f'%(d['value'], d['kernels']['prune_ms'], d['kernels']['tree_only_eval_ms']), {k:round(v['states_per_s']) for k,v in d['mcmc'].items()})
PY
