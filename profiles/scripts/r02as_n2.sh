# the driver's 8-GPU command: weak cfg3 value + the north-star block (10 000 x 1 000 000 over 8 ranks, sparse, global size factors)
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02as_bench_n2.json 2> gpurun_out/r02as_bench_n2.err
echo "rc=$? wall=$(( $(date +%s) - S ))s"
tail -3 gpurun_out/r02as_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02as_bench_n2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
ns=d.get('north_star') or {}
print({k:ns.get(k) for k in ('sites_per_gpu','residency','ms_per_evaluation','dataset_evals_per_s','tree_only_ms','prune_ms','slots')}, ns.get('leaf_stress'))
print('ns mcmc', {k:v['states_per_s'] for k,v in (ns.get('mcmc') or {}).items()})
print('ns parity', ns.get('parity'))
print('parity', d.get('parity'))
PY
