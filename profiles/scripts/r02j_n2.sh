mkdir -p gpurun_out
export NCCL_DEBUG=INFO
S=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02j_n2.json 2> gpurun_out/r02j_n2.err
echo "rc=$? wall=$(( $(date +%s) - S ))s"
grep -c "Init COMPLETE\|comm 0x" gpurun_out/r02j_n2.json gpurun_out/r02j_n2.err | head
tail -c 600 gpurun_out/r02j_n2.err
python - <<'PY'
import json
for line in open('gpurun_out/r02j_n2.json'):
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line)
        print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
        ns=d.get('north_star') or {}
        print({k:ns.get(k) for k in ('sites_per_gpu','residency','ms_per_evaluation','dataset_evals_per_s','tree_only_ms','prune_ms','setup_s')})
        print(ns.get('leaf_stress')); print(ns.get('parity'))
PY
