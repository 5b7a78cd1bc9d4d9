# k_walk with the q pairs requested one op ahead and a separate code path for a forwarded child 1 (-DSV_WALK_QPIPE)
mkdir -p gpurun_out
SIEVE_B200_LIB=$PWD/build/variants/libsieve_b200_qpipe.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -x -q -k "walk_kernels or lazy or full_evaluation or smallest or streamed or sparse or shapes or distance" 2>&1 | tail -3
Q="--steps 30 --warmup 3 --no-cpu-baseline --mcmc-states 300 --no-variants --no-dropin --no-north-star"
for v in qpipe qpipe2; do
  lib=""; case $v in qpipe*) lib=$PWD/build/variants/libsieve_b200_qpipe.so;; esac
  SIEVE_B200_LIB=$lib python bench.py $Q > gpurun_out/r02at_$v.json 2> gpurun_out/r02at_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02at_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.2f ms/step %.4f prune %.4f tree-only %.4f leaf %.4f'%(d['value'], d['ms_per_step'], d['kernels']['prune_ms'], d['kernels']['tree_only_eval_ms'], d['kernels']['leaf_ms']), {k:round(v['states_per_s']) for k,v in d['mcmc'].items()}, d['parity']['logP'])
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-400:])
PY
