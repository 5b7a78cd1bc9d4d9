mkdir -p gpurun_out
Q="--steps 20 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star"
python bench.py $Q > gpurun_out/r02o_base.json 2> gpurun_out/r02o_base.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02o_base.json').read().strip().splitlines()[-1])
print(d['value'], d['kernels'])
PY
