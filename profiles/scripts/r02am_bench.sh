mkdir -p gpurun_out
python bench.py --no-cpu-baseline --mcmc-states 100 --no-variants --no-dropin > gpurun_out/r02am_bench.json 2> gpurun_out/r02am_bench.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02am_bench.json").read().strip().splitlines()[-1]); ns=d["north_star"]
print(d["value"], d["e2e"]["value"], ns["residency"], ns["residency_fallbacks"], ns["ms_per_evaluation"], ns["leaf_stress"]["leaf_ms"])
PY
tail -3 gpurun_out/r02am_bench.err
