# closing check on the final commit, in the driver's order: GPU suite, smoke, reference arm, our arm
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('reference arm', d['value'], d['cpu_baseline']['cores'])"
python bench.py --gpus 1 > gpurun_out/r02aw_bench.json 2> gpurun_out/r02aw_bench.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02aw_bench.json').read().strip().splitlines()[-1])
print('value %.1f e2e %.1f dropin %.1f traffic %s clocks %s'%(d['value'], d['e2e']['value'], d['e2e_dropin']['value'], d['roofline']['traffic'], d['clocks']))
print('north star', d['north_star']['ms_per_evaluation'], d['north_star']['residency'], d['parity']['max_rel_site_logl'], d['north_star']['parity']['max_rel_site_logl'])
PY
