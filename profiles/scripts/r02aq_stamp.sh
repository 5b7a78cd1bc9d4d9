mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['traffic'], d['roofline']['traffic_source'], d['library_build_id'])"
