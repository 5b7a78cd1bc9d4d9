# k_walk<., GEN>: the lean walk for 16-double matrices (drop-in call sequence); bit-identity against k_prune, then the drop-in number
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SIEVE_B200_GENERIC=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('generic k_walk: value %.1f prune %.3f'%(d['value'], d['kernels']['prune_ms']), d['parity']['max_rel_site_logl'])"
SIEVE_B200_GENERIC=1 SIEVE_B200_WALK=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-dropin --no-north-star 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('generic k_prune: value %.1f prune %.3f'%(d['value'], d['kernels']['prune_ms']))"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --mcmc-states 0 --no-variants --no-north-star 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.1f e2e %.1f dropin %.1f (%.3f ms)'%(d['value'], d['e2e']['value'], d['e2e_dropin']['value'], d['e2e_dropin']['ms_per_step']))"
