# two GPUs: the fused peer-memory exchange against ncclAllGather (test), then the driver's 2-GPU command both ways
mkdir -p gpurun_out
python -m pytest tests/test_gpu_comm2.py tests/test_gpu_parity.py -m gpu -x -q -k "two_rank or native_comm" 2>&1 | tail -15
for mode in 1 0; do
  S=$(date +%s)
  SIEVE_B200_COMM_DEBUG=1 SIEVE_B200_P2P=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$mode bench.py --gpus 2 --steps 20 --warmup 3 --no-north-star > gpurun_out/r02ad_n2_p2p$mode.json 2> gpurun_out/r02ad_n2_p2p$mode.err
  echo "p2p=$mode rc=$? wall=$(( $(date +%s) - S ))s"; grep "sieve_comm_init" gpurun_out/r02ad_n2_p2p$mode.err | head -2
done
python - <<'PY'
import json
for m in (1,0):
    try:
        d=json.loads(open('gpurun_out/r02ad_n2_p2p%d.json'%m).read().strip().splitlines()[-1])
        print('p2p',m,'value %.1f ms %.4f e2e %.1f'%(d['value'], d['ms_per_step'], d['e2e']['value']), {k:round(v['states_per_s']) for k,v in d['mcmc'].items()}, d['parity']['combine_rel_vs_oracle'], d['parity']['logP'], d['config']['parallelism'][:60])
    except Exception as e:
        print('p2p',m,'ERR',e, open('gpurun_out/r02ad_n2_p2p%d.err'%m).read()[-1500:])
PY
