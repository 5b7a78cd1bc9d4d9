# eight GPUs: per-evaluation exchange fused into k_reduce over peer memory (default) against ncclAllGather (SIEVE_B200_P2P=0)
mkdir -p gpurun_out
for mode in 1 0 1 0; do
  for rep in a; do
  SIEVE_B200_COMM_DEBUG=1 SIEVE_B200_P2P=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$mode bench.py --gpus 8 --steps 40 --warmup 3 --no-north-star --no-cpu-baseline --no-variants --no-dropin --mcmc-states 600 > gpurun_out/r02ae_n8_p2p${mode}_$RANDOM.json 2> gpurun_out/r02ae_n8_p2p$mode.err
  echo "p2p=$mode rc=$?"; grep -c "fused into k_reduce" gpurun_out/r02ae_n8_p2p$mode.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02ae_n8_p2p*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f[-22:], 'value %.1f ms %.4f e2e %.1f'%(d['value'], d['ms_per_step'], d['e2e']['value']), {k:round(v['states_per_s']) for k,v in d['mcmc'].items()}, d['parity']['combine_rel_vs_oracle'], d['parity']['logP_repeatable'])
    except Exception as e:
        print(f,'ERR',e)
PY
