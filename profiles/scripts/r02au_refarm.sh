# the reference arm exactly as the driver launches it (N = 1 and under torchrun with N = 2: rank 0 prints, the other exits 0)
mkdir -p gpurun_out
S=$(date +%s); python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r02au_ref_n1.json 2> gpurun_out/r02au_ref_n1.err; echo "rc=$? wall=$(( $(date +%s) - S ))s"
python -c "
import json; d=json.loads(open('gpurun_out/r02au_ref_n1.json').read().strip().splitlines()[-1]); print(d['impl'], d['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['sample'][:80], d['config'])"
S=$(date +%s); python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02au_ref_n2.json 2> gpurun_out/r02au_ref_n2.err; echo "rc=$? wall=$(( $(date +%s) - S ))s"
grep -c '"impl"' gpurun_out/r02au_ref_n2.json
