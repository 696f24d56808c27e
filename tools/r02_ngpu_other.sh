# Dev tool (GPU box, N GPUs): the non-headline BASELINE configs at N GPUs (weak scaling: one batch of items per GPU).
N=${1:-8}
for w in video occupancy nerf; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --workload $w --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_${w}_${N}gpu.json 2> gpurun_out/bench_r02_${w}_${N}gpu.err
  echo "$w rc=$?"; tail -2 gpurun_out/bench_r02_${w}_${N}gpu.err | cut -c1-300
  python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_${w}_${N}gpu.json').read().strip().splitlines()[-1]); print('$w', d['n_gpus'], d['value'], d['e2e']['value'], d['roofline']['frac'])"
done
