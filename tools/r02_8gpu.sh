# Dev tool (GPU box, 8 GPUs): the multi-GPU bench line incl. the strong-scaling leg.
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_r02_8gpu.json 2> gpurun_out/bench_r02_8gpu.err
echo "rc=$?"; tail -3 gpurun_out/bench_r02_8gpu.err; wc -c gpurun_out/bench_r02_8gpu.json
