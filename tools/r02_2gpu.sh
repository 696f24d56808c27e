# Dev tool (GPU box, 2 GPUs): the multi-GPU bench line incl. the strong-scaling leg.
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err
echo "rc=$?"; tail -3 gpurun_out/bench_r02_2gpu.err; wc -c gpurun_out/bench_r02_2gpu.json
