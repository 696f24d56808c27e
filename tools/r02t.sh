timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --workload mesh --steps 3 --warmup 2 > gpurun_out/r02t_bench_mesh.json 2> gpurun_out/r02t_bench_mesh.err; tail -c 2500 gpurun_out/r02t_bench_mesh.json; tail -3 gpurun_out/r02t_bench_mesh.err
