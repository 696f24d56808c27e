# Dev tool (GPU box): every bench line once (short).
for w in c1 video occupancy nerf; do timeout 400 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/r02e_bench_$w.json 2> gpurun_out/r02e_bench_$w.err; tail -3 gpurun_out/r02e_bench_$w.err; done
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02e_bench_image.json 2> gpurun_out/r02e_bench_image.err; tail -3 gpurun_out/r02e_bench_image.err
