# Dev tool: the round-end evidence run (on the GPU box, via gpurun): tests, smoke, benches, ncu launch list + full captures.
set -x
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -8
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err
for w in occupancy video nerf; do timeout 200 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_r01e_$w.json 2>/dev/null; done
timeout 300 python bench.py --precision bf16x3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01e_bf16x3.json 2>/dev/null
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01e_ref.json 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r01e_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01e_launches_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:image_umma -s 1 -c 1 -o gpurun_out/r01e_image_f16f8 python bench.py --batch 8 --res 1024 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu -i gpurun_out/r01e_image_f16f8.ncu-rep --page raw --csv > gpurun_out/r01e_image_f16f8_raw.csv
for w in occupancy video nerf; do
  timeout 400 ncu --set full --clock-control none -k regex:${w}_umma -s 1 -c 1 -o /tmp/r01e_${w} python bench.py --workload $w --steps 1 --warmup 1 > /dev/null 2>&1
  ncu -i /tmp/r01e_${w}.ncu-rep --page raw --csv > gpurun_out/r01e_${w}_f16f8_raw.csv
done
ls -la gpurun_out/
