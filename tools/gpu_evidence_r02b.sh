# Round-2 (second half) evidence run (1 GPU): smoke, every bench line, reference arm, ncu launch list of the bench command,
# ncu --set full of the video and occupancy kernels (summarised on the box: the reports are too big to bring back)
set -x
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02b_smoke.log 2>&1; tail -3 gpurun_out/r02b_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; tail -2 gpurun_out/bench_r02.err
for w in c1 video occupancy nerf mesh planes; do timeout 400 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_r02_$w.json 2> gpurun_out/bench_r02_$w.err; tail -2 gpurun_out/bench_r02_$w.err; done
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_ref.json 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02b_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02b_launches_bench.log 2>&1
for w in video occupancy nerf; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:${w}_umma -s 1 -c 1 -o /tmp/r02b_$w python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  python tools/ncu_stalls.py /tmp/r02b_$w.ncu-rep > gpurun_out/r02b_${w}_ncu_stalls.md
  ncu -i /tmp/r02b_$w.ncu-rep --page raw --csv > gpurun_out/r02b_${w}_raw.csv
done
# the occupancy capture above is the point-list launch (-s 1); the lattice launch of the timed step is the third one
timeout 500 ncu --set full --clock-control none --import-source on -k regex:occupancy_umma -s 2 -c 1 -o /tmp/r02b_occ python bench.py --workload occupancy --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_stalls.py /tmp/r02b_occ.ncu-rep > gpurun_out/r02b_occupancy_lattice_ncu_stalls.md
ncu -i /tmp/r02b_occ.ncu-rep --page raw --csv > gpurun_out/r02b_occupancy_lattice_raw.csv
# timelines of one tile (profiling build) and memcheck of the kernels changed in the second half of round 2
for k in video nerf; do KIND=$k DDMI_B200_LIB=$PWD/ddmi_b200/libddmi_b200_prof.so timeout 120 python tools/profile_timeline.py > gpurun_out/r02b_${k}_timeline.txt 2>&1; done
KIND=occ PTS=lattice DDMI_B200_LIB=$PWD/ddmi_b200/libddmi_b200_prof.so timeout 120 python tools/profile_timeline.py > gpurun_out/r02b_occupancy_timeline_lattice.txt 2>&1
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x -k 'video or occupancy or lattice or mesh or generat' > gpurun_out/r02b_memcheck.log 2>&1; tail -3 gpurun_out/r02b_memcheck.log
ls -la gpurun_out | head -40
