for w in occupancy video nerf; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:${w}_umma -s 1 -c 1 -o /tmp/r02s_${w} python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  python tools/ncu_stalls.py /tmp/r02s_${w}.ncu-rep > gpurun_out/r02s_${w}_ncu.md
  ncu -i /tmp/r02s_${w}.ncu-rep --page raw --csv > gpurun_out/r02s_${w}_raw.csv
done
ls -la gpurun_out/
