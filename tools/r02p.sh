export DDMI_B200_IMAGE_TS=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:image_umma -s 1 -c 1 -o gpurun_out/r02p_image_ts python bench.py --batch 8 --res 1024 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02p_ncu.log 2>&1
ls -la gpurun_out/r02p_image_ts.ncu-rep
