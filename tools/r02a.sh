# Dev tool (GPU box): round-2 first trip -- parity tests, epilogue microbench, stage timeline, baseline bench.
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02a_tests.log
timeout 200 python tools/microbench.py > gpurun_out/r02a_microbench.log 2>&1
DDMI_B200_LIB=ddmi_b200/libddmi_b200_prof.so timeout 200 python tools/profile_timeline.py > gpurun_out/r02a_timeline.log 2>&1
DDMI_B200_LIB=ddmi_b200/libddmi_b200_prof.so timeout 200 python tools/profile_image.py > gpurun_out/r02a_profimg.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -3 gpurun_out/r02a_tests.log; cat gpurun_out/r02a_microbench.log
