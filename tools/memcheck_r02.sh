# Dev tool (GPU box): compute-sanitizer memcheck over the parity tests that exercise every kernel (round 2: + TMEM-resident image
# kernel, noise, TMA windows, sample_pdf, marching cubes)
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x -k 'golden or ragged or straddle or empty or f16f8 or marching or tma or noise or sample_pdf' > gpurun_out/r02_memcheck.log 2>&1
tail -6 gpurun_out/r02_memcheck.log
