# Dev tool (GPU box): what-if bounds of the image kernel pipeline (profiling build).
export DDMI_B200_LIB=ddmi_b200/libddmi_b200_prof.so
for d in 0 1 2 4 5 6 7; do DBG=$d timeout 200 python tools/profile_image.py 2>&1 | tail -2; done
