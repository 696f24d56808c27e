export DDMI_B200_IMAGE_TS=1
timeout 600 python -m pytest tests -m gpu -x -q -k "image or repeated" 2>&1 | tail -4
timeout 200 python tools/profile_image.py 2>&1 | tail -2 | head -1
timeout 200 python tools/profile_image.py 32 2048 2>&1 | tail -2 | head -1
export DDMI_B200_LIB=ddmi_b200/libddmi_b200_prof.so
timeout 200 python tools/profile_image.py 2>&1 | tail -2
DBG=2 timeout 200 python tools/profile_image.py 2>&1 | tail -2
timeout 200 python tools/profile_timeline.py > gpurun_out/r02n_timeline_ts.log 2>&1
