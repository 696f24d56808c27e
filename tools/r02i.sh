# N-split A/B: parity tests per mode, then device-resident timing (shipping lib) and in-kernel counters (prof lib)
for ns in mixed full off; do
  echo "=== NSPLIT=$ns"
  DDMI_B200_NSPLIT=$ns timeout 900 python -m pytest tests -m gpu -x -q -k "image or repeated" 2>&1 | tail -3
  DDMI_B200_NSPLIT=$ns timeout 200 python tools/profile_image.py 2>&1 | tail -2 | head -1
  DDMI_B200_NSPLIT=$ns timeout 200 python tools/profile_image.py 32 2048 2>&1 | tail -2 | head -1
  DDMI_B200_NSPLIT=$ns DDMI_B200_LIB=ddmi_b200/libddmi_b200_prof.so timeout 200 python tools/profile_image.py 2>&1 | tail -2
done
