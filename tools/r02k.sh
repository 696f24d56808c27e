# TMEM-resident-activation image kernel (DDMI_B200_IMAGE_TS=1): parity, then device-resident timing vs the shared-memory kernel
export DDMI_B200_IMAGE_TS=1
for ns in full mixed; do
  echo "=== TS NSPLIT=$ns"
  DDMI_B200_NSPLIT=$ns timeout 600 python -m pytest tests -m gpu -x -q -k "image or repeated" 2>&1 | tail -6
  DDMI_B200_NSPLIT=$ns timeout 200 python tools/profile_image.py 2>&1 | tail -2 | head -1
  DDMI_B200_NSPLIT=$ns timeout 200 python tools/profile_image.py 32 2048 2>&1 | tail -2 | head -1
done
echo "=== SS (reference point)"
DDMI_B200_IMAGE_TS=0 timeout 200 python tools/profile_image.py 2>&1 | tail -2 | head -1
