export DDMI_B200_LIB=ddmi_b200/libddmi_b200_prof.so
echo "=== SS off"; DDMI_B200_IMAGE_TS=0 timeout 200 python tools/profile_image.py 2>&1 | tail -2
for ns in full mixed; do
  echo "=== TS NSPLIT=$ns"
  DDMI_B200_IMAGE_TS=1 DDMI_B200_NSPLIT=$ns timeout 200 python tools/profile_image.py 2>&1 | tail -2
  for d in 1 2 3; do DBG=$d DDMI_B200_IMAGE_TS=1 DDMI_B200_NSPLIT=$ns timeout 200 python tools/profile_image.py 2>&1 | tail -2; done
done
DDMI_B200_IMAGE_TS=1 DDMI_B200_NSPLIT=full timeout 200 python tools/profile_timeline.py > gpurun_out/r02m_timeline_ts_full.log 2>&1
DDMI_B200_IMAGE_TS=0 timeout 200 python tools/profile_timeline.py > gpurun_out/r02m_timeline_ss_off.log 2>&1
