# Dev tool (GPU box): does a deeper weight ring help?  (experiment build: results garbage, timing only)
for lib in ddmi_b200/libddmi_b200.so ddmi_b200/libddmi_b200_exp.so; do
  DDMI_B200_LIB=$lib timeout 200 python tools/profile_image.py 2>&1 | tail -2
done
