#!/bin/bash
# prebuild the variant libraries named in VARIANTS ("name:defines;name:defines") in-tree
IFS=';' read -ra VS <<< "$VARIANTS"
for spec in "${VS[@]}"; do
  v="${spec%%:*}"; defs="${spec#*:}"
  MGS_VARIANT=$v MGS_NVCC_DEFINES="$defs" python -c "from manigaussian_b200 import build; print(build.build())" &
done
wait
