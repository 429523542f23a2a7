#!/bin/bash
# usage (GPU box): scratch/ab_refine_variants.sh A X ...  -- Refine timings per library variant
cp minrms_b200/libminrms_b200.so /tmp/lib_A.so
for v in "$@"; do
  if [ "$v" = A ]; then cp /tmp/lib_A.so minrms_b200/libminrms_b200.so; else cp scratch/variants/lib_$v.so minrms_b200/libminrms_b200.so; fi
  for sz in "300 280" "380 400"; do echo "VARIANT $v $(python scratch/refine_pair.py $sz | tail -1)"; done
done
cp /tmp/lib_A.so minrms_b200/libminrms_b200.so
