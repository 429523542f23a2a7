#!/bin/bash
# usage: scratch/ab_refine.sh "<flags A>" "<flags B>" ...   (bench + Refine batch timing)
for f in "$@"; do
  MR_NVCC_EXTRA="$f" python -m minrms_b200.build --force > /dev/null 2>&1 || { echo "build failed: $f"; continue; }
  echo "FLAGS [$f]"
  MR_DEBUG_REFINE=1 timeout 100 python scratch/time_c1.py 2>&1 | grep "refine batch" | tail -1
  timeout 100 python scratch/time_c1.py 2>&1 | tail -2 | cut -c1-30,55-75,150-260
done
python -m minrms_b200.build --force > /dev/null 2>&1
