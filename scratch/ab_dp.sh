#!/bin/bash
# DP-kernel-only timing A/B (results invalid): scratch/ab_dp.sh "<flags A>" "<flags B>" ...
for f in "$@"; do
  MR_NVCC_EXTRA="$f" python -m minrms_b200.build --force > /dev/null 2>&1 || { echo "build failed: $f"; continue; }
  MR_EXP_DP_ONLY=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']
print('FLAGS [$f] ms/step %.2f dp_gcups_kernel %.0f avg_launch_ms %.3f' % (d['ms_per_step'], r['achieved'], r['avg_launch_ms']))" || echo "FLAGS [$f] failed"
done
python -m minrms_b200.build --force > /dev/null 2>&1
