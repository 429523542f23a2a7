#!/bin/bash
# usage: scratch/ab.sh "<flags A>" "<flags B>" ...
for f in "$@"; do
  MR_NVCC_EXTRA="$f" python -m minrms_b200.build --force > /dev/null 2>&1 || { echo "build failed: $f"; continue; }
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']
print('FLAGS [$f] value %.0f ms/step %.2f dp_gcups_kernel %.0f dp_ms %.2f share %s rmsd100 %.6f' % (d['value'], d['ms_per_step'], r['achieved'], r['avg_launch_ms']*2, r['share_of_step'], d['result_check']['rmsd_100']))" || echo "FLAGS [$f] failed"
done
python -m minrms_b200.build --force > /dev/null 2>&1
