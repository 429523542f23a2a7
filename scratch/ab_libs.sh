#!/bin/bash
# A/B of prebuilt library variants (scratch/variants/*.so, built here with MR_BUILD_OUT / MR_NVCC_EXTRA):
#   scratch/ab_libs.sh pp0 pp1 pp2
for v in "$@"; do
  MINRMS_B200_LIB=$PWD/scratch/variants/$v.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']
print('LIB [$v] value %.0f ms/step %.3f dp_gcups_kernel %.0f dp_ms %.3f share %s rmsd100 %.6f' % (d['value'], d['ms_per_step'], r['achieved'], r['avg_launch_ms'], r['share_of_step'], d['result_check']['rmsd_100']))" || echo "LIB [$v] failed"
done
