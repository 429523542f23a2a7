#!/bin/bash
# usage (on the GPU box): scratch/ab_variants.sh A B C ...   (A = the library as built; others = scratch/variants/lib_X.so)
cp minrms_b200/libminrms_b200.so /tmp/lib_A.so
for v in "$@"; do
  if [ "$v" = A ]; then cp /tmp/lib_A.so minrms_b200/libminrms_b200.so; else cp scratch/variants/lib_$v.so minrms_b200/libminrms_b200.so; fi
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']
print('VARIANT $v value %.0f ms/step %.2f dp_gcups_kernel %.0f dp_launch_ms %.3f x%d share %s rmsd100 %.6f' % (d['value'], d['ms_per_step'], r['achieved'], r['avg_launch_ms'], r['launches'], r['share_of_step'], d['result_check']['rmsd_100']))" || echo "VARIANT $v failed"
done
cp /tmp/lib_A.so minrms_b200/libminrms_b200.so
