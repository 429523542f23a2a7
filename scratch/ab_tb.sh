#!/bin/bash
# A/B of library variants on the traceback + Kabsch kernel: scratch/ab_tb.sh tb8 tb10 ... ("main" = the in-tree library)
for v in "$@"; do
  if [ "$v" = main ]; then unset MINRMS_B200_LIB; else export MINRMS_B200_LIB=$PWD/scratch/variants/$v.so; fi
  python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; sh=r['share_of_step']
print('LIB [$v] value %.0f ms/step %.3f dp_ms %.3f tb_ms %.3f rmsd100 %.6f' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], sh['traceback_kabsch']*d['ms_per_step'], d['result_check']['rmsd_100']))" || echo "LIB [$v] failed"
done
