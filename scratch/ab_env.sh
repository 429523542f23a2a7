#!/bin/bash
# A/B of runtime switches: scratch/ab_env.sh "MINRMS_B200_TB_PAIRS=0" "MINRMS_B200_TB_PAIRS=1" ...
for v in "$@"; do
  env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; sh=r['share_of_step']
print('ENV [$v] value %.0f ms/step %.3f e2e %.0f dp_ms %.3f tb_ms %.3f rmsd100 %.6f' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['avg_launch_ms'], sh['traceback_kabsch']*d['ms_per_step'], d['result_check']['rmsd_100']))" || echo "ENV [$v] failed"
done
