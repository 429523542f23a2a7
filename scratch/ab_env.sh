#!/bin/bash
# A/B of runtime switches: scratch/ab_env.sh "MR_TB_OVERLAP=0" "MR_TB_SLICES=4" ...
for v in "$@"; do
  env $v python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']
print('ENV [$v] value %.0f ms/step %.3f e2e %.0f dp_ms %.3f share %s rmsd100 %.6f' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['avg_launch_ms'], r['share_of_step'], d['result_check']['rmsd_100']))" || echo "ENV [$v] failed"
done
