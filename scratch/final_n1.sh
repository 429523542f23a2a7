#!/bin/bash
# scratch/final_n1.sh : bench lines of one GPU for the round's profiles (configs 1-5 + the reference arm)
OUT=gpurun_out
T=${1:-r02f}
python bench.py > $OUT/${T}_n1_c2.json 2> $OUT/${T}_n1_c2.err || tail -3 $OUT/${T}_n1_c2.err
python bench.py --impl reference > $OUT/${T}_n1_c2_ref.json 2> $OUT/${T}_n1_c2_ref.err || tail -3 $OUT/${T}_n1_c2_ref.err
python bench.py --config 1 > $OUT/${T}_n1_c1.json 2> $OUT/${T}_n1_c1.err || tail -3 $OUT/${T}_n1_c1.err
python bench.py --config 3 --steps 2 --warmup 3 > $OUT/${T}_n1_c3.json 2> $OUT/${T}_n1_c3.err || tail -3 $OUT/${T}_n1_c3.err
python bench.py --config 4 --steps 3 --warmup 3 > $OUT/${T}_n1_c4.json 2> $OUT/${T}_n1_c4.err || tail -3 $OUT/${T}_n1_c4.err
python bench.py --config 5 --steps 1 > $OUT/${T}_n1_c5.json 2> $OUT/${T}_n1_c5.err || tail -3 $OUT/${T}_n1_c5.err
for f in $OUT/${T}_n1_*.json; do python - "$f" <<'PY'
import json, sys
lines = [l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")]
if not lines: print(sys.argv[1], "no json"); sys.exit(0)
d = json.loads(lines[-1]); r = d.get("roofline") or {}
print(sys.argv[1].split("/")[-1], "value %.0f %s, ms/step %.2f, e2e %s, dp ms %s frac %s, full_run %s" % (
    d.get("value", 0), d.get("unit"), d.get("ms_per_step", 0), (d.get("e2e") or {}).get("value"),
    r.get("avg_launch_ms"), r.get("frac"), (d.get("full_run") or {}).get("ms")))
PY
done
