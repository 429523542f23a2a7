#!/bin/bash
OUT=gpurun_out; T=r02g
python bench.py > $OUT/${T}_n1_c2.json 2> $OUT/${T}_n1_c2.err || tail -3 $OUT/${T}_n1_c2.err
python bench.py --config 1 > $OUT/${T}_n1_c1.json 2> $OUT/${T}_n1_c1.err || tail -3 $OUT/${T}_n1_c1.err
python bench.py --config 4 --steps 3 --warmup 3 > $OUT/${T}_n1_c4.json 2> $OUT/${T}_n1_c4.err || tail -3 $OUT/${T}_n1_c4.err
for f in $OUT/${T}_n1_*.json; do python - "$f" <<'PY'
import json, sys
lines = [l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")]
if not lines: print(sys.argv[1], "no json"); sys.exit(0)
d = json.loads(lines[-1]); r = d.get("roofline") or {}
print(sys.argv[1].split("/")[-1], "value %.0f %s, ms/step %.2f, e2e %s, dp ms %s frac %s, full_run %s, nw %s" % (
    d.get("value", 0), d.get("unit"), d.get("ms_per_step", 0), (d.get("e2e") or {}).get("value"),
    r.get("avg_launch_ms"), r.get("frac"), (d.get("full_run") or {}).get("ms"), (d.get("nw_filter") or {}).get("avg_launch_ms")))
PY
done
