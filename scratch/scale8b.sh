#!/bin/bash
# scratch/scale8b.sh N : the filter-heavy configurations on N GPUs (gpurun --gpus N)
N=${1:-8}
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641"
$TR bench.py --gpus $N --config 1 --steps 200 --warmup 5 --scaling weak > $OUT/r02f_n${N}_c1_weak.json 2> $OUT/r02f_n${N}_c1_weak.err || tail -3 $OUT/r02f_n${N}_c1_weak.err
$TR bench.py --gpus $N --config 4 --steps 2 --warmup 3 --scaling strong > $OUT/r02f_n${N}_c4_strong.json 2> $OUT/r02f_n${N}_c4_strong.err || tail -3 $OUT/r02f_n${N}_c4_strong.err
for f in $OUT/r02f_n${N}_c1_weak.json $OUT/r02f_n${N}_c4_strong.json; do python -c "
import json,sys
l=[x for x in open(sys.argv[1]).read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]); print(sys.argv[1], d['value'], d['ms_per_step'], d.get('reduction_ms'), (d.get('e2e') or {}).get('value'), d.get('result_check'))" $f; done
