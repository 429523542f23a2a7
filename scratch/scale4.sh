#!/bin/bash
# scratch/scale4.sh N : weak and strong scaling of config 2 on N GPUs (gpurun --gpus N)
N=${1:-4}
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631"
$TR bench.py --gpus $N --steps 100 --warmup 5 --scaling weak   > $OUT/r02f_n${N}_c2_weak.json   2> $OUT/r02f_n${N}_c2_weak.err   || tail -3 $OUT/r02f_n${N}_c2_weak.err
$TR bench.py --gpus $N --steps 100 --warmup 5 --scaling strong > $OUT/r02f_n${N}_c2_strong.json 2> $OUT/r02f_n${N}_c2_strong.err || tail -3 $OUT/r02f_n${N}_c2_strong.err
for f in $OUT/r02f_n${N}_c2_weak.json $OUT/r02f_n${N}_c2_strong.json; do python -c "
import json,sys
l=[x for x in open(sys.argv[1]).read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]); print(sys.argv[1], d['value'], d['ms_per_step'], d.get('reduction_ms'), (d.get('e2e') or {}).get('value'), d.get('result_check'))" $f; done
