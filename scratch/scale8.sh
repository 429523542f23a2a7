#!/bin/bash
# scratch/scale8.sh N : scaling measurements on N GPUs of one box (gpurun --gpus N)
N=${1:-8}
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621"
$TR bench.py --gpus $N --steps 100 --warmup 5 --scaling weak   > $OUT/r02f_n${N}_c2_weak.json   2> $OUT/r02f_n${N}_c2_weak.err   || tail -3 $OUT/r02f_n${N}_c2_weak.err
$TR bench.py --gpus $N --steps 100 --warmup 5 --scaling strong > $OUT/r02f_n${N}_c2_strong.json 2> $OUT/r02f_n${N}_c2_strong.err || tail -3 $OUT/r02f_n${N}_c2_strong.err
$TR bench.py --gpus $N --config 3 --steps 3 --warmup 3 --scaling strong > $OUT/r02f_n${N}_c3_strong.json 2> $OUT/r02f_n${N}_c3_strong.err || tail -3 $OUT/r02f_n${N}_c3_strong.err
$TR bench.py --gpus $N --config 5 --chains 64 --refine --workers 4 --steps 1 --warmup-batches 0 > $OUT/r02f_n${N}_c5.json 2> $OUT/r02f_n${N}_c5.err || tail -3 $OUT/r02f_n${N}_c5.err
EXE=$PWD/oracle/_ref/strict/minrms_cuda
for g in 1 2 4 $N; do
  d=$(mktemp -d); cp tests/golden/hemoglobin.pdb tests/golden/myoglobin.pdb $d/
  s=$(date +%s%N)
  (cd $d && MINRMS_CUDA_GPUS=$g $EXE -fm 4 -HS -ir -r "hemoglobin.pdb,*.A" myoglobin.pdb > run.log 2>&1) || echo "minrms_cuda failed"
  e=$(date +%s%N)
  echo "CLI config 2, MINRMS_CUDA_GPUS=$g: $(( (e - s) / 1000000 )) ms wall, $(ls $d/*.msf | wc -l) msf files, md5 $(cat $d/align*.msf | grep -v 'MSF:' | md5sum | cut -c1-12)" | tee -a $OUT/r02_cli_gpus.log
  rm -rf $d
done
