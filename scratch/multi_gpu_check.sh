#!/bin/bash
# scratch/multi_gpu_check.sh N : the multi-GPU paths on N GPUs of one box (gpurun --gpus N)
N=${1:-2}
OUT=gpurun_out
python -m pytest tests/test_gpu_cli.py -m gpu -q 2>&1 | tail -3
for sc in weak strong; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
     bench.py --gpus $N --steps 60 --warmup 5 --scaling $sc > $OUT/bench_n${N}_$sc.json 2> $OUT/bench_n${N}_$sc.err || tail -5 $OUT/bench_n${N}_$sc.err
done
# the C++ drop-in binary, one host thread per GPU (MINRMS_CUDA_GPUS): wall time of the whole command
EXE=$PWD/oracle/_ref/strict/minrms_cuda
for cfg in "-fm 4 -of 8.0,0.33 -HS -minN 0.33 -max-rmsd 3.5 -ir -r" "-fm 4 -HS -ir -r"; do
  for g in 1 $N; do
    d=$(mktemp -d); cp tests/golden/hemoglobin.pdb tests/golden/myoglobin.pdb $d/
    s=$(date +%s%N)
    (cd $d && MINRMS_CUDA_GPUS=$g $EXE $cfg "hemoglobin.pdb,*.A" myoglobin.pdb > run.log 2>&1) || echo "minrms_cuda failed"
    e=$(date +%s%N)
    echo "CLI [$cfg] MINRMS_CUDA_GPUS=$g: $(( (e - s) / 1000000 )) ms wall, $(ls $d/*.msf | wc -l) msf files, md5 of all msf bodies $(cat $d/align*.msf | grep -v 'MSF:' | md5sum | cut -c1-12)"
    rm -rf $d
  done
done
