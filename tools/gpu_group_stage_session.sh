#!/bin/bash
# usage: tools/gpu_group_stage_session.sh <tag> — on a multi-GPU box: one beam, ONE stage process,
# channel groups over 8 / 4 / 2 GPUs through the rings (64 integrations each)
tag=$1; out=gpurun_out
n=$(nvidia-smi -L | wc -l)
export OMP_NUM_THREADS=$(nproc)
for g in 0,1,2,3,4,5,6,7 0,2,4,6 0,4; do
  k=$(echo $g | tr ',' '\n' | wc -l)
  [ $k -le $n ] || continue
  timeout 200 python tools/run_ring_e2e.py --gpu $g --nblocks 64 --nbufs 3 > $out/${tag}_group_$k.json 2> $out/${tag}_group_$k.err
  echo "gpus $g rc=$?"; cat $out/${tag}_group_$k.json; tail -2 $out/${tag}_group_$k.err
done
