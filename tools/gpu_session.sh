#!/bin/bash
# usage: tools/gpu_session.sh <tag> <ngpus> — box facts, GPU tests of the stage, bench at N=ngpus
tag=$1; n=$2
out=gpurun_out
{ nproc; free -g | head -2; ipcs -lm | head -8; df -h /dev/shm | tail -1; nvidia-smi topo -m; } > $out/${tag}_box.txt 2>&1
python -m pytest tests/test_host_ring.py -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_stagetest.log
if [ "$n" -gt 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err
else
  python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
fi
echo "rc=$?"
tail -3 $out/${tag}_stagetest.log
tail -5 $out/${tag}_bench_n$n.err
head -c 6000 $out/${tag}_bench_n$n.json
