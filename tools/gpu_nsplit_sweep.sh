#!/bin/bash
# usage: tools/gpu_nsplit_sweep.sh <tag> — kernel-only step time against the time-split count, both kernels
tag=$1; out=gpurun_out/${tag}_nsplit_sweep.txt
: > $out
for k in tma ldg; do
  for ns in 37 74 111 148 185 222 259 296 444; do
    python bench.py --kernel $k --nsplit $ns --steps 64 --warmup 5 --no-e2e --no-cpu --no-ring --no-live --beamset 0 2>/dev/null |
      python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$k nsplit $ns: chained %.5f ms %.1f GB/s   isolated %.5f ms %.1f GB/s' % (d['ms_per_step'], d['value'], r['launch_ms'], r['achieved']))" >> $out
  done
done
cat $out
