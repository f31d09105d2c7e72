#!/bin/bash
# usage: tools/gpu_nsplit_ab.sh <tag> — interleaved A/B of split counts on ONE box (rules out drift)
tag=$1; out=gpurun_out/${tag}_nsplit_ab.txt
: > $out
for round in 1 2 3; do
  for ns in 74 222 111 148; do
    python bench.py --kernel ldg --nsplit $ns --steps 128 --warmup 8 --no-e2e --no-cpu --no-ring --no-live --beamset 0 2>/dev/null |
      python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('round $round ldg nsplit $ns: chained %.5f ms %.1f GB/s   isolated %.5f ms %.1f GB/s  clocks %s' % (d['ms_per_step'], d['value'], r['launch_ms'], r['achieved'], d['clocks']['sm_mhz']))" >> $out
  done
done
cat $out
