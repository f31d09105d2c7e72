#!/bin/bash
# usage: tools/gpu_nsplit_sweep2.sh <tag> — split count x tuning variant (B2P_VARIANT), kernel-only
tag=$1; out=gpurun_out/${tag}_nsplit_variant_sweep.txt
: > $out
run() { # kernel nsplit variant
  B2P_VARIANT=$3 python bench.py --kernel $1 --nsplit $2 --steps 64 --warmup 5 --no-e2e --no-cpu --no-ring --no-live --beamset 0 2>/dev/null |
    python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$1 variant $3 nsplit $2: chained %.5f ms %.1f GB/s   isolated %.5f ms %.1f GB/s' % (d['ms_per_step'], d['value'], r['launch_ms'], r['achieved']))" >> $out
}
for v in 0 2 3; do for ns in 37 55 74 92 111; do run ldg $ns $v; done; done
for v in 0 1 2 3; do for ns in 37 74; do run tma $ns $v; done; done
cat $out
